"""TEST INFRASTRUCTURE: ctypes front-end of the host emulation build (tests/host_emu/bt_emu.cpp) -- the same
per-environment programs as the CUDA kernels, one lane per environment.  Never imported by the product."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_SRC = os.path.join(_HERE, "host_emu", "bt_emu.cpp")
_LIB = os.path.join(_HERE, "host_emu", "libbt_emu.so")
_CSRC = os.path.join(_ROOT, "brax_tracking_b200", "csrc")


class StatePtrs(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("qpos", "qvel", "act", "qacc_warmstart", "time", "xpos")]


def build(force=False):
    deps = [_SRC] + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith(".h")] + [os.path.join(_ROOT, "include", "bt_api.h")]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(d) > os.path.getmtime(_LIB) for d in deps):
        subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off", "-I" + os.path.join(_ROOT, "include"), "-I" + _CSRC,
                        "-o", _LIB, _SRC], check=True)
    return _LIB


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def table_args(tables):
    names = list(tables.keys())
    arrs = [np.ascontiguousarray(tables[k]) for k in names]
    n = len(names)
    c_names = (C.c_char_p * n)(*[k.encode() for k in names])
    c_data = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    c_counts = (C.c_int64 * n)(*[a.size for a in arrs])
    c_isf = (C.c_int * n)(*[1 if a.dtype == np.float32 else 0 for a in arrs])
    for a in arrs:
        assert a.dtype in (np.float32, np.int32), a.dtype
    return n, c_names, c_data, c_counts, c_isf, arrs


class Emu:
    def __init__(self, tables):
        self.lib = C.CDLL(build())
        self.lib.emu_last_error.restype = C.c_char_p
        self.t = tables
        g = lambda k: int(tables[k][0])
        self.nq, self.nv, self.nu, self.na, self.nbody = g("nq"), g("nv"), g("nu"), g("na"), g("nbody")
        self.obs_size, self.smem, self.ncon = g("obs_size"), g("smem_floats"), g("ncon")
        n, names, data, counts, isf, keep = table_args(tables)
        self._keep = keep
        h = C.c_void_p()
        rc = self.lib.emu_model_create(n, names, data, counts, isf, C.byref(h))
        if rc:
            raise RuntimeError(self.lib.emu_last_error().decode())
        self.h = h

    def off(self, name):
        return int(self.t["o_" + name][0])

    # ---- state helpers (numpy float32, [N, dim]) ----
    def new_state(self, N):
        z = lambda d: np.zeros((N, d), dtype=np.float32)
        return dict(qpos=z(self.nq), qvel=z(self.nv), act=z(max(self.na, 1))[:, :self.na].copy(), qacc_warmstart=z(self.nv),
                    time=np.zeros(N, np.float32), xpos=z(3 * self.nbody))

    def sp(self, st):
        s = StatePtrs()
        for k in ("qpos", "qvel", "act", "qacc_warmstart", "time", "xpos"):
            a = st[k]
            assert a.dtype == np.float32 and a.flags.c_contiguous
            setattr(s, k, a.ctypes.data if a.size else None)
        return s

    def new_outputs(self, N):
        return dict(obs=np.zeros((N, self.obs_size), np.float32), reward=np.zeros(N, np.float32), done=np.zeros(N, np.float32),
                    metrics=np.zeros((N, 12), np.float32), info_f=np.zeros((N, 5), np.float32), info_i=np.zeros((N, 2), np.int32))

    def reset(self, keys, fixed_start_frame=-1, clip_idx=None):
        keys = np.ascontiguousarray(keys, dtype=np.uint32)
        N = keys.shape[0]
        st, out = self.new_state(N), self.new_outputs(N)
        self.lib.emu_reset(self.h, N, _ptr(keys), int(fixed_start_frame), self.sp(st), _ptr(out["obs"]), _ptr(out["reward"]), _ptr(out["done"]),
                           _ptr(out["metrics"]), _ptr(out["info_f"]), _ptr(out["info_i"]), _ptr(clip_idx))
        return st, out

    def step(self, st, out, first_st, first_obs, first_info_i, action, clip_idx=None):
        action = np.ascontiguousarray(action, dtype=np.float32)
        N = action.shape[0]
        self.lib.emu_step(self.h, N, _ptr(action), self.sp(st), self.sp(first_st), _ptr(first_obs), _ptr(first_info_i),
                          _ptr(out["obs"]), _ptr(out["reward"]), _ptr(out["done"]), _ptr(out["metrics"]), _ptr(out["info_f"]),
                          _ptr(out["info_i"]), _ptr(clip_idx))

    def physics_step(self, st, ctrl, n_substeps):
        ctrl = None if ctrl is None else np.ascontiguousarray(ctrl, dtype=np.float32)
        self.lib.emu_physics_step(self.h, st["qpos"].shape[0], _ptr(ctrl), self.sp(st), n_substeps)

    def reward_obs(self, st, out, action, clip_idx=None):
        action = np.ascontiguousarray(action, dtype=np.float32)
        self.lib.emu_reward_obs(self.h, action.shape[0], _ptr(action), self.sp(st), _ptr(out["info_i"]), _ptr(out["obs"]),
                                _ptr(out["reward"]), _ptr(out["done"]), _ptr(out["metrics"]), _ptr(out["info_f"]), _ptr(clip_idx))

    def forward_debug(self, st, ctrl, stop=0):
        N = st["qpos"].shape[0]
        ctrl = None if ctrl is None else np.ascontiguousarray(ctrl, dtype=np.float32)
        scratch = np.zeros((N, self.smem), np.float32)
        cdist = np.zeros((N, max(self.ncon, 1)), np.float32)
        niter = np.zeros(N, np.int32)
        self.lib.emu_forward_debug(self.h, N, _ptr(ctrl), self.sp(st), stop, _ptr(scratch), _ptr(cdist), _ptr(niter))
        return scratch, cdist, niter
