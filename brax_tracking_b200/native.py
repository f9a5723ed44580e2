"""ctypes binding of the C ABI (include/bt_api.h, libbt_b200.so).

PyTorch is used only for device memory and streams: every call hands raw device pointers and the current
CUDA stream to the library.  There is NO fallback: if the shared library is missing or a call fails, this
module raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
# BT_B200_LIB: another build of the same library (compiler-flag experiments); no fallback either way
LIB_PATH = os.environ.get("BT_B200_LIB") or os.path.join(_PKG, "libbt_b200.so")

NUM_METRICS, NUM_INFO_F, NUM_INFO_I = 12, 5, 2
METRIC_NAMES = ["pos_reward", "quat_reward", "joint_reward", "angvel_reward", "bodypos_reward", "endeff_reward",
                "reward_quadctrl", "reward_alive", "too_far", "bad_pose", "bad_quat", "fall"]  # fruitfly.py:481-494
INFO_F_NAMES = ["summed_pos_distance", "quat_distance", "joint_distance", "steps", "truncation"]
INFO_I_NAMES = ["cur_frame", "steps_taken_cur_frame"]
STATE_FIELDS = ("qpos", "qvel", "act", "qacc_warmstart", "time", "xpos")

# exported symbols, exactly as declared in include/bt_api.h
SYMBOLS = ["bt_model_create", "bt_model_destroy", "bt_model_dims", "bt_model_launch", "bt_reset", "bt_step",
           "bt_physics_step", "bt_pipeline_init", "bt_reward_obs", "bt_forward_debug", "bt_last_error", "bt_launch_count",
           "bt_ppo_tanh_normal_fwd", "bt_ppo_tanh_normal_bwd", "bt_ppo_flat_adam"]


class StatePtrs(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in STATE_FIELDS]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU or PyTorch fallback for the tracking step)")
        l = C.CDLL(LIB_PATH)
        l.bt_last_error.restype = C.c_char_p
        l.bt_launch_count.restype = C.c_int64
        for s in SYMBOLS:
            getattr(l, s)
        _lib = l
    return _lib


def _check(rc):
    if rc != 0:
        raise RuntimeError(f"libbt_b200: error {rc}: {lib().bt_last_error().decode()}")


def launch_count() -> int:
    return int(lib().bt_launch_count())


def _bt_strides(t):
    """(b, t) element strides of a [B, T, A] tensor whose innermost dim is contiguous."""
    if t.dim() != 3 or (t.shape[2] > 1 and t.stride(2) != 1):
        raise ValueError(f"expected a [B, T, A] tensor with a contiguous innermost dim, got strides {t.stride()}")
    return C.c_int64(t.stride(0)), C.c_int64(t.stride(1))


def ppo_tanh_normal(logits, raw, noise, grads=None):
    """bt_ppo_tanh_normal_fwd (grads None): returns (log_prob, entropy_term) as [T, B]; bt_ppo_tanh_normal_bwd
    (grads = (glp, gent), contiguous [T, B]): returns d/dlogits [B, T, 2A].  CUDA float32 tensors, current stream."""
    import torch
    B, T, A2 = logits.shape
    A = A2 // 2
    if not (logits.is_cuda and logits.is_contiguous() and logits.dtype == torch.float32 and raw.dtype == torch.float32
            and noise.dtype == torch.float32 and tuple(raw.shape) == (B, T, A) and tuple(noise.shape) == (B, T, A)):
        raise ValueError("ppo_tanh_normal: logits [B, T, 2A] contiguous float32 on CUDA, raw / noise [B, T, A] float32")
    stream = C.c_void_p(torch.cuda.current_stream(logits.device).cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    osb, ost = C.c_int64(1), C.c_int64(B)                                    # outputs / output gradients are [T, B] contiguous
    if grads is None:
        lp, ent = torch.empty(T, B, device=logits.device), torch.empty(T, B, device=logits.device)
        _check(lib().bt_ppo_tanh_normal_fwd(B, T, A, p(logits), p(raw), *_bt_strides(raw), p(noise), *_bt_strides(noise), p(lp), p(ent),
                                            osb, ost, stream))
        return lp, ent
    glp, gent = grads
    if not (glp.is_contiguous() and gent.is_contiguous() and tuple(glp.shape) == (T, B) and tuple(gent.shape) == (T, B)):
        raise ValueError("ppo_tanh_normal: output gradients must be contiguous [T, B]")
    out = torch.empty_like(logits)
    _check(lib().bt_ppo_tanh_normal_bwd(B, T, A, p(logits), p(raw), *_bt_strides(raw), p(noise), *_bt_strides(noise), p(glp), p(gent),
                                        osb, ost, p(out), stream))
    return out


def ppo_flat_adam(p, g, m, v, step, lr: float, b1: float, b2: float, eps: float, gscale: float):
    """bt_ppo_flat_adam: one fused optax.adam update of the flat parameter buffer ``p`` from the flat (summed) gradient ``g`` scaled by
    ``gscale`` (= 1 / world); ``step`` is a 0-dim CUDA float tensor (updates applied so far; advanced by the caller)."""
    import torch
    for t in (p, g, m, v):
        if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32 and t.numel() == p.numel()):
            raise ValueError("ppo_flat_adam: p, g, m, v must be contiguous float32 CUDA tensors of one size")
    stream = C.c_void_p(torch.cuda.current_stream(p.device).cuda_stream)
    _check(lib().bt_ppo_flat_adam(C.c_int64(p.numel()), C.c_void_p(p.data_ptr()), C.c_void_p(g.data_ptr()), C.c_void_p(m.data_ptr()),
                                  C.c_void_p(v.data_ptr()), C.c_void_p(step.data_ptr()), C.c_float(lr), C.c_float(b1), C.c_float(b2),
                                  C.c_float(eps), C.c_float(gscale), stream))


class NativeModel:
    """Owns a ``BtModel*`` (device copy of the packed tables)."""

    def __init__(self, tables: Dict[str, np.ndarray], device: int = 0):
        l = lib()
        names = list(tables.keys())
        arrs = [np.ascontiguousarray(tables[k]) for k in names]
        for k, a in zip(names, arrs):
            if a.dtype not in (np.float32, np.int32):
                raise TypeError(f"table {k} has dtype {a.dtype}")
        n = len(names)
        c_names = (C.c_char_p * n)(*[k.encode() for k in names])
        c_data = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        c_counts = (C.c_int64 * n)(*[a.size for a in arrs])
        c_isf = (C.c_int * n)(*[1 if a.dtype == np.float32 else 0 for a in arrs])
        h = C.c_void_p()
        _check(l.bt_model_create(n, c_names, c_data, c_counts, c_isf, int(device), C.byref(h)))
        self._h = h
        self.device = int(device)
        self.tables = tables
        d = (C.c_int * 9)()
        _check(l.bt_model_dims(h, d))
        self.nq, self.nv, self.nu, self.na, self.nbody, self.obs_size, self.smem_floats, self.ncon, self.n_clips = [int(x) for x in d]
        g = (C.c_int * 3)()
        _check(l.bt_model_launch(h, g))
        self.warps_per_cta, self.max_ctas, self.smem_bytes = int(g[0]), int(g[1]), int(g[2])

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and _lib is not None:
            _lib.bt_model_destroy(h)
            self._h = None

    def offset(self, region: str) -> int:
        return int(self.tables["o_" + region][0])

    # ---- torch helpers -------------------------------------------------------------------------
    def _torch(self):
        import torch
        return torch

    def _dev(self):
        return self._torch().device("cuda", self.device)

    def new_state(self, n: int):
        torch = self._torch()
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self._dev())
        return dict(qpos=z(n, self.nq), qvel=z(n, self.nv), act=z(n, self.na), qacc_warmstart=z(n, self.nv), time=z(n),
                    xpos=z(n, 3 * self.nbody))

    def new_outputs(self, n: int):
        torch = self._torch()
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self._dev())
        return dict(obs=z(n, self.obs_size), reward=z(n), done=z(n), metrics=z(n, NUM_METRICS), info_f=z(n, NUM_INFO_F),
                    info_i=torch.zeros(n, NUM_INFO_I, dtype=torch.int32, device=self._dev()))

    def _sp(self, st, n) -> StatePtrs:
        torch = self._torch()
        shapes = dict(qpos=(n, self.nq), qvel=(n, self.nv), act=(n, self.na), qacc_warmstart=(n, self.nv), time=(n,),
                      xpos=(n, 3 * self.nbody))
        s = StatePtrs()
        for k in STATE_FIELDS:
            t = st[k]
            if t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape) != shapes[k] or t.device != self._dev():
                raise ValueError(f"state['{k}'] must be a contiguous float32 {shapes[k]} tensor on cuda:{self.device}")
            setattr(s, k, t.data_ptr() if t.numel() else None)
        return s

    def _p(self, t, shape, dtype=None):
        torch = self._torch()
        dtype = dtype or torch.float32
        # page-locked host tensors are accepted too: under unified addressing the kernel writes them through the same pointer
        # (zero-copy device->host, used for the observation rows of host-side consumers)
        on_dev = t.device == self._dev() or (t.device.type == "cpu" and t.is_pinned())
        if t.dtype != dtype or not t.is_contiguous() or tuple(t.shape) != tuple(shape) or not on_dev:
            raise ValueError(f"expected a contiguous {dtype} tensor of shape {tuple(shape)} on cuda:{self.device} (or pinned host memory), "
                             f"got {t.dtype} {tuple(t.shape)} on {t.device}")
        return C.c_void_p(t.data_ptr())

    def _stream(self):
        return C.c_void_p(self._torch().cuda.current_stream(self.device).cuda_stream)

    # ---- entry points ----------------------------------------------------------------------------
    def _clip(self, clip_idx, n):
        """[n] int32 clip indices (RodentMultiClip) or None: a NULL pointer, the single-clip behaviour"""
        return None if clip_idx is None else self._p(clip_idx, (n,), self._torch().int32)

    def reset(self, keys, state, out, fixed_start_frame: int = -1, clip_idx=None):
        torch = self._torch()
        n = keys.shape[0]
        _check(lib().bt_reset(self._h, n, self._p(keys, (n, 2), torch.uint32 if keys.dtype == torch.uint32 else torch.int32),
                              int(fixed_start_frame), self._sp(state, n), self._p(out["obs"], (n, self.obs_size)), self._p(out["reward"], (n,)),
                              self._p(out["done"], (n,)), self._p(out["metrics"], (n, NUM_METRICS)),
                              self._p(out["info_f"], (n, NUM_INFO_F)), self._p(out["info_i"], (n, NUM_INFO_I), torch.int32),
                              self._clip(clip_idx, n), self._stream()))

    def step(self, action, state, first, first_obs, first_info_i, out, clip_idx=None):
        torch = self._torch()
        n = action.shape[0]
        _check(lib().bt_step(self._h, n, self._p(action, (n, self.nu)), self._sp(state, n), self._sp(first, n),
                             self._p(first_obs, (n, self.obs_size)), self._p(first_info_i, (n, NUM_INFO_I), torch.int32),
                             self._p(out["obs"], (n, self.obs_size)), self._p(out["reward"], (n,)), self._p(out["done"], (n,)),
                             self._p(out["metrics"], (n, NUM_METRICS)), self._p(out["info_f"], (n, NUM_INFO_F)),
                             self._p(out["info_i"], (n, NUM_INFO_I), torch.int32), self._clip(clip_idx, n), self._stream()))

    def physics_step(self, ctrl, state, n_substeps: int):
        n = state["qpos"].shape[0]
        _check(lib().bt_physics_step(self._h, n, self._p(ctrl, (n, self.nu)), self._sp(state, n), int(n_substeps), self._stream()))

    def pipeline_init(self, state):
        n = state["qpos"].shape[0]
        _check(lib().bt_pipeline_init(self._h, n, self._sp(state, n), self._stream()))

    def reward_obs(self, action, state, out, clip_idx=None):
        torch = self._torch()
        n = action.shape[0]
        _check(lib().bt_reward_obs(self._h, n, self._p(action, (n, self.nu)), self._sp(state, n),
                                   self._p(out["info_i"], (n, NUM_INFO_I), torch.int32), self._p(out["obs"], (n, self.obs_size)),
                                   self._p(out["reward"], (n,)), self._p(out["done"], (n,)), self._p(out["metrics"], (n, NUM_METRICS)),
                                   self._p(out["info_f"], (n, NUM_INFO_F)), self._clip(clip_idx, n), self._stream()))

    def kinematics(self, qpos):
        """Batched forward kinematics (mjx smooth.kinematics; preprocessing/preprocess.py:144-204): qpos [n, nq] ->
        (xpos [n, nbody, 3], xquat [n, nbody, 4]) through the tree pass of the step kernel."""
        n = qpos.shape[0]
        st = self.new_state(n)
        st["qpos"].copy_(qpos)
        scratch, _, _ = self.forward_debug(None, st, stop=1)
        ox, oq = self.offset("xpos"), self.offset("xquat")
        return (scratch[:, ox:ox + 3 * self.nbody].reshape(n, self.nbody, 3).clone(),
                scratch[:, oq:oq + 4 * self.nbody].reshape(n, self.nbody, 4).clone())

    def forward_debug(self, ctrl: Optional["object"], state, stop: int = 0):
        torch = self._torch()
        n = state["qpos"].shape[0]
        scratch = torch.zeros(n, self.smem_floats, dtype=torch.float32, device=self._dev())
        cdist = torch.zeros(n, max(self.ncon, 1), dtype=torch.float32, device=self._dev())
        niter = torch.zeros(n, dtype=torch.int32, device=self._dev())
        cp = self._p(ctrl, (n, self.nu)) if ctrl is not None else None
        _check(lib().bt_forward_debug(self._h, n, cp, self._sp(state, n), int(stop), C.c_void_p(scratch.data_ptr()),
                                      C.c_void_p(cdist.data_ptr()), C.c_void_p(niter.data_ptr()), self._stream()))
        return scratch, cdist, niter
