"""Numpy-in / numpy-out adapters over the two builds of the same per-environment programs (TEST INFRASTRUCTURE):
``CudaBackend`` = the product (libbt_b200.so through the C ABI, device tensors via torch),
``EmuBackend``  = tests/host_emu (one lane per environment, CPU) for the `-m "not gpu"` suite."""
from __future__ import annotations

import numpy as np

STATE = ("qpos", "qvel", "act", "qacc_warmstart", "time", "xpos")
OUT = ("obs", "reward", "done", "metrics", "info_f", "info_i")


def state_from_oracle(ps, N):
    """oracle pipeline_state dict -> fresh float32 [N, dim] arrays"""
    f32 = lambda x: np.array(x, dtype=np.float32, copy=True)
    return {k: (f32(ps[k]).reshape(N, -1) if k != "time" else f32(ps[k])) for k in STATE}


class EmuBackend:
    name = "emu"

    def __init__(self, tables):
        import emu
        self.e = emu.Emu(tables)
        self.t = tables

    def new_outputs(self, N):
        return self.e.new_outputs(N)

    def reset(self, keys, fixed_start_frame=-1, clip_idx=None):
        return self.e.reset(keys, fixed_start_frame, clip_idx)

    def step(self, st, out, first, first_obs, first_ii, action, clip_idx=None):
        self.e.step(st, out, first, first_obs, first_ii, action, clip_idx)
        return st, out

    def physics_step(self, st, ctrl, n):
        self.e.physics_step(st, ctrl, n)
        return st

    def reward_obs(self, st, out, action, clip_idx=None):
        self.e.reward_obs(st, out, action, clip_idx)
        return out

    def pipeline_init(self, st):
        self.e.physics_step(st, None, 0)
        return st

    def forward_debug(self, st, ctrl, stop=0):
        return self.e.forward_debug(st, ctrl, stop)


class CudaBackend:
    name = "cuda"

    def __init__(self, tables, device=0):
        import torch
        from brax_tracking_b200 import native
        self.torch = torch
        self.nm = native.NativeModel(tables, device)
        self.t = tables

    def _d(self, a):
        t = self.torch.from_numpy(np.ascontiguousarray(a))
        return t.cuda(self.nm.device)

    def _dst(self, st):
        return {k: self._d(st[k]) for k in STATE}

    def _back(self, dev, host):
        for k, v in dev.items():
            if k in host:
                host[k][...] = v.cpu().numpy().reshape(host[k].shape)

    def new_outputs(self, N):
        return {k: v.cpu().numpy() for k, v in self.nm.new_outputs(N).items()}

    def reset(self, keys, fixed_start_frame=-1, clip_idx=None):
        N = keys.shape[0]
        st, out = self.nm.new_state(N), self.nm.new_outputs(N)
        dc = None if clip_idx is None else self._d(clip_idx)
        self.nm.reset(self._d(np.ascontiguousarray(keys, dtype=np.uint32).view(np.int32)), st, out, fixed_start_frame, clip_idx=dc)
        if dc is not None:
            clip_idx[...] = dc.cpu().numpy()
        return {k: v.cpu().numpy() for k, v in st.items()}, {k: v.cpu().numpy() for k, v in out.items()}

    def step(self, st, out, first, first_obs, first_ii, action, clip_idx=None):
        dst, dout = self._dst(st), {k: self._d(out[k]) for k in OUT}
        self.nm.step(self._d(np.asarray(action, np.float32)), dst, self._dst(first), self._d(first_obs), self._d(first_ii), dout,
                     clip_idx=None if clip_idx is None else self._d(clip_idx))
        self._back(dst, st); self._back(dout, out)
        return st, out

    def physics_step(self, st, ctrl, n):
        dst = self._dst(st)
        self.nm.physics_step(self._d(np.asarray(ctrl, np.float32)), dst, n)
        self._back(dst, st)
        return st

    def reward_obs(self, st, out, action, clip_idx=None):
        dst, dout = self._dst(st), {k: self._d(out[k]) for k in OUT}
        self.nm.reward_obs(self._d(np.asarray(action, np.float32)), dst, dout, clip_idx=None if clip_idx is None else self._d(clip_idx))
        self._back(dout, out)
        return out

    def pipeline_init(self, st):
        dst = self._dst(st)
        self.nm.pipeline_init(dst)
        self._back(dst, st)
        return st

    def forward_debug(self, st, ctrl, stop=0):
        sc, cd, ni = self.nm.forward_debug(None if ctrl is None else self._d(np.asarray(ctrl, np.float32)), self._dst(st), stop)
        return sc.cpu().numpy(), cd.cpu().numpy(), ni.cpu().numpy()
