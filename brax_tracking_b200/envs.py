"""Host-side mirror of the reference's environment interface for the tracking hot path.

Reference interface (paths relative to the reference checkout):
  * ``PipelineEnv.reset(rng) -> State`` / ``step(State, action) -> State`` as implemented by
    ``Fruitfly_Tethered`` / ``Fruitfly_Tethered_Free`` (envs/fruitfly.py:20-341, :346-668) and intended by
    ``RodentSingleClip`` (envs/rodent.py:19-353; canonical definition in SURVEY.md Appendix B.3),
  * ``custom_wrappers.wrap`` -> ``AutoResetWrapperTracking(VmapWrapper(EpisodeWrapper(env)))``
    (custom_brax/custom_wrappers.py:14-80),
  * the attributes other layers read: ``sys``, ``dt``, ``action_size``, ``observation_size``,
    ``_steps_for_cur_frame``, ``_thorax_idx``, ``_free_jnt``, ``_reset_noise_scale`` (main.py:86,147,243,275,283).

Differences that come with the B200 design: environments are always batched (the ``VmapWrapper``
is built in: ``rng`` is ``[n, 2]`` uint32 JAX keys, ``action`` is ``[n, nu]``), tensors are ``torch`` CUDA tensors,
and ``step`` updates the state buffers in place and returns a ``State`` viewing them (the reference returns fresh
arrays; XLA donates them).  ``wrap(env).step`` is ONE kernel launch: physics x n_frames + reward + obs +
episode bookkeeping + auto-reset selection.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, Optional

import numpy as np

from . import assets, configs, mjcf, model as model_mod, native
from .clips import ReferenceClip


@dataclass
class State:
    """brax.envs.base.State (pipeline_state, obs, reward, done, metrics, info) over batched device buffers."""
    pipeline_state: Dict[str, Any]
    obs: Any
    reward: Any
    done: Any
    metrics: Dict[str, Any] = field(default_factory=dict)
    info: Dict[str, Any] = field(default_factory=dict)
    _raw: Dict[str, Any] = field(default_factory=dict, repr=False)


def _views(out) -> tuple:
    metrics = {k: out["metrics"][:, i] for i, k in enumerate(native.METRIC_NAMES)}
    info = {k: out["info_f"][:, i] for i, k in enumerate(native.INFO_F_NAMES)}
    info.update({k: out["info_i"][:, i] for i, k in enumerate(native.INFO_I_NAMES)})
    if "clip_idx" in out:
        info["clip_idx"] = out["clip_idx"]
    return metrics, info


class TrackingEnv:
    """Single-clip imitation env on the fused sm_100a step (batched; one process drives one GPU)."""

    def __init__(self, mj_model: mjcf.Model, reference_clip, env_args: dict, device: int = 0):
        """``reference_clip``: a ``ReferenceClip`` / dict of its arrays, or a list of them (multi-clip, ``RodentMultiClip``): stacked
        on a leading clip axis as ``preprocess.save_reference_clip`` stores them (preprocessing/preprocess.py:254-258)."""
        if isinstance(reference_clip, (list, tuple)):
            cs = [c.as_dict() if isinstance(c, ReferenceClip) else dict(c) for c in reference_clip]
            clip = {k: np.stack([np.asarray(c[k]) for c in cs]) for k in cs[0]}
        else:
            clip = reference_clip.as_dict() if isinstance(reference_clip, ReferenceClip) else dict(reference_clip)
        self.sys = mj_model
        self._cfg = configs.resolve(mj_model, env_args)
        self._tables = model_mod.pack(mj_model, self._cfg, clip)
        self._native = native.NativeModel(self._tables, device)
        self._ref_traj = clip
        c = self._cfg
        self._steps_for_cur_frame = c["steps_for_cur_frame"]
        self._thorax_idx = c["torso_idx"]
        self._joint_idxs, self._body_idxs, self._endeff_idxs = c["joint_idxs"], c["body_idxs"], c["endeff_idxs"]
        self._free_jnt = c["free_jnt"]
        self._reset_noise_scale = c["reset_noise_scale"]
        self._ref_len = c["ref_len"]
        self._n_frames = c["n_frames"]
        self.episode_length = c["episode_length"]

    # ---- brax Env surface ----
    @property
    def dt(self) -> float:
        return self.sys.timestep * self._n_frames

    @property
    def action_size(self) -> int:
        return self.sys.nu

    @property
    def observation_size(self) -> int:
        return self._native.obs_size

    @property
    def backend(self) -> str:
        return "b200"

    def _keys(self, rng):
        import torch
        if isinstance(rng, np.ndarray):
            rng = torch.from_numpy(np.ascontiguousarray(rng.astype(np.uint32)).view(np.int32))
        if rng.dtype == torch.uint32:
            rng = rng.view(torch.int32)
        return rng.to(device=self._native._dev(), dtype=torch.int32).contiguous()

    def _new_clip_idx(self, out, n, clip_idx=None):
        """multi-clip models carry ``info['clip_idx']`` ([n] int32; drawn by the reset kernel, or given for render rollouts)"""
        if self._native.n_clips <= 1:
            return None
        import torch
        c = torch.zeros(n, dtype=torch.int32, device=self._native._dev())
        if clip_idx is not None:
            c.copy_(torch.as_tensor(clip_idx, dtype=torch.int32))
        out["clip_idx"] = c
        return c

    def reset(self, rng) -> State:
        """Fruitfly_Tethered_Free.reset (fruitfly.py:449-495) for a batch of JAX keys ``[n, 2]``."""
        keys = self._keys(rng)
        n = keys.shape[0]
        st, out = self._native.new_state(n), self._native.new_outputs(n)
        self._native.reset(keys, st, out, clip_idx=self._new_clip_idx(out, n))
        metrics, info = _views(out)
        info.pop("steps"); info.pop("truncation")  # those belong to the EpisodeWrapper
        return State(st, out["obs"], out["reward"], out["done"], metrics, info, _raw=out)

    def step(self, state: State, action) -> State:
        """Fruitfly_Tethered_Free.step (fruitfly.py:497-596), unwrapped: pipeline_step + reward/obs, in place."""
        self._native.physics_step(action, state.pipeline_state, self._n_frames)
        self._native.reward_obs(action, state.pipeline_state, state._raw, clip_idx=state._raw.get("clip_idx"))
        return state

    def pipeline_init(self, qpos, qvel) -> Dict[str, Any]:
        """PipelineEnv.pipeline_init = mjx.forward on (qpos, qvel) with zero act/ctrl (fruitfly.py:477)."""
        n = qpos.shape[0]
        st = self._native.new_state(n)
        st["qpos"].copy_(qpos); st["qvel"].copy_(qvel)
        self._native.pipeline_init(st)
        return st

    def pipeline_step(self, pipeline_state, action) -> Dict[str, Any]:
        """PipelineEnv.pipeline_step = mjx.step x n_frames (fruitfly.py:500), in place."""
        self._native.physics_step(action, pipeline_state, self._n_frames)
        return pipeline_state


def _load(name, mj_model):
    return mj_model if mj_model is not None else assets.load_model(name)


def RodentSingleClip(reference_clip, mj_model: Optional[mjcf.Model] = None, device: int = 0, **overrides) -> TrackingEnv:
    """envs/rodent.py:19-136 with the canonical fixes of SURVEY.md Appendix B.3."""
    return TrackingEnv(_load("rodent", mj_model), reference_clip, dict(configs.RODENT_ENV_ARGS, **overrides), device)


def RodentMultiClip(reference_clips, mj_model: Optional[mjcf.Model] = None, device: int = 0, **overrides) -> TrackingEnv:
    """envs/rodent.py:377 -- an EMPTY class body in the reference (a SyntaxError).  Defined here as RodentSingleClip
    over several clips stacked on a leading axis (preprocessing/preprocess.py:254-258): every environment draws its clip at reset
    (``randint(rng_pos, (), 0, n_clips)``, the key the single-clip reset splits off and never uses, fruitfly.py:451), keeps it in
    ``info['clip_idx']`` and through auto-resets, and every clip gather of the step is offset by it."""
    return TrackingEnv(_load("rodent", mj_model), list(reference_clips), dict(configs.RODENT_ENV_ARGS, **overrides), device)


def Fruitfly_Tethered_Free(reference_clip, mj_model: Optional[mjcf.Model] = None, device: int = 0, **overrides) -> TrackingEnv:
    """envs/fruitfly.py:343-447."""
    return TrackingEnv(_load("fly_free", mj_model), reference_clip, dict(configs.FLY_FREEJNT_ENV_ARGS, **overrides), device)


def Fruitfly_Tethered(reference_clip, mj_model: Optional[mjcf.Model] = None, device: int = 0, **overrides) -> TrackingEnv:
    """envs/fruitfly.py:17-120."""
    return TrackingEnv(_load("fly_tethered", mj_model), reference_clip, dict(configs.FLY_ENV_ARGS, **overrides), device)


class AutoResetWrapperTracking:
    """custom_brax/custom_wrappers.py:43-80 fused with brax's EpisodeWrapper / VmapWrapper: one launch per step."""

    def __init__(self, env: TrackingEnv, episode_length: Optional[int] = None, action_repeat: int = 1):
        if action_repeat != 1:
            raise NotImplementedError("action_repeat != 1 is not used by the reference configs (train_fly.yaml:18)")
        if episode_length is not None and episode_length != env.episode_length:
            # the episode length is a model constant of the fused kernel: rebuild the tables with the new value
            env._cfg["episode_length"] = int(episode_length)
            env._tables = model_mod.pack(env.sys, env._cfg, env._ref_traj)
            env._native = native.NativeModel(env._tables, env._native.device)
            env.episode_length = int(episode_length)
        self.env = env

    def __getattr__(self, name):  # brax Wrapper.__getattr__ forwarding
        return getattr(self.env, name)

    def reset(self, rng) -> State:
        env = self.env
        keys = env._keys(rng)
        n = keys.shape[0]
        st, out = env._native.new_state(n), env._native.new_outputs(n)
        env._native.reset(keys, st, out, clip_idx=env._new_clip_idx(out, n))
        metrics, info = _views(out)
        # custom_wrappers.py:46-52: cache the first state / obs / frame counters for the auto-reset selection
        first = {k: v.clone() for k, v in st.items()}
        info["first_pipeline_state"] = first
        info["first_obs"] = out["obs"].clone()
        first_info_i = out["info_i"].clone()
        info["first_cur_frame"] = first_info_i[:, 0]
        info["first_steps_taken_cur_frame"] = first_info_i[:, 1]
        raw = dict(out, first=first, first_obs=info["first_obs"], first_info_i=first_info_i)
        return State(st, out["obs"], out["reward"], out["done"], metrics, info, _raw=raw)

    def step(self, state: State, action) -> State:
        r = state._raw
        self.env._native.step(action, state.pipeline_state, r["first"], r["first_obs"], r["first_info_i"], r, clip_idx=r.get("clip_idx"))
        return state

    def bind_host_obs(self, state: State):
        """For host-side consumers of the observations (a CPU policy, logging): from now on ``state.obs`` lives in page-locked
        host memory and the step kernel writes each environment's row there directly (mapped pointer, zero-copy), spread
        over the launch instead of a device->host copy after it.  Valid after the stream is synchronised.  Returns the tensor."""
        import torch
        h = torch.empty(state.obs.shape, dtype=torch.float32).pin_memory()
        h.copy_(state.obs)
        state._raw["obs"] = h
        state.obs = h
        return h


class RenderRolloutWrapperTracking:
    """custom_brax/custom_wrappers.py:82-125: "always resets to 0" -- deterministic start frame for evaluation / rendering
    rollouts (main.py:131-151); steps are the bare env's (no episode wrapper, no auto-reset)."""

    def __init__(self, env: TrackingEnv):
        self.env = env

    def __getattr__(self, name):
        return getattr(self.env, name)

    def reset(self, rng, clip_idx=None) -> State:
        """``clip_idx`` ([n] ints, multi-clip models only): the clip every environment replays (default: clip 0)."""
        env = self.env
        keys = env._keys(rng)
        n = keys.shape[0]
        st, out = env._native.new_state(n), env._native.new_outputs(n)
        env._native.reset(keys, st, out, fixed_start_frame=0, clip_idx=env._new_clip_idx(out, n, clip_idx))
        metrics, info = _views(out)
        info.pop("steps"); info.pop("truncation")
        return State(st, out["obs"], out["reward"], out["done"], metrics, info, _raw=out)

    def step(self, state: State, action) -> State:
        return self.env.step(state, action)


def wrap(env: TrackingEnv, episode_length: int = 1000, action_repeat: int = 1, randomization_fn=None) -> AutoResetWrapperTracking:
    """custom_brax/custom_wrappers.py:14-40."""
    if randomization_fn is not None:
        raise NotImplementedError("domain randomisation is not used by the reference configs")
    return AutoResetWrapperTracking(env, episode_length, action_repeat)
