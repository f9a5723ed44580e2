"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product package).

NumPy restatement of the *environment layer* of the reference hot path, batched over envs:

  * ``Fruitfly_Tethered_Free.reset / step / _get_obs / _bounded_quat_dist``
    (/root/reference/envs/fruitfly.py:449-668) -- the free-root template that also defines the
    canonical rodent env (rodent root seeding rodent.py:154-159),
  * ``Fruitfly_Tethered`` differences (fruitfly.py:122-341),
  * ``AutoResetWrapperTracking`` (/root/reference/custom_brax/custom_wrappers.py:44-80),
  * brax ``EpisodeWrapper`` (third-party),
  * JAX threefry2x32 ``split / randint / uniform`` (third-party).

Physics (``pipeline_init`` / ``pipeline_step``) is delegated to ``oracle.Oracle`` (mjx_oracle.c).
**parity unpinned** (SURVEY.md section 8c): no reference golden vectors exist; the threefry core is pinned
by the published known-answer vectors in tests/test_prng.py.
"""
from __future__ import annotations

import numpy as np

# ------------------------------------------------------------------------------------------------
# JAX PRNG (legacy threefry2x32, non-partitionable) -- SURVEY.md Appendix D
# ------------------------------------------------------------------------------------------------
_U32 = np.uint32


def _rotl(x, d):
    return ((x << _U32(d)) | (x >> _U32(32 - d))).astype(_U32)


def threefry2x32(key, x0, x1):
    """key: (k0,k1) uint32 scalars or arrays broadcastable with x0/x1."""
    with np.errstate(over="ignore"):
        k0, k1 = _U32(key[0]) if np.isscalar(key[0]) else key[0].astype(_U32), _U32(key[1]) if np.isscalar(key[1]) else key[1].astype(_U32)
        ks = [k0, k1, (k0 ^ k1 ^ _U32(0x1BD11BDA)).astype(_U32) if not np.isscalar(k0) else _U32(k0 ^ k1 ^ _U32(0x1BD11BDA))]
        x0 = (np.asarray(x0, dtype=_U32) + ks[0]).astype(_U32)
        x1 = (np.asarray(x1, dtype=_U32) + ks[1]).astype(_U32)
        rots = [[13, 15, 26, 6], [17, 29, 16, 24]]
        for g in range(5):
            for r in rots[g % 2]:
                x0 = (x0 + x1).astype(_U32)
                x1 = _rotl(x1, r)
                x1 = (x1 ^ x0).astype(_U32)
            x0 = (x0 + ks[(g + 1) % 3]).astype(_U32)
            x1 = (x1 + ks[(g + 2) % 3] + _U32(g + 1)).astype(_U32)
    return x0, x1


def _threefry_counts(key, n):
    """threefry_2x32(key, iota(n)) with JAX's split-in-halves convention; an odd-length counter array is padded with a
    ZERO (jax/_src/prng.py::threefry_2x32: ``concatenate([count.ravel(), np.uint32([0])])``), not with n."""
    cnt = np.arange(n + (n % 2), dtype=_U32)
    if n % 2:
        cnt[-1] = 0
    half = len(cnt) // 2
    y0, y1 = threefry2x32(key, cnt[:half], cnt[half:])
    return np.concatenate([y0, y1])[:n]


def split(key, num=2):
    """jax.random.split for one key (k0,k1) -> [num,2] uint32."""
    return _threefry_counts(key, 2 * num).reshape(num, 2)


def random_bits(key, n):
    return _threefry_counts(key, n)


def uniform(key, n, minval, maxval):
    bits = random_bits(key, n)
    f = ((bits >> _U32(9)) | _U32(0x3F800000)).view(np.float32) - np.float32(1.0)
    lo, hi = np.float32(minval), np.float32(maxval)
    return np.maximum(lo, f * (hi - lo) + lo).astype(np.float32)


def randint(key, minval, maxval):
    k = split(key, 2)
    hi_bits = int(random_bits((k[0, 0], k[0, 1]), 1)[0])
    lo_bits = int(random_bits((k[1, 0], k[1, 1]), 1)[0])
    span = int(maxval - minval)
    if span <= 0:
        span = 1
    mult = (2 ** 16) % span
    mult = (mult * mult) % span
    off = ((hi_bits % span) * mult + (lo_bits % span)) & 0xFFFFFFFF
    return int(minval + off % span)


# ------------------------------------------------------------------------------------------------
# brax.math helpers (third-party), batched on leading axes
# ------------------------------------------------------------------------------------------------
def rotate(v, q):
    s = q[..., 0:1]
    u = q[..., 1:]
    return 2 * np.sum(u * v, -1, keepdims=True) * u + (s * s - np.sum(u * u, -1, keepdims=True)) * v + 2 * s * np.cross(u, v)


def quat_mul(a, b):
    return np.stack([
        a[..., 0] * b[..., 0] - a[..., 1] * b[..., 1] - a[..., 2] * b[..., 2] - a[..., 3] * b[..., 3],
        a[..., 0] * b[..., 1] + a[..., 1] * b[..., 0] + a[..., 2] * b[..., 3] - a[..., 3] * b[..., 2],
        a[..., 0] * b[..., 2] - a[..., 1] * b[..., 3] + a[..., 2] * b[..., 0] + a[..., 3] * b[..., 1],
        a[..., 0] * b[..., 3] + a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1] + a[..., 3] * b[..., 0],
    ], axis=-1)


def quat_inv(q):
    return q * np.array([1, -1, -1, -1], dtype=q.dtype)


def relative_quat(q1, q2):
    return quat_mul(q2, quat_inv(q1))


def _gather_idx(idx, n):
    """JAX gather semantics for x[..., idx]: negative wraps once, then clamp."""
    idx = np.asarray(idx, dtype=np.int64)
    idx = np.where(idx < 0, idx + n, idx)
    return np.clip(idx, 0, n - 1)


METRIC_NAMES = ["pos_reward", "quat_reward", "joint_reward", "angvel_reward", "bodypos_reward", "endeff_reward",
                "reward_quadctrl", "reward_alive", "too_far", "bad_pose", "bad_quat", "fall"]


class EnvOracle:
    """Batched restatement of wrap(Env) = AutoResetWrapperTracking(VmapWrapper(EpisodeWrapper(Env)))."""

    def __init__(self, physics, clip, cfg, dtype=np.float32):
        """physics: oracle.Oracle; clip: dict of arrays (position, quaternion, joints, body_positions,
        angular_velocity ...), time-major; cfg: dict with the env constants (see brax_tracking_b200.envs.EnvConfig)."""
        self.o = physics
        self.m = physics.m
        self.dt = np.dtype(dtype)
        # several clips stacked on a leading axis (preprocess.py:254-258) are kept as ONE table of n_clips * T rows; an environment
        # with clip index c reads row c * T + frame (the frame clamps stay per clip)
        multi = np.asarray(clip["joints"]).ndim == 3
        self.n_clips = int(np.asarray(clip["joints"]).shape[0]) if multi else 1
        self.T = int(np.asarray(clip["joints"]).shape[1 if multi else 0])
        self.clip = {k: np.asarray(v, dtype=self.dt).reshape((self.n_clips * self.T,) + np.asarray(v).shape[(2 if multi else 1):])
                     for k, v in clip.items()}
        self.c = dict(cfg)

    def animals(self):
        """Per-animal constants (brax_tracking_b200.configs.resolve): every reference env has one animal; the two-rodent
        model of BASELINE.json configs[3] applies the single-animal expressions to each animal's own slices."""
        c = self.c
        nj = self.clip["joints"].shape[1]
        return c.get("animals") or [dict(qadr=0, dadr=0, nj=nj, jbase=0, torso_idx=c["torso_idx"], joint_idxs=c["joint_idxs"],
                                         body_idxs=c["body_idxs"], endeff_idxs=c["endeff_idxs"])]

    # -- fruitfly.py:598-646 ---------------------------------------------------------------------
    def get_obs(self, qpos, qvel, xpos, cur_frame, clip_idx=None):
        c, dt = self.c, self.dt
        N = qpos.shape[0]
        L = c["ref_len"]
        start = np.clip(cur_frame + 1, 0, self.T - L)  # dynamic_slice clamps the start
        if clip_idx is not None:
            start = start + np.asarray(clip_idx) * self.T
        win = start[:, None] + np.arange(L)[None, :]   # [N, L]
        parts = [qpos, qvel]
        free = c["free_jnt"]
        ans = self.animals()
        NA = len(ans)
        cpos = self.clip["position"].reshape(-1, NA, 3)
        cquat = self.clip["quaternion"].reshape(-1, NA, 4)
        for a, an in enumerate(ans):
            qa = an["qadr"]
            # tethered: fruitfly.py:271-319 -- no pos/quat terms, full qpos joints, offsets rotated by qpos[3:7] (joint angles!)
            quat = qpos[:, qa + 3:qa + 7]
            if free:
                parts.append(rotate(cpos[win, a] - qpos[:, None, qa:qa + 3], quat[:, None, :]).reshape(N, -1))
                parts.append(relative_quat(quat[:, None, :], cquat[win, a]).reshape(N, -1))
                q0 = qa + 7
            else:
                q0 = qa
            jd = self.clip["joints"][win][:, :, an["jbase"]:an["jbase"] + an["nj"]] - qpos[:, None, q0:q0 + an["nj"]]
            parts.append(jd[:, :, _gather_idx(an["joint_idxs"], an["nj"])].reshape(N, -1))
            bidx = _gather_idx(an["body_idxs"], self.m.nbody)
            bd = (self.clip["body_positions"][win] - xpos[:, None, :, :])[:, :, bidx]  # [N,L,nb,3]
            parts.append(rotate(bd, quat[:, None, None, :]).reshape(N, -1))
        return np.concatenate(parts, axis=1).astype(dt)

    @staticmethod
    def bounded_quat_dist(source, target):
        source = source / np.linalg.norm(source, axis=-1, keepdims=True)
        target = target / np.linalg.norm(target, axis=-1, keepdims=True)
        dist = 2 * np.sum(source * target, -1) ** 2 - 1
        dist = np.minimum(dist.dtype.type(1.0), dist)
        return 0.5 * np.arccos(dist)

    # -- reset: fruitfly.py:449-495 (+ rodent.py:154-159), EpisodeWrapper.reset, AutoReset.reset -----
    def reset(self, keys, fixed_start_frame=-1, start_frames=None, clip_idx=None):
        """keys: [N,2] uint32 (one JAX key per env, as jax.random.split(key_env, num_envs)).
        fixed_start_frame >= 0: RenderRolloutWrapperTracking.reset (custom_wrappers.py:85-125): split(rng, 3), that frame.
        start_frames (test-only): per-env start frames replacing the randint draw (late-clip cases)."""
        c, dt, m = self.c, self.dt, self.m
        N = keys.shape[0]
        qpos = np.zeros((N, m.nq), dtype=dt)
        qvel = np.zeros((N, m.nv), dtype=dt)
        start = np.zeros(N, dtype=np.int32)
        clip_sel = np.zeros(N, dtype=np.int32) if clip_idx is None else np.asarray(clip_idx, np.int32).copy()
        lo, hi = -c["reset_noise_scale"], c["reset_noise_scale"]
        for e in range(N):
            k = split((keys[e, 0], keys[e, 1]), 4 if fixed_start_frame < 0 else 3)
            rng, rng1, rng2 = (k[0, 0], k[0, 1]), (k[1, 0], k[1, 1]), (k[2, 0], k[2, 1])
            start[e] = randint(rng, 0, 44) if fixed_start_frame < 0 else fixed_start_frame
            if start_frames is not None:
                start[e] = start_frames[e]
            if self.n_clips > 1 and fixed_start_frame < 0:
                # RodentMultiClip (an empty class in the reference): clip = randint(rng_pos, (), 0, n_clips), rng_pos = the fourth key
                clip_sel[e] = randint((k[3, 0], k[3, 1]), 0, self.n_clips)
            crow = int(clip_sel[e]) * self.T
            q0 = m.qpos0.astype(dt).copy()
            if c["seed_root_from_clip"] and fixed_start_frame < 0:
                fs = min(max(int(start[e]), 0), self.T - 1)   # JAX gather clamps
                ans = self.animals()
                for a, an in enumerate(ans):                  # every animal from its own copy of the clip
                    qa = an["qadr"]
                    q0[qa:qa + 2] = self.clip["position"].reshape(-1, len(ans), 3)[crow + fs, a, :2]
                    q0[qa + 3:qa + 7] = self.clip["quaternion"].reshape(-1, len(ans), 4)[crow + fs, a]
            qpos[e] = q0 + uniform(rng1, m.nq, lo, hi).astype(dt)
            qvel[e] = uniform(rng2, m.nv, lo, hi).astype(dt)
        st = dict(qpos=qpos, qvel=qvel, act=np.zeros((N, m.na), dtype=dt), qacc_warmstart=np.zeros((N, m.nv), dtype=dt),
                  time=np.zeros(N, dtype=dt))
        ps = self.o.pipeline_batch(st, None, 0, forward_only=True)  # pipeline_init = mjx.forward
        ps = {k: np.asarray(v, dtype=dt) for k, v in ps.items()}
        obs = self.get_obs(ps["qpos"], ps["qvel"], ps["xpos"], start, clip_sel)
        state = dict(
            pipeline_state=ps, obs=obs, reward=np.zeros(N, dtype=dt), done=np.zeros(N, dtype=dt),
            metrics={k: np.zeros(N, dtype=dt) for k in METRIC_NAMES},
            info=dict(cur_frame=start.copy(), steps_taken_cur_frame=np.zeros(N, dtype=np.int32), clip_idx=clip_sel,
                      summed_pos_distance=np.zeros(N, dtype=dt), quat_distance=np.zeros(N, dtype=dt),
                      joint_distance=np.zeros(N, dtype=dt),
                      steps=np.zeros(N, dtype=dt), truncation=np.zeros(N, dtype=dt)),
        )
        info = state["info"]
        info["first_pipeline_state"] = {k: v.copy() for k, v in ps.items()}
        info["first_obs"] = obs.copy()
        info["first_cur_frame"] = info["cur_frame"].copy()
        info["first_steps_taken_cur_frame"] = info["steps_taken_cur_frame"].copy()
        return state

    # -- env.step: fruitfly.py:497-596 -------------------------------------------------------------
    def env_step(self, state, action):
        c, dt, m = self.c, self.dt, self.m
        ps0 = state["pipeline_state"]
        ps = self.o.pipeline_batch(ps0, np.asarray(action, dtype=self.o.dtype), c["n_frames"])
        ps = {k: np.asarray(v, dtype=dt) for k, v in ps.items()}
        return self.reward_obs(state, ps, action)

    def reward_obs(self, state, ps, action):
        """Everything in env.step after pipeline_step (fruitfly.py:502-596)."""
        c, dt = self.c, self.dt
        f32 = dt.type
        info = dict(state["info"])
        N = ps["qpos"].shape[0]
        stc = info["steps_taken_cur_frame"] + 1
        hit = stc == c["steps_for_cur_frame"]
        cur = info["cur_frame"] + np.where(hit, 1, 0).astype(np.int32)
        stc = stc * np.where(hit, 0, 1).astype(np.int32)
        info["steps_taken_cur_frame"], info["cur_frame"] = stc, cur
        fi = np.clip(cur, 0, self.T - 1)  # JAX gather clamps
        cidx = info.get("clip_idx")
        if cidx is not None:
            fi = fi + np.asarray(cidx) * self.T
        qpos, qvel, xpos = ps["qpos"], ps["qvel"], ps["xpos"]
        free = c["free_jnt"]
        ans = self.animals()
        NA = len(ans)
        cpos = self.clip["position"].reshape(-1, NA, 3)
        cquat = self.clip["quaternion"].reshape(-1, NA, 4)
        cang = self.clip["angular_velocity"].reshape(-1, NA, 3)
        min_z, max_z = c["healthy_z_range"]
        z32 = lambda: np.zeros(N, dtype=dt)
        tot = {k: z32() for k in ("pos", "quat", "joint", "angvel", "bodypos", "endeff", "healthy")}
        flags = {k: z32() for k in ("too_far", "bad_pose", "bad_quat", "fall")}
        dist = {}
        with np.errstate(invalid="ignore"):
            for a, an in enumerate(ans):
                # ---- the single-animal expressions of fruitfly.py:514-552 on this animal's slices
                qa, da = an["qadr"], an["dadr"]
                if free:
                    pos_distance = qpos[:, qa:qa + 3] - cpos[fi, a]
                    pos_reward = f32(c["pos_reward_weight"]) * np.exp(f32(-400) * np.sum(pos_distance, -1) ** 2)
                    quat_distance = self.bounded_quat_dist(qpos[:, qa + 3:qa + 7], cquat[fi, a]) ** 2
                    quat_reward = f32(c["quat_reward_weight"]) * np.exp(f32(-4.0) * quat_distance)
                    q0 = qa + 7
                else:
                    pos_distance = np.zeros((N, 3), dtype=dt)
                    quat_distance = z32()
                    pos_reward = z32()
                    quat_reward = z32()
                    q0 = qa
                joint_distance = np.sum(qpos[:, q0:q0 + an["nj"]] - self.clip["joints"][fi][:, an["jbase"]:an["jbase"] + an["nj"]], -1) ** 2
                joint_reward = f32(c["joint_reward_weight"]) * np.exp(f32(-0.5) * joint_distance)
                angvel_reward = f32(c["angvel_reward_weight"]) * np.exp(f32(-0.5) * np.sum(qvel[:, da + 3:da + 6] - cang[fi, a], -1) ** 2)
                bidx = _gather_idx(an["body_idxs"], self.m.nbody)
                eidx = _gather_idx(an["endeff_idxs"], self.m.nbody)
                tb = self.clip["body_positions"][fi]
                bodypos_reward = f32(c["bodypos_reward_weight"]) * np.exp(f32(-6.0) * np.sum((xpos[:, bidx] - tb[:, bidx]).reshape(N, -1), -1) ** 2)
                endeff_reward = f32(c["endeff_reward_weight"]) * np.exp(f32(-0.75) * np.sum((xpos[:, eidx] - tb[:, eidx]).reshape(N, -1), -1) ** 2)
                z = xpos[:, an["torso_idx"], 2]
                is_healthy = np.where(z < f32(min_z), f32(0), f32(1))
                is_healthy = np.where(z > f32(max_z), f32(0), is_healthy)
                if c["terminate_when_unhealthy"]:
                    healthy_reward = np.full(N, c["healthy_reward"], dtype=dt)
                else:
                    healthy_reward = f32(c["healthy_reward"]) * is_healthy
                summed_pos_distance = np.sum((pos_distance * np.array([1.0, 1.0, 0.2], dtype=dt)) ** 2, -1)
                # ---- combination over animals (DESIGN.md "config 4"): reward terms add, flags and distances take the max
                for k, v in (("pos", pos_reward), ("quat", quat_reward), ("joint", joint_reward), ("angvel", angvel_reward),
                             ("bodypos", bodypos_reward), ("endeff", endeff_reward), ("healthy", healthy_reward)):
                    tot[k] = (tot[k] + v).astype(dt)
                for k, v in (("too_far", np.where(summed_pos_distance > f32(c["too_far_dist"]), f32(1), f32(0))),
                             ("bad_pose", np.where(joint_distance > f32(c["bad_pose_dist"]), f32(1), f32(0))),
                             ("bad_quat", np.where(quat_distance > f32(c["bad_quat_dist"]), f32(1), f32(0))),
                             ("fall", f32(1) - is_healthy)):
                    flags[k] = np.maximum(flags[k], v).astype(dt)
                for k, v in (("summed_pos_distance", summed_pos_distance), ("quat_distance", quat_distance), ("joint_distance", joint_distance)):
                    v = v.astype(dt)
                    dist[k] = v if a == 0 else np.where((v > dist[k]) | np.isnan(v), v, dist[k])   # max; NaN propagates
        pos_reward, quat_reward, joint_reward, angvel_reward = tot["pos"], tot["quat"], tot["joint"], tot["angvel"]
        bodypos_reward, endeff_reward, healthy_reward = tot["bodypos"], tot["endeff"], tot["healthy"]
        too_far, bad_pose, bad_quat, fall = flags["too_far"], flags["bad_pose"], flags["bad_quat"], flags["fall"]
        info["joint_distance"] = dist["joint_distance"]
        info["summed_pos_distance"] = dist["summed_pos_distance"]
        info["quat_distance"] = dist["quat_distance"]
        action = np.asarray(action, dtype=dt)
        ctrl_cost = f32(c["ctrl_cost_weight"]) * np.sum(np.square(action), -1)
        obs = self.get_obs(qpos, qvel, xpos, cur, cidx)
        reward = (joint_reward + pos_reward + quat_reward + angvel_reward + bodypos_reward + endeff_reward
                  + healthy_reward - ctrl_cost)
        done = fall if c["terminate_when_unhealthy"] else np.zeros(N, dtype=dt)
        done = np.max(np.stack([done, too_far, bad_pose, bad_quat]), axis=0)
        reward = np.nan_to_num(reward)
        obs = np.nan_to_num(obs)
        nan = np.zeros(N, dtype=bool)
        for k in ("qpos", "qvel", "act", "qacc_warmstart", "xpos", "time"):
            nan |= np.isnan(ps[k].reshape(N, -1)).any(axis=1)
        nan |= np.isnan(action.reshape(N, -1)).any(axis=1)   # data.ctrl = action is a leaf of the flattened mjx.Data (fruitfly.py:572)
        done = np.maximum(done, nan.astype(dt))
        metrics = dict(pos_reward=pos_reward, quat_reward=quat_reward, joint_reward=joint_reward,
                       angvel_reward=angvel_reward, bodypos_reward=bodypos_reward, endeff_reward=endeff_reward,
                       reward_quadctrl=-ctrl_cost, reward_alive=healthy_reward, too_far=too_far, bad_pose=bad_pose,
                       bad_quat=bad_quat, fall=fall)
        metrics = {k: np.asarray(v, dtype=dt) for k, v in metrics.items()}
        return dict(pipeline_state=ps, obs=obs.astype(dt), reward=reward.astype(dt), done=done.astype(dt),
                    metrics=metrics, info=info)

    # -- wrap(): EpisodeWrapper.step + AutoResetWrapperTracking.step ---------------------------------
    def step(self, state, action, physics_override=None):
        c, dt = self.c, self.dt
        info = dict(state["info"])
        # AutoResetWrapperTracking.step: custom_wrappers.py:54-59
        info["steps"] = np.where(state["done"] > 0, np.zeros_like(info["steps"]), info["steps"])
        st = dict(state)
        st["info"] = info
        st["done"] = np.zeros_like(state["done"])
        # EpisodeWrapper.step (action_repeat = 1)
        if physics_override is None:
            ns = self.env_step(st, action)
        else:
            ns = self.reward_obs(st, physics_override, action)
        steps = ns["info"]["steps"] + 1
        L = c["episode_length"]
        done = np.where(steps >= L, np.ones_like(ns["done"]), ns["done"])
        ns["info"]["truncation"] = np.where(steps >= L, 1 - ns["done"], np.zeros_like(ns["done"]))
        ns["info"]["steps"] = steps
        ns["done"] = done
        # AutoResetWrapperTracking.step: custom_wrappers.py:62-80
        d = done > 0
        ps = {}
        for k, v in ns["pipeline_state"].items():
            first = info["first_pipeline_state"][k]
            dd = d.reshape([-1] + [1] * (v.ndim - 1))
            ps[k] = np.where(dd, first, v)
        ns["pipeline_state"] = ps
        ns["obs"] = np.where(d[:, None], info["first_obs"], ns["obs"])
        ns["info"]["cur_frame"] = np.where(d, info["first_cur_frame"], ns["info"]["cur_frame"])
        ns["info"]["steps_taken_cur_frame"] = np.where(d, info["first_steps_taken_cur_frame"], ns["info"]["steps_taken_cur_frame"])
        for k in ("first_pipeline_state", "first_obs", "first_cur_frame", "first_steps_taken_cur_frame"):
            ns["info"][k] = info[k]
        return ns
