"""Backend-agnostic parity cases: the per-environment programs (CUDA through the C ABI, or the 1-lane host
emulation) against the oracle on identical seeded inputs.  Tolerances are stated here, next to the assertions.

Why the trajectory cases are *teacher forced*: the reference configuration (4 CG iterations, contact switching,
milligram bodies) is chaotic in fp32 -- the oracle's own float32 build departs from its float64 build by ~1e-6 after
one control step, ~1e-3 after two and O(1) after five (DESIGN.md "parity").  A free-running comparison after 10 or
100 control steps therefore measures the Lyapunov exponent, not the implementation.  Instead every control step of a
100-step trajectory is started from the ORACLE's state, the integer / flag outputs are required to be bit-exact, and
the one-step float error distribution is required to be no worse than the float32 oracle's own."""
from __future__ import annotations

import numpy as np

import common
import oracle as oracle_mod
from backends import state_from_oracle


def random_states(m, N, seed=0):
    rng = np.random.default_rng(seed)
    st = dict(qpos=np.zeros((N, m.nq), np.float32), qvel=np.zeros((N, m.nv), np.float32), act=np.zeros((N, m.na), np.float32),
              qacc_warmstart=np.zeros((N, m.nv), np.float32), time=np.zeros(N, np.float32), xpos=np.zeros((N, 3 * m.nbody), np.float32))
    roots = [int(m.jnt_qposadr[j]) for j in range(m.njnt) if m.jnt_type[j] == 0]
    hinge = np.ones(m.nq, bool)
    for qa in roots:
        hinge[qa:qa + 7] = False
    for e in range(N):
        q = m.qpos0.copy()
        q[hinge] += rng.uniform(-0.2, 0.2, int(hinge.sum()))
        for k, qa in enumerate(roots):
            q[qa + 2] += rng.uniform(-0.01, 0.02)
            quat = q[qa + 3:qa + 7] + rng.uniform(-0.2, 0.2, 4)
            q[qa + 3:qa + 7] = quat / np.linalg.norm(quat)
            if k > 0:     # further animals stand beside the first one (0.3 m apart: no inter-animal contact)
                q[qa:qa + 2] += [0.0, 0.3 * k]
        st["qpos"][e] = q
        st["qvel"][e] = rng.uniform(-1, 1, m.nv)
        st["act"][e] = rng.uniform(-0.5, 0.5, m.na)
        st["qacc_warmstart"][e] = rng.uniform(-5, 5, m.nv)
    ctrl = rng.uniform(-1, 1, (N, m.nu)).astype(np.float32)
    return st, ctrl


def touching_states(name, N, seed=0, min_active=1):
    """States of a two-animal model in which inter-animal contacts are ACTIVE (penetrating): the second root is placed beside the
    first by rejection sampling against the oracle's collision pass."""
    m, cfg, clip, tables = common.setup(name)
    o, _ = common.oracles(name)
    rng = np.random.default_rng(seed)
    roots = [int(m.jnt_qposadr[j]) for j in range(m.njnt) if m.jnt_type[j] == 0]
    assert len(roots) == 2
    cross = np.asarray(tables["con_xref"]) >= 0
    base, ctrl = random_states(m, 1, seed=seed + 100)
    out = {k: np.repeat(np.zeros_like(v), N, 0) for k, v in base.items()}
    ctrls = np.zeros((N, m.nu), np.float32)
    found = tries = 0
    while found < N:
        tries += 1
        assert tries < 4000, "no touching configuration found"
        st1, c1 = random_states(m, 1, seed=int(rng.integers(1 << 30)))
        q = st1["qpos"][0].astype(np.float64)
        q0, q1 = roots
        yaw = rng.uniform(-np.pi, np.pi)
        q[q1:q1 + 2] = q[q0:q0 + 2] + rng.uniform(-0.07, 0.07, 2)
        q[q1 + 2] = q[q0 + 2] + rng.uniform(-0.005, 0.005)
        q[q1 + 3:q1 + 7] = [np.cos(yaw / 2), 0, 0, np.sin(yaw / 2)]
        o.set_state(q, st1["qvel"][0].astype(np.float64), st1["act"][0].astype(np.float64), st1["qacc_warmstart"][0].astype(np.float64),
                    c1[0].astype(np.float64))
        o.forward()
        d = np.asarray(o.d.con_dist)
        # penetrating but not absurdly deep (a few mm at most: the regime a simulation can reach)
        act = cross & (d < 0)
        if act.sum() >= min_active and d[cross].min() > -0.004:
            for k in out:
                out[k][found] = st1[k][0]
            out["qpos"][found] = q.astype(np.float32)
            ctrls[found] = c1[0]
            found += 1
    return out, ctrls


def check_forward_intermediates(backend, name, N=8, states=None):
    """kinematics / smooth forces / tree-sparse factor+solve / collision / CG vs the dense float64 oracle."""
    m, cfg, clip, tables = common.setup(name)
    o, _ = common.oracles(name)
    o32 = oracle_mod.Oracle(m, np.float32)
    st, ctrl = random_states(m, N) if states is None else states
    full, cdist, niter = backend.forward_debug(st, ctrl, 0)
    at5, _, _ = backend.forward_debug(st, ctrl, 5)
    R = lambda s, e, nm, size: common.region(tables, s[e], nm, size).astype(np.float64)
    ties, outliers = [], []
    for e in range(N):
        o.set_state(st["qpos"][e].astype(np.float64), st["qvel"][e].astype(np.float64), st["act"][e].astype(np.float64),
                    st["qacc_warmstart"][e].astype(np.float64), ctrl[e].astype(np.float64))
        o.forward()
        d = o.d
        np.testing.assert_allclose(R(full, e, "xpos", 3 * m.nbody), d.xpos, atol=2e-6)          # fp32 rounding over a 39-deep chain
        np.testing.assert_allclose(R(full, e, "xquat", 4 * m.nbody), d.xquat, atol=2e-6)
        sc = np.abs(d.qfrc_smooth).max()
        np.testing.assert_allclose(R(at5, e, "qfrc_smooth", m.nv), d.qfrc_smooth, atol=2e-5 * sc)
        sa = np.abs(d.qacc_smooth).max()
        np.testing.assert_allclose(R(at5, e, "qacc_smooth", m.nv), d.qacc_smooth, atol=5e-4 * sa)  # cond(M) ~ 1e4..1e5
        tie = False
        if m.a["pair_ncon"].sum():
            np.testing.assert_allclose(cdist[e], d.con_dist, atol=2e-6)
            # The reference's collision functions break ties discontinuously (parallel capsules: which end of the overlap
            # carries the contact point; frame tangents at the |n_y| = 0.5 switch).  Where rounding decides such a tie
            # differently, distance and normal still agree but the contact point / tangent basis jump, and the truncated CG
            # lands somewhere else: those environments are checked up to the collision stage only.
            nc = cdist.shape[1]
            geo = common.region(tables, full[e], "T", 12 * nc).reshape(nc, 12).astype(np.float64)
            refp = common.region(tables, full[e], "ref", 3 * int(tables["nroot"][0])).reshape(-1, 3).astype(np.float64)
            pos = geo[:, :3] + refp[tables["con_ref"][:nc]]
            act = np.asarray(d.con_dist) < np.asarray(tables["con_includemargin"][:nc], np.float64)
            dpos = np.abs(pos - np.asarray(d.con_pos).reshape(nc, 3)).max(1)
            dfr = np.abs(geo[:, 3:] - np.asarray(d.con_frame).reshape(nc, 9)).max(1)
            np.testing.assert_allclose(geo[act, 3:6], np.asarray(d.con_frame).reshape(nc, 9)[act, :3], atol=5e-6)   # normals always agree
            tie = bool(((dpos > 1e-4) | (dfr > 1e-4))[act].any())
            ties.append(tie)
        if tie:
            continue
        # the truncated CG (4 iterations) amplifies rounding near cone-zone / active-set switches: allow 10x the float32
        # oracle's own departure from float64 on top of the fp32 bound
        o32.set_state(st["qpos"][e], st["qvel"][e], st["act"][e], st["qacc_warmstart"][e], ctrl[e])
        o32.forward()
        slack_a = 10 * np.abs(o32.d.qacc.astype(np.float64) - d.qacc).max()
        slack_f = 10 * np.abs(o32.d.qfrc_constraint.astype(np.float64) - d.qfrc_constraint).max()
        # ... and a cone-zone switch decided differently by rounding moves the 4-iteration iterate by a finite amount: at most
        # one environment in 16 may leave the tight bound, and then by no more than 5 % of the acceleration scale
        sq = max(sa, np.abs(d.qacc).max())
        sf = max(np.abs(d.qfrc_constraint).max(), 1e-3)
        ea = np.abs(R(full, e, "qacc", m.nv) - d.qacc).max()
        ef = np.abs(R(full, e, "qfrc_c", m.nv) - d.qfrc_constraint).max()
        if ea <= 2e-3 * sq + slack_a and ef <= 5e-2 * sf + slack_f:
            continue
        outliers.append(e)
        assert ea <= 5e-2 * sq and ef <= 0.5 * sf, (e, ea, sq, ef, sf)
    assert len(outliers) <= max(1, N // 16), f"CG outputs off the tight bound in environments {outliers}"
    assert sum(ties) <= max(1, N // 8), f"{sum(ties)} of {N} environments on a collision tie-break"


def check_reset(backend, name, N=32, seed=3):
    m, cfg, clip, tables = common.setup(name)
    _, eo = common.oracles(name)
    keys = common.jax_keys(N, seed=seed)
    st, out = backend.reset(keys)
    s0 = eo.reset(keys)
    assert np.array_equal(out["info_i"][:, 0], s0["info"]["cur_frame"])               # threefry randint: bit exact
    assert np.array_equal(out["info_i"][:, 1], s0["info"]["steps_taken_cur_frame"])
    assert np.array_equal(st["qvel"], s0["pipeline_state"]["qvel"])                    # threefry uniform: bit exact
    np.testing.assert_allclose(st["qpos"], s0["pipeline_state"]["qpos"], atol=1.2e-7)   # root quaternion renormalised (1 ulp)
    np.testing.assert_allclose(out["obs"], s0["obs"], atol=2e-5)
    sw = np.abs(s0["pipeline_state"]["qacc_warmstart"]).max()
    np.testing.assert_allclose(st["qacc_warmstart"], s0["pipeline_state"]["qacc_warmstart"], atol=1e-3 * sw)
    assert not out["done"].any() and not out["reward"].any() and not out["metrics"].any()
    return st, out, s0


FLAG_METRICS = ("too_far", "bad_pose", "bad_quat", "fall")
FLOAT_METRICS = ("pos_reward", "quat_reward", "joint_reward", "angvel_reward", "bodypos_reward", "endeff_reward", "reward_quadctrl",
                 "reward_alive")
INFO_FLOATS = ("summed_pos_distance", "quat_distance", "joint_distance")


def _metric_scale(cfg, k):
    """absolute scale of a reward-term metric = its weight (fruitfly.py:514-537: w * exp(-c * d)); 1 for the others"""
    return max(abs(float(cfg.get(k + "_weight", 1.0))), 1.0) if k.endswith("_reward") else 1.0


def check_teacher_forced(backend, name, N=16, T=100, seed=5, act_scale=0.3, episode_length=None, start_frames=None, require_done=True):
    """Wrapped step (AutoReset o Episode o env.step), every control step started from the ORACLE's state.
    `start_frames`: override the per-environment clip frame after the reset (late-clip cases: the `cur_frame + 1` window of
    `_get_obs` clamps to `[T - ref_len, T)`, fruitfly.py:602-611, and `clip[cur_frame]` clamps to the last frame)."""
    m, cfg, clip, tables = common.setup(name, episode_length)
    o64, eo = common.oracles(name, np.float64, episode_length)
    o32 = oracle_mod.Oracle(m, np.float32)
    keys = common.jax_keys(N, seed=seed)
    acts = common.actions(T, N, m.nu, seed=seed + 2, scale=act_scale)
    s = eo.reset(keys, start_frames=start_frames)
    first = state_from_oracle(s["info"]["first_pipeline_state"], N)
    first_obs = np.array(s["info"]["first_obs"], np.float32)
    first_ii = np.stack([s["info"]["first_cur_frame"], s["info"]["first_steps_taken_cur_frame"]], 1).astype(np.int32)
    n_done = n_live = 0
    wmax = max(_metric_scale(cfg, k) for k in FLOAT_METRICS)   # reward terms scale with their weights (fly joint reward: 50)
    eq, ev, oq, ov, er, eo_med, eo_max = [], [], [], [], [], [], []
    em = {k: [] for k in FLOAT_METRICS + INFO_FLOATS}
    max_frame = 0
    for t in range(T):
        st = state_from_oracle(s["pipeline_state"], N)
        out = backend.new_outputs(N)
        out["done"][:] = s["done"]
        out["info_f"][:, 3] = s["info"]["steps"]
        out["info_i"][:, 0] = s["info"]["cur_frame"]
        out["info_i"][:, 1] = s["info"]["steps_taken_cur_frame"]
        p32 = o32.pipeline_batch({k: np.array(v, np.float32) for k, v in s["pipeline_state"].items()}, acts[t], cfg["n_frames"])
        backend.step(st, out, first, first_obs, first_ii, acts[t])
        s_prev = s
        s = eo.step(s, acts[t])
        # ---- bit-exact: done flags, frame counters, episode counters, reset selection
        assert np.array_equal(out["done"], s["done"]), f"done flags differ at step {t}"
        assert np.array_equal(out["info_i"][:, 0], s["info"]["cur_frame"]), f"cur_frame differs at step {t}"
        assert np.array_equal(out["info_i"][:, 1], s["info"]["steps_taken_cur_frame"])
        assert np.array_equal(out["info_f"][:, 3], s["info"]["steps"])
        assert np.array_equal(out["info_f"][:, 4], s["info"]["truncation"])
        for k in FLAG_METRICS:
            assert np.array_equal(out["metrics"][:, common_metric(k)], s["metrics"][k]), f"{k} differs at step {t}"
        d = s["done"] > 0
        n_done += int(d.sum())
        if d.any():  # auto-reset restored the cached first state exactly
            for k in ("qpos", "qvel", "act", "qacc_warmstart", "time", "xpos"):
                assert np.array_equal(st[k][d], first[k][d]), k
            assert np.array_equal(out["obs"][d], first_obs[d])
        nd = ~d
        n_live += int(nd.sum())
        ref = s["pipeline_state"]
        eq.append(np.abs(st["qpos"] - ref["qpos"])[nd].max(1)); ev.append(np.abs(st["qvel"] - ref["qvel"])[nd].max(1))
        oq.append(np.abs(p32["qpos"] - ref["qpos"])[nd].max(1)); ov.append(np.abs(p32["qvel"] - ref["qvel"])[nd].max(1))
        er.append(np.abs(out["reward"] - s["reward"]))
        # ---- env layer of every LIVE environment, every step, at fp32 rounding level: the oracle's reward / obs functions
        #      evaluated on the PRODUCT's post-step physics state (so the one-step physics departure does not blur it):
        #      post-step observation (fruitfly.py:554 -> :598-646: the window clip[cur_frame + 1 : cur_frame + 1 + ref_len] AFTER
        #      the frame counter advanced), the 8 float metrics, the 3 info floats (fruitfly.py:530,548-549,579-592), the reward
        if nd.any():
            over = {k: (st[k].reshape(N, m.nbody, 3) if k == "xpos" else st[k]) for k in st}
            with np.errstate(all="ignore"):
                chk = eo.step(s_prev, acts[t], physics_override=over)
            np.testing.assert_allclose(out["obs"][nd], chk["obs"][nd], atol=2e-5, rtol=1e-6, err_msg=f"obs at step {t}")
            np.testing.assert_allclose(out["reward"][nd], chk["reward"][nd], atol=1e-4 * wmax, err_msg=f"reward at step {t}")
            for k in FLOAT_METRICS:
                np.testing.assert_allclose(out["metrics"][nd, common_metric(k)], chk["metrics"][k][nd], atol=1e-4 * _metric_scale(cfg, k),
                                           err_msg=f"{k} at step {t}")
            for i, k in enumerate(INFO_FLOATS):
                np.testing.assert_allclose(out["info_f"][nd, i], chk["info"][k][nd], atol=1e-5, rtol=1e-4, err_msg=f"{k} at step {t}")
            # ... and against the oracle's own physics: bounded by the one-step state departure
            eobs = np.abs(out["obs"][nd] - s["obs"][nd])
            eo_med.append(np.median(eobs, 1)); eo_max.append(eobs.max(1))
        # terminal-step values of every environment (the auto-reset does not touch metrics / info floats): bulk only
        for k in FLOAT_METRICS:
            em[k].append(np.abs(out["metrics"][:, common_metric(k)] - s["metrics"][k]) / _metric_scale(cfg, k))
        for i, k in enumerate(INFO_FLOATS):
            em[k].append(np.abs(out["info_f"][:, i] - s["info"][k]) / np.maximum(1.0, np.abs(s["info"][k])))
        max_frame = max(max_frame, int(s["info"]["cur_frame"].max()))
    assert n_done > 0 or not require_done, "the trajectory never exercised the auto-reset path"
    assert n_live > 0, "no live environment was ever compared"
    eq, ev, oq, ov, er, eo_med, eo_max = (np.concatenate(x) for x in (eq, ev, oq, ov, er, eo_med, eo_max))
    # ---- fp32 tolerance after ONE control step (n_frames substeps) from identical inputs:
    #      absolute bounds on the bulk, and no worse than 3x (bulk) / 10x (extreme tail) the float32 oracle's own departure from float64
    assert np.median(eq) < 1e-4 and np.median(ev) < 2e-2, (np.median(eq), np.median(ev))
    for p, k in ((50, 3), (90, 3), (99, 10), (100, 10)):   # the extreme tail of a few hundred samples is itself noisy
        assert np.percentile(eq, p) <= k * np.percentile(oq, p) + 1e-5, (p, np.percentile(eq, p), np.percentile(oq, p))
        assert np.percentile(ev, p) <= k * np.percentile(ov, p) + 1e-3, (p, np.percentile(ev, p), np.percentile(ov, p))
    # reward / reward terms: |err| relative to the term's weight
    assert np.median(er) < 1e-4 * wmax and er.max() < 2e-2 * wmax, (np.median(er), er.max())
    # observation rows: typical element at fp32 rounding, worst element bounded by the one-step qvel tolerance
    assert np.median(eo_med) < 1e-4 and eo_max.max() < max(2e-2, 1.5 * ev.max()), (np.median(eo_med), eo_max.max(), ev.max())
    res = dict(n_done=n_done, n_live=n_live, max_frame=max_frame, qpos_med=float(np.median(eq)), qpos_max=float(eq.max()),
               qvel_med=float(np.median(ev)), qvel_max=float(ev.max()), reward_med=float(np.median(er)), reward_max=float(er.max()),
               obs_med=float(np.median(eo_med)), obs_max=float(eo_max.max()))
    for k, v in em.items():
        v = np.concatenate(v)
        assert np.isfinite(v).all(), k
        assert np.median(v) < 1e-4, (k, np.median(v), v.max(), res)
        res[k + "_max"] = float(v.max())
    return res


def check_late_clip(backend, name, N=16, T=24, episode_length=None):
    """Clip-end clamps: start frames T-14 .. T+1 so that within the run `cur_frame + 1 + ref_len` passes the end of the clip
    (dynamic_slice clamps the window start, fruitfly.py:602-611) and `cur_frame` itself passes T - 1 (gather clamps)."""
    m, cfg, clip, tables = common.setup(name, episode_length)
    Tc = int(np.asarray(clip["joints"]).shape[0])
    start = (Tc - 14 + np.arange(N) % 16).astype(np.int32)
    r = check_teacher_forced(backend, name, N=N, T=T, seed=9, episode_length=episode_length, start_frames=start, require_done=False)
    assert r["max_frame"] >= Tc, "the run never indexed past the end of the clip"
    return r


def check_unwrapped_step(backend, name, N=8, T=12, seed=13):
    """The bare env (no wrappers): `pipeline_init` (fruitfly.py:477), `env.step` = `pipeline_step` + reward / obs
    (fruitfly.py:497-596) through bt_pipeline_init / bt_physics_step / bt_reward_obs, teacher-forced against the oracle."""
    m, cfg, clip, tables = common.setup(name)
    o64, eo = common.oracles(name)
    keys = common.jax_keys(N, seed=seed)
    acts = common.actions(T, N, m.nu, seed=seed + 1, scale=0.3)
    s = eo.reset(keys, fixed_start_frame=0)         # RenderRolloutWrapperTracking.reset: frame 0
    # ---- pipeline_init = mjx.forward on (qpos, qvel): xpos and the warm start it leaves behind
    st = state_from_oracle(s["pipeline_state"], N)
    st["xpos"][:] = 0; st["qacc_warmstart"][:] = 0
    backend.pipeline_init(st)
    np.testing.assert_allclose(st["xpos"], s["pipeline_state"]["xpos"].reshape(N, -1), atol=2e-6)
    sw = np.abs(s["pipeline_state"]["qacc_warmstart"]).max()
    np.testing.assert_allclose(st["qacc_warmstart"], s["pipeline_state"]["qacc_warmstart"], atol=1e-3 * sw)
    wmax = max(_metric_scale(cfg, k) for k in FLOAT_METRICS)
    o32 = oracle_mod.Oracle(m, np.float32)
    eq, oq = [], []
    for t in range(T):
        st = state_from_oracle(s["pipeline_state"], N)
        out = backend.new_outputs(N)
        out["info_i"][:, 0] = s["info"]["cur_frame"]
        out["info_i"][:, 1] = s["info"]["steps_taken_cur_frame"]
        p32 = o32.pipeline_batch({k: np.array(v, np.float32) for k, v in s["pipeline_state"].items()}, acts[t], cfg["n_frames"])
        backend.physics_step(st, acts[t], cfg["n_frames"])
        backend.reward_obs(st, out, acts[t])
        # the env layer on the product's own physics state, at fp32 rounding level (see check_teacher_forced)
        over = {k: (st[k].reshape(N, m.nbody, 3) if k == "xpos" else st[k]) for k in st}
        chk = eo.reward_obs(s, over, acts[t])
        s = eo.env_step(s, acts[t])
        assert np.array_equal(out["done"], s["done"]), f"done differs at step {t}"
        assert np.array_equal(out["info_i"][:, 0], s["info"]["cur_frame"]) and np.array_equal(out["info_i"][:, 1], s["info"]["steps_taken_cur_frame"])
        for k in FLAG_METRICS:
            assert np.array_equal(out["metrics"][:, common_metric(k)], s["metrics"][k]), k
        np.testing.assert_allclose(out["obs"], chk["obs"], atol=2e-5, rtol=1e-6, err_msg=f"obs at step {t}")
        np.testing.assert_allclose(out["reward"], chk["reward"], atol=1e-4 * wmax)
        for k in FLOAT_METRICS:
            np.testing.assert_allclose(out["metrics"][:, common_metric(k)], chk["metrics"][k], atol=1e-4 * _metric_scale(cfg, k), err_msg=k)
        for i, k in enumerate(INFO_FLOATS):
            np.testing.assert_allclose(out["info_f"][:, i], chk["info"][k], atol=1e-5, rtol=1e-4, err_msg=k)
        eq.append(np.abs(st["qpos"] - s["pipeline_state"]["qpos"]).max(1)); oq.append(np.abs(p32["qpos"] - s["pipeline_state"]["qpos"]).max(1))
        assert np.abs(out["reward"] - s["reward"]).max() < 2e-2 * wmax
    eq, oq = np.concatenate(eq), np.concatenate(oq)
    # physics of the unwrapped step: as in check_teacher_forced, relative to the float32 oracle's own one-step departure
    assert np.median(eq) < 1e-4 and eq.max() <= 10 * oq.max() + 1e-5, (np.median(eq), eq.max(), oq.max())
    stats = dict(qpos_med=float(np.median(eq)), qpos_max=float(eq.max()), o32_qpos_max=float(oq.max()))
    assert int(s["info"]["cur_frame"].max()) == T // 2      # two control steps per mocap frame, from frame 0
    return stats


def check_nan_guard(backend, name, N=12, seed=17):
    """NaN guard (fruitfly.py:569-577): a NaN anywhere in the pipeline state => done = 1 with nan_to_num'ed reward / obs; the
    auto-reset wrapper then restores the cached first state, and the next step is clean.  Also the element-wise nan_to_num
    of the observation row (NaN -> 0, +-inf -> +-FLT_MAX) on the unwrapped path, where no reset replaces the row."""
    m, cfg, clip, tables = common.setup(name)
    o64, eo = common.oracles(name)
    keys = common.jax_keys(N, seed=seed)
    acts = common.actions(3, N, m.nu, seed=seed + 1, scale=0.3)
    s = eo.reset(keys)
    first = state_from_oracle(s["info"]["first_pipeline_state"], N)
    first_obs = np.array(s["info"]["first_obs"], np.float32)
    first_ii = np.stack([s["info"]["first_cur_frame"], s["info"]["first_steps_taken_cur_frame"]], 1).astype(np.int32)
    s = eo.step(s, acts[0])                                  # one clean step first
    free = m.jnt_type[0] == 0
    ps = {k: np.array(v, copy=True) for k, v in s["pipeline_state"].items()}
    a1 = acts[1].copy()
    hinge0 = 7 if free else 0
    ps["qpos"][0, hinge0 + 3] = np.nan                       # a joint angle
    ps["qvel"][1, m.nv - 1] = np.nan                         # a leaf joint velocity
    if free:
        ps["qpos"][2, 4] = np.nan                            # root quaternion component
    if m.na:
        ps["act"][3, 0] = np.nan                             # actuator activation
    a1[4, 1] = np.nan                                        # the action itself (data.ctrl is part of the flattened Data)
    ps["qacc_warmstart"][5, 2] = np.nan                      # warm start only: `warm.cost < smooth.cost` is False -> qacc_smooth, no NaN survives
    poisoned = np.zeros(N, bool); poisoned[[0, 1, 4]] = True
    if free: poisoned[2] = True
    if m.na: poisoned[3] = True
    s["pipeline_state"] = ps
    st = state_from_oracle(ps, N)
    out = backend.new_outputs(N)
    out["done"][:] = s["done"]; out["info_f"][:, 3] = s["info"]["steps"]
    out["info_i"][:, 0] = s["info"]["cur_frame"]; out["info_i"][:, 1] = s["info"]["steps_taken_cur_frame"]
    backend.step(st, out, first, first_obs, first_ii, a1)
    with np.errstate(all="ignore"):
        s = eo.step(s, a1)
    assert (s["done"][poisoned] == 1).all(), "the oracle itself did not flag the poisoned environments"
    assert np.array_equal(out["done"], s["done"]), (out["done"], s["done"])
    assert s["done"][5] == 0 and np.isfinite(st["qpos"][5]).all()                 # NaN warm start alone is harmless
    assert np.isfinite(out["reward"]).all() and np.isfinite(out["obs"]).all()
    # nan_to_num(reward): a NaN reward reads 0 in both
    np.testing.assert_array_equal(out["reward"][poisoned] == 0.0, s["reward"][poisoned] == 0.0)
    assert (out["reward"][[0, 4]] == 0.0).all()               # joint / ctrl sums contain the NaN for certain
    for k in ("qpos", "qvel", "act", "qacc_warmstart", "time", "xpos"):          # restored from the cached first state
        assert np.array_equal(st[k][poisoned], first[k][poisoned]), k
        assert np.isfinite(st[k]).all(), k
    assert np.array_equal(out["obs"][poisoned], first_obs[poisoned])
    assert np.array_equal(out["info_i"][poisoned, 0], first_ii[poisoned, 0])
    live = s["done"] == 0
    np.testing.assert_allclose(out["reward"][live], s["reward"][live], atol=2e-2 * max(_metric_scale(cfg, k) for k in FLOAT_METRICS))
    # ---- the step after: every environment is clean again and tracks the oracle
    st = state_from_oracle(s["pipeline_state"], N)
    out["done"][:] = s["done"]; out["info_f"][:, 3] = s["info"]["steps"]
    out["info_i"][:, 0] = s["info"]["cur_frame"]; out["info_i"][:, 1] = s["info"]["steps_taken_cur_frame"]
    backend.step(st, out, first, first_obs, first_ii, acts[2])
    s = eo.step(s, acts[2])
    assert np.array_equal(out["done"], s["done"]) and np.array_equal(out["info_f"][:, 3], s["info"]["steps"])
    assert (out["info_f"][poisoned, 3] == 1).all()            # episode counter restarted (custom_wrappers.py:55-58)
    dq = np.abs(st["qpos"] - s["pipeline_state"]["qpos"])      # one control step in contact: bulk at fp32 rounding (check_teacher_forced has the distribution test)
    assert np.median(dq) < 1e-5 and dq.max() < 2e-3, (np.median(dq), dq.max())
    # ---- unwrapped: element-wise nan_to_num of the observation row and of the reward
    ps = {k: np.array(v, copy=True) for k, v in s["pipeline_state"].items()}
    ps["qvel"][0, 3] = np.nan; ps["qvel"][1, 5] = np.inf; ps["qvel"][2, 6] = -np.inf; ps["xpos"].reshape(N, -1)[3, 3 * 5 + 1] = np.nan
    st = state_from_oracle(ps, N)
    o2 = backend.new_outputs(N)
    o2["info_i"][:, 0] = s["info"]["cur_frame"]; o2["info_i"][:, 1] = s["info"]["steps_taken_cur_frame"]
    backend.reward_obs(st, o2, acts[2])
    with np.errstate(all="ignore"):
        ref = eo.reward_obs(dict(s), {k: np.asarray(v, np.float32) for k, v in ps.items()}, acts[2])
    assert np.isfinite(o2["obs"]).all() and np.isfinite(o2["reward"]).all()
    assert np.array_equal(o2["done"], ref["done"]) and o2["done"][0] == 1 and o2["done"][3] == 1 and o2["done"][1] == ref["done"][1]
    fmax = np.finfo(np.float32).max
    assert o2["obs"][0, m.nq + 3] == 0.0 and o2["obs"][1, m.nq + 5] == fmax and o2["obs"][2, m.nq + 6] == -fmax
    np.testing.assert_allclose(o2["obs"], ref["obs"], atol=2e-5, rtol=1e-6)
    return dict(poisoned=int(poisoned.sum()))


def common_metric(name):
    from brax_tracking_b200 import native
    return native.METRIC_NAMES.index(name)


def check_physics_1_10_100(backend, name, N=8, seed=11):
    """Free-running qpos/qvel after 1, 10 and 100 control steps.
      *   1 step : fp32 tolerance against the float64 oracle.
      *  10 steps: the typical (median over envs) departure from the float64 oracle stays within 10x of the float32 oracle's
                   own departure -- same arithmetic precision, same chaos.
      * 100 steps: both float32 trajectories have decorrelated from the float64 one by then (the divergence saturates at the
                   size of the attractor), so only statistics are comparable: states stay finite and inside the joint
                   ranges, the ensemble root height agrees, and the ZERO-ACTION trajectory (a contraction: the animal
                   settles) comes to rest at the same height."""
    m, cfg, clip, tables = common.setup(name)
    o64, eo = common.oracles(name)
    o32 = oracle_mod.Oracle(m, np.float32)
    keys = common.jax_keys(N, seed=seed)
    s0 = eo.reset(keys)
    res = {}
    free = m.jnt_type[0] == 0
    lim = m.jnt_limited.astype(bool) & (m.jnt_type == 3)
    qa = m.jnt_qposadr[lim]
    lo, hi = m.jnt_range[lim, 0], m.jnt_range[lim, 1]
    for label, scale in (("zero", 0.0), ("policy", 0.3)):
        acts = common.actions(100, N, m.nu, seed=seed, scale=scale)
        p64 = {k: np.asarray(v, np.float64) for k, v in s0["pipeline_state"].items()}
        p32 = {k: np.array(v, np.float32) for k, v in s0["pipeline_state"].items()}
        st = state_from_oracle(s0["pipeline_state"], N)
        for t in range(1, 101):
            p64 = o64.pipeline_batch(p64, acts[t - 1].astype(np.float64), cfg["n_frames"])
            p32 = o32.pipeline_batch(p32, acts[t - 1], cfg["n_frames"])
            backend.physics_step(st, acts[t - 1], cfg["n_frames"])
            if t in (1, 10, 100):
                e_env = np.abs(st["qpos"] - p64["qpos"]).max(1); e32_env = np.abs(p32["qpos"] - p64["qpos"]).max(1)
                v_env = np.abs(st["qvel"] - p64["qvel"]).max(1); v32_env = np.abs(p32["qvel"] - p64["qvel"]).max(1)
                res[(label, t)] = (float(np.median(e_env)), float(np.median(e32_env)), float(np.median(v_env)), float(np.median(v32_env)))
                assert np.isfinite(st["qpos"]).all() and np.isfinite(st["qvel"]).all()
                if t == 1:
                    assert e_env.max() < 5e-5 and v_env.max() < 2e-2, (label, t, e_env.max(), v_env.max())  # fp32 tolerance, 1 control step
                elif t == 10:
                    assert np.median(e_env) <= 10 * np.median(e32_env) + 1e-3, (label, t, np.median(e_env), np.median(e32_env))
                    assert np.median(v_env) <= 10 * np.median(v32_env) + 5e-2, (label, t, np.median(v_env), np.median(v32_env))
                else:
                    q = st["qpos"][:, qa]
                    assert (q > lo - 0.5).all() and (q < hi + 0.5).all(), "joint angles left their ranges"
                    if free:
                        assert abs(st["qpos"][:, 2].mean() - p64["qpos"][:, 2].mean()) < 2e-2, "ensemble root height"
        if label == "zero" and free:
            np.testing.assert_allclose(st["qpos"][:, 2], p64["qpos"][:, 2], atol=5e-3)   # same resting height
    return res


def check_multi_clip(make_backend, N=24, T=12, n_clips=3, seed=19):
    """RodentMultiClip: n_clips clips stacked on a leading axis; the reset draws every environment's
    clip (randint(rng_pos, (), 0, n_clips), bit exact), seeds the root from THAT clip, and every clip gather of the step (reward
    targets, observation window) is offset by it; auto-resets keep the clip."""
    from brax_tracking_b200 import clips, configs, model, presets
    import env_oracle
    m, args, _ = presets.load("rodent")
    cfg = configs.resolve(m, args)
    cfg["episode_length"] = 6
    cs = [clips.synthetic_clip(m, True, seed=k, amplitude=0.2 + 0.1 * k).as_dict() for k in range(n_clips)]
    stacked = {k: np.stack([c[k] for c in cs]) for k in cs[0]}
    tables = model.pack(m, cfg, stacked)
    assert int(tables["n_clips"][0]) == n_clips
    b = make_backend(tables)
    o64, _ = common.oracles("rodent")
    eo = env_oracle.EnvOracle(o64, stacked, cfg, dtype=np.float32)
    keys = common.jax_keys(N, seed=seed)
    cidx = np.full(N, -1, np.int32)
    st, out = b.reset(keys, clip_idx=cidx)
    s = eo.reset(keys)
    assert np.array_equal(cidx, s["info"]["clip_idx"]) and len(set(cidx.tolist())) == n_clips     # drawn clips: bit exact, all used
    assert np.array_equal(out["info_i"][:, 0], s["info"]["cur_frame"])
    np.testing.assert_allclose(st["qpos"], s["pipeline_state"]["qpos"], atol=1.2e-7)               # root seeded from the env's own clip
    np.testing.assert_allclose(out["obs"], s["obs"], atol=2e-5)
    # the same keys on clip 0 alone differ wherever another clip was drawn
    s_one = env_oracle.EnvOracle(o64, cs[0], cfg, dtype=np.float32).reset(keys)
    assert (np.abs(s_one["obs"] - s["obs"]).max(1) > 1e-3)[cidx != 0].all()
    first = state_from_oracle(s["info"]["first_pipeline_state"], N)
    first_obs = np.array(s["info"]["first_obs"], np.float32)
    first_ii = np.stack([s["info"]["first_cur_frame"], s["info"]["first_steps_taken_cur_frame"]], 1).astype(np.int32)
    acts = common.actions(T, N, m.nu, seed=seed + 1, scale=0.3)
    n_done = 0
    for t in range(T):
        stt = state_from_oracle(s["pipeline_state"], N)
        o2 = b.new_outputs(N)
        o2["done"][:] = s["done"]; o2["info_f"][:, 3] = s["info"]["steps"]
        o2["info_i"][:, 0] = s["info"]["cur_frame"]; o2["info_i"][:, 1] = s["info"]["steps_taken_cur_frame"]
        b.step(stt, o2, first, first_obs, first_ii, acts[t], clip_idx=cidx)
        over = {k: (stt[k].reshape(N, m.nbody, 3) if k == "xpos" else stt[k]) for k in stt}
        chk = eo.step(s, acts[t], physics_override=over)
        s = eo.step(s, acts[t])
        assert np.array_equal(o2["done"], s["done"]) and np.array_equal(o2["info_i"][:, 0], s["info"]["cur_frame"])
        live = s["done"] == 0
        n_done += int((~live).sum())
        np.testing.assert_allclose(o2["obs"][live], chk["obs"][live], atol=2e-5, rtol=1e-6)        # the env's own clip window
        np.testing.assert_allclose(o2["reward"][live], chk["reward"][live], atol=1e-4)
        assert np.array_equal(o2["obs"][~live], first_obs[~live])
        assert np.array_equal(s["info"]["clip_idx"], cidx)                                          # the clip survives auto-resets
    assert n_done > 0
    return dict(clips=np.bincount(cidx, minlength=n_clips).tolist(), n_done=n_done)
