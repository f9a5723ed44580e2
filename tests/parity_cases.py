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
    free = m.jnt_type[0] == 0
    for e in range(N):
        q = m.qpos0.copy()
        hinge = slice(7, None) if free else slice(0, None)
        q[hinge] += rng.uniform(-0.2, 0.2, q[hinge].shape)
        if free:
            q[2] += rng.uniform(-0.01, 0.02)
            quat = np.array([1.0, 0, 0, 0]) + rng.uniform(-0.2, 0.2, 4)
            q[3:7] = quat / np.linalg.norm(quat)
        st["qpos"][e] = q
        st["qvel"][e] = rng.uniform(-1, 1, m.nv)
        st["act"][e] = rng.uniform(-0.5, 0.5, m.na)
        st["qacc_warmstart"][e] = rng.uniform(-5, 5, m.nv)
    ctrl = rng.uniform(-1, 1, (N, m.nu)).astype(np.float32)
    return st, ctrl


def check_forward_intermediates(backend, name, N=8):
    """kinematics / smooth forces / tree-sparse factor+solve / collision / CG vs the dense float64 oracle."""
    m, cfg, clip, tables = common.setup(name)
    o, _ = common.oracles(name)
    o32 = oracle_mod.Oracle(m, np.float32)
    st, ctrl = random_states(m, N)
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


def check_teacher_forced(backend, name, N=16, T=100, seed=5, act_scale=0.3, episode_length=None):
    m, cfg, clip, tables = common.setup(name, episode_length)
    o64, eo = common.oracles(name, np.float64, episode_length)
    o32 = oracle_mod.Oracle(m, np.float32)
    keys = common.jax_keys(N, seed=seed)
    acts = common.actions(T, N, m.nu, seed=seed + 2, scale=act_scale)
    s = eo.reset(keys)
    first = state_from_oracle(s["info"]["first_pipeline_state"], N)
    first_obs = np.array(s["info"]["first_obs"], np.float32)
    first_ii = np.stack([s["info"]["first_cur_frame"], s["info"]["first_steps_taken_cur_frame"]], 1).astype(np.int32)
    n_done = 0
    eq, ev, oq, ov, er = [], [], [], [], []
    for t in range(T):
        st = state_from_oracle(s["pipeline_state"], N)
        out = backend.new_outputs(N)
        out["done"][:] = s["done"]
        out["info_f"][:, 3] = s["info"]["steps"]
        out["info_i"][:, 0] = s["info"]["cur_frame"]
        out["info_i"][:, 1] = s["info"]["steps_taken_cur_frame"]
        p32 = o32.pipeline_batch({k: np.array(v, np.float32) for k, v in s["pipeline_state"].items()}, acts[t], cfg["n_frames"])
        backend.step(st, out, first, first_obs, first_ii, acts[t])
        s = eo.step(s, acts[t])
        # ---- bit-exact: done flags, frame counters, episode counters, reset selection
        assert np.array_equal(out["done"], s["done"]), f"done flags differ at step {t}"
        assert np.array_equal(out["info_i"][:, 0], s["info"]["cur_frame"]), f"cur_frame differs at step {t}"
        assert np.array_equal(out["info_i"][:, 1], s["info"]["steps_taken_cur_frame"])
        assert np.array_equal(out["info_f"][:, 3], s["info"]["steps"])
        assert np.array_equal(out["info_f"][:, 4], s["info"]["truncation"])
        for k in ("too_far", "bad_pose", "bad_quat", "fall"):
            assert np.array_equal(out["metrics"][:, common_metric(k)], s["metrics"][k]), f"{k} differs at step {t}"
        d = s["done"] > 0
        n_done += int(d.sum())
        if d.any():  # auto-reset restored the cached first state exactly
            assert np.array_equal(st["qpos"][d], first["qpos"][d]) and np.array_equal(out["obs"][d], first_obs[d])
        nd = ~d
        ref = s["pipeline_state"]
        eq.append(np.abs(st["qpos"] - ref["qpos"])[nd].max(1)); ev.append(np.abs(st["qvel"] - ref["qvel"])[nd].max(1))
        oq.append(np.abs(p32["qpos"] - ref["qpos"])[nd].max(1)); ov.append(np.abs(p32["qvel"] - ref["qvel"])[nd].max(1))
        er.append(np.abs(out["reward"] - s["reward"]))
    assert n_done > 0, "the trajectory never exercised the auto-reset path"
    eq, ev, oq, ov, er = (np.concatenate(x) for x in (eq, ev, oq, ov, er))
    # ---- fp32 tolerance after ONE control step (n_frames substeps) from identical inputs:
    #      absolute bounds on the bulk, and no worse than 3x (bulk) / 10x (extreme tail) the float32 oracle's own departure from float64
    assert np.median(eq) < 1e-4 and np.median(ev) < 2e-2, (np.median(eq), np.median(ev))
    for p, k in ((50, 3), (90, 3), (99, 10), (100, 10)):   # the extreme tail of a few hundred samples is itself noisy
        assert np.percentile(eq, p) <= k * np.percentile(oq, p) + 1e-5, (p, np.percentile(eq, p), np.percentile(oq, p))
        assert np.percentile(ev, p) <= k * np.percentile(ov, p) + 1e-3, (p, np.percentile(ev, p), np.percentile(ov, p))
    assert np.median(er) < 1e-4 and er.max() < 2e-2, (np.median(er), er.max())
    return dict(n_done=n_done, qpos_med=float(np.median(eq)), qpos_max=float(eq.max()), qvel_med=float(np.median(ev)))


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
