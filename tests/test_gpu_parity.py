"""GPU parity tests proper: the sm_100a kernels, called through the C ABI, against the oracle on identical inputs."""
import numpy as np
import pytest

import common
import parity_cases as pc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rodent_cuda():
    from backends import CudaBackend
    return CudaBackend(common.setup("rodent")[3])


def test_forward_intermediates(rodent_cuda):
    pc.check_forward_intermediates(rodent_cuda, "rodent", N=16)


def test_reset(rodent_cuda):
    pc.check_reset(rodent_cuda, "rodent", N=64)


def test_teacher_forced_wrapped_step_100(rodent_cuda):
    r = pc.check_teacher_forced(rodent_cuda, "rodent", N=16, T=100)
    print(r)


def test_late_clip_window_clamps(rodent_cuda):
    """`cur_frame + 1` window clamp of _get_obs (fruitfly.py:602-611) and the clip[cur_frame] gather clamp, past the clip end."""
    print(pc.check_late_clip(rodent_cuda, "rodent", N=16, T=24))


def test_nan_guard(rodent_cuda):
    """fruitfly.py:569-577: NaN in qpos / qvel / act / action => done, finite reward / obs, restore from the cached first state."""
    print(pc.check_nan_guard(rodent_cuda, "rodent"))


def test_unwrapped_env_step_and_pipeline_init(rodent_cuda):
    """bt_pipeline_init, bt_physics_step + bt_reward_obs (the bare env.step, fruitfly.py:497-596) against the oracle."""
    print(pc.check_unwrapped_step(rodent_cuda, "rodent", N=8, T=12))


def test_env_classes_against_the_oracle():
    """The host-side mirror of the reference classes on CUDA: TrackingEnv.reset / .step / .pipeline_init,
    RenderRolloutWrapperTracking (custom_wrappers.py:82-125) and ppo.evaluate_rollout (main.py:136-258) traces."""
    import torch
    from brax_tracking_b200 import envs, native, ppo
    m, cfg, clip, _ = common.setup("rodent")
    _, eo = common.oracles("rodent")
    n, T = 8, 10
    keys = common.jax_keys(n, seed=31)
    acts = common.actions(T, n, m.nu, seed=32, scale=0.3)
    env = envs.RodentSingleClip(clip, mj_model=m)
    # ---- bare env: reset draws the training start frames; step = pipeline_step + reward / obs, no wrappers
    s = eo.reset(keys)
    state = env.reset(keys)
    assert np.array_equal(state.info["cur_frame"].cpu().numpy(), s["info"]["cur_frame"])
    np.testing.assert_allclose(state.obs.cpu().numpy(), s["obs"], atol=2e-5)
    assert "steps" not in state.info and "first_obs" not in state.info            # those belong to the wrappers
    for t in range(3):
        # teacher forcing: the oracle's state goes in, one env.step comes out
        for k in native.STATE_FIELDS:
            state.pipeline_state[k].copy_(torch.from_numpy(np.asarray(s["pipeline_state"][k], np.float32).reshape(state.pipeline_state[k].shape)))
        state._raw["info_i"][:, 0].copy_(torch.from_numpy(s["info"]["cur_frame"])); state._raw["info_i"][:, 1].copy_(torch.from_numpy(s["info"]["steps_taken_cur_frame"]))
        state = env.step(state, torch.from_numpy(acts[t]).cuda())
        # the oracle's env layer on the product's own post-step physics state: fp32 rounding level (parity_cases.check_teacher_forced)
        over = {k: state.pipeline_state[k].cpu().numpy() for k in native.STATE_FIELDS}
        over["xpos"] = over["xpos"].reshape(n, m.nbody, 3)
        chk = eo.reward_obs(s, over, acts[t])
        s = eo.env_step(s, acts[t])
        assert np.array_equal(state.done.cpu().numpy(), s["done"])
        assert np.array_equal(state.info["cur_frame"].cpu().numpy(), s["info"]["cur_frame"])
        np.testing.assert_allclose(state.obs.cpu().numpy(), chk["obs"], atol=2e-5, rtol=1e-6)
        np.testing.assert_allclose(state.reward.cpu().numpy(), chk["reward"], atol=1e-4)
        for k in pc.FLOAT_METRICS:
            np.testing.assert_allclose(state.metrics[k].cpu().numpy(), chk["metrics"][k], atol=1e-4, err_msg=k)
        for k in pc.INFO_FLOATS:
            np.testing.assert_allclose(state.info[k].cpu().numpy(), chk["info"][k], atol=1e-5, rtol=1e-4, err_msg=k)
        assert np.median(np.abs(state.obs.cpu().numpy() - s["obs"])) < 1e-4      # and the bulk against the oracle's own physics
        np.testing.assert_allclose(state.reward.cpu().numpy(), s["reward"], atol=2e-2)
    # ---- pipeline_init (fruitfly.py:477)
    ps = env.pipeline_init(torch.from_numpy(np.asarray(s["pipeline_state"]["qpos"], np.float32)).cuda(),
                           torch.from_numpy(np.asarray(s["pipeline_state"]["qvel"], np.float32)).cuda())
    o64, _ = common.oracles("rodent")
    fwd = o64.pipeline_batch(dict(s["pipeline_state"], act=np.zeros_like(s["pipeline_state"]["act"]),
                                  qacc_warmstart=np.zeros_like(s["pipeline_state"]["qacc_warmstart"])), None, 0, forward_only=True)
    np.testing.assert_allclose(ps["xpos"].cpu().numpy(), np.asarray(fwd["xpos"]).reshape(n, -1), atol=2e-6)
    # ---- render-rollout wrapper: frame 0, split(rng, 3), qpos0 + noise, bare steps
    renv = envs.RenderRolloutWrapperTracking(env)
    rs = renv.reset(keys)
    s0 = eo.reset(keys, fixed_start_frame=0)
    assert not rs.info["cur_frame"].any().item()
    assert np.array_equal(rs.pipeline_state["qvel"].cpu().numpy(), s0["pipeline_state"]["qvel"])
    np.testing.assert_allclose(rs.obs.cpu().numpy(), s0["obs"], atol=2e-5)
    # ---- evaluation rollout traces with an open-loop "policy" (free-running: the first steps stay within the 1-step bounds)
    it = iter(range(T))
    tr = ppo.evaluate_rollout(env, lambda obs: (torch.from_numpy(acts[next(it)]).cuda(),), keys, num_steps=T)
    so, ref_rew, ref_cur, ref_h = s0, [], [], []
    for t in range(T):
        so = eo.env_step(so, acts[t])
        ref_rew.append(so["reward"]); ref_cur.append(so["info"]["cur_frame"]); ref_h.append(so["pipeline_state"]["xpos"][:, cfg["torso_idx"], 2])
    assert np.array_equal(tr["cur_frame"], np.array(ref_cur))                      # frame index trace: exact, all steps
    np.testing.assert_allclose(tr["reward"][:2], np.array(ref_rew)[:2], atol=2e-2)
    np.testing.assert_allclose(tr["torso_height"][:2], np.array(ref_h)[:2], atol=1e-4)
    assert np.isfinite(tr["reward"]).all() and tr["pos_reward"].shape == (T, n)


def test_multi_clip_gather():
    """RodentMultiClip (envs/rodent.py:377; preprocess.py:254-258): clip_idx drawn in the reset kernel, gathered in the step kernel;
    and the host-side env class carrying info['clip_idx']."""
    import torch
    from backends import CudaBackend
    print(pc.check_multi_clip(CudaBackend, N=48, T=12))
    from brax_tracking_b200 import clips, envs
    m, cfg, clip, _ = common.setup("rodent")
    cs = [clips.synthetic_clip(m, True, seed=k, amplitude=0.2 + 0.1 * k) for k in range(3)]
    env = envs.wrap(envs.RodentMultiClip(cs, mj_model=m), episode_length=20)
    keys = common.jax_keys(64, seed=19)
    state = env.reset(keys)
    c0 = state.info["clip_idx"].clone()
    assert set(c0.cpu().tolist()) == {0, 1, 2}
    for t in range(25):
        state = env.step(state, torch.zeros(64, m.nu, device="cuda"))
    assert torch.equal(state.info["clip_idx"], c0) and torch.isfinite(state.obs).all()
    r = envs.RenderRolloutWrapperTracking(env.env).reset(keys[:4], clip_idx=[2, 1, 0, 2])
    assert r.info["clip_idx"].cpu().tolist() == [2, 1, 0, 2] and not r.info["cur_frame"].any().item()


def test_physics_1_10_100(rodent_cuda):
    print(pc.check_physics_1_10_100(rodent_cuda, "rodent", N=8))


@pytest.mark.parametrize("name", ["fly_free", "fly_tethered"])
def test_fly_elliptic_cone(name):
    """configs[2]: fruit-fly imitation env, 8192-env kernel variant (2 dof slots, 1 contact slot per lane)."""
    from backends import CudaBackend
    b = CudaBackend(common.setup(name)[3])
    pc.check_forward_intermediates(b, name, N=16)
    pc.check_reset(b, name, N=64)
    print(pc.check_physics_1_10_100(b, name, N=16))
    pc.check_unwrapped_step(b, name, N=8, T=8)
    pc.check_nan_guard(b, name)
    bt = CudaBackend(common.setup(name, 12)[3])
    print(pc.check_teacher_forced(bt, name, N=16, T=40, episode_length=12))
    if name == "fly_tethered":   # (the free fly is not seeded from the clip: at a late start frame it is `too_far` at every step)
        print(pc.check_late_clip(bt, name, N=16, T=24, episode_length=12))


@pytest.mark.parametrize("name", ["rodent", "fly_free", "fly_tethered", "rodent_pair"])
def test_cuda_matches_golden(name):
    from backends import CudaBackend
    from test_golden import run_against_golden
    run_against_golden(CudaBackend, name)


def test_full_size_properties(rodent_cuda):
    """BASELINE.json's full size (8192 envs, configs[1]) through size-independent properties: the batch result is
    deterministic, independent of batch size and of the position of an environment in the batch (environments never
    interact), which ties every row of the full-size run to the oracle-checked small runs; flags and counters stay exact."""
    b, N, T = rodent_cuda, 8192, 4
    m, cfg, clip, tables = common.setup("rodent")
    keys = common.jax_keys(N, seed=21)
    acts = common.actions(T, N, m.nu, seed=23, scale=0.5)

    def rollout(keys, acts):
        st, out = b.reset(keys)
        first = {k: v.copy() for k, v in st.items()}
        first_obs, first_ii = out["obs"].copy(), out["info_i"].copy()
        for t in range(acts.shape[0]):
            b.step(st, out, first, first_obs, first_ii, acts[t])
        return st, out

    st_a, out_a = rollout(keys, acts)
    st_b, out_b = rollout(keys, acts)
    for k in st_a:      # determinism: no atomics, fixed reduction order
        assert np.array_equal(st_a[k], st_b[k]), k
    for k in out_a:
        assert np.array_equal(out_a[k], out_b[k], equal_nan=True), k
    # batch-size independence: the first 48 environments of the 8192 batch == a 48-environment batch, bit for bit
    n = 48
    st_s, out_s = rollout(keys[:n], acts[:, :n])
    for k in st_a:
        assert np.array_equal(st_a[k][:n], st_s[k]), k
    for k in out_a:
        assert np.array_equal(out_a[k][:n], out_s[k], equal_nan=True), k
    # position independence: the reversed batch gives the reversed result
    st_r, out_r = rollout(keys[::-1].copy(), acts[:, ::-1].copy())
    for k in st_a:
        assert np.array_equal(st_a[k], st_r[k][::-1]), k
    assert np.array_equal(out_a["obs"], out_r["obs"][::-1]) and np.array_equal(out_a["reward"], out_r["reward"][::-1])
    # the same 48 environments against the oracle (teacher forcing is not needed for 4 steps: float tolerance of check_reset /
    # the 1-step bound compounded), integer outputs exact
    _, eo = common.oracles("rodent")
    s = eo.reset(keys[:n])
    alive = np.ones(n, bool)
    for t in range(T):
        s = eo.step(s, acts[t, :n])
        alive &= s["done"] == 0
    alive &= out_s["done"] == 0   # free-running: a threshold flag may legitimately flip at rounding level; compare the rest
    assert alive.sum() > n // 2
    assert np.array_equal(out_s["info_i"][alive, 0], s["info"]["cur_frame"][alive])
    assert np.array_equal(out_s["info_i"][alive, 1], s["info"]["steps_taken_cur_frame"][alive])
    assert np.array_equal(out_s["info_f"][alive, 3], s["info"]["steps"][alive])
    # (free-running floats are chaotic after a few control steps -- parity_cases.py; the float bounds are asserted by the
    # teacher-forced and 1 / 10 / 100-step cases above)
    assert np.isfinite(out_a["obs"]).all() and np.isfinite(out_a["reward"]).all()
    assert set(np.unique(out_a["done"])) <= {0.0, 1.0}


def test_host_bound_observations():
    """wrap(env).bind_host_obs: the kernel writes the observation rows into page-locked host memory (zero-copy);
    the rows equal the device-resident ones of an identical rollout, bit for bit."""
    import torch
    from brax_tracking_b200 import envs
    m, cfg, clip, _ = common.setup("rodent")
    n = 96
    keys = common.jax_keys(n, seed=4)
    acts = torch.from_numpy(common.actions(3, n, m.nu, seed=6, scale=0.5)).cuda()
    env = envs.wrap(envs.RodentSingleClip(clip, mj_model=m), episode_length=cfg["episode_length"])
    sd, sh = env.reset(keys), env.reset(keys)
    h = env.bind_host_obs(sh)
    assert h.is_pinned() and sh.obs is h
    for t in range(3):
        sd, sh = env.step(sd, acts[t]), env.step(sh, acts[t])
        torch.cuda.synchronize()
        assert np.array_equal(sd.obs.cpu().numpy(), h.numpy())
        assert np.array_equal(sd.reward.cpu().numpy(), sh.reward.cpu().numpy())


def test_two_rodent_model_with_inter_animal_contacts():
    """configs[3]: rodent_pair.xml, 4096-env kernel variant (5 dof slots, 4 contact slots per lane): two kinematic trees with
    contacts BETWEEN them (per-tree reference points in J v / J' f), both animals tracked."""
    from backends import CudaBackend
    name = "rodent_pair"
    m, cfg, clip, tables = common.setup(name)
    assert int(tables["obs_size"][0]) == 1234 and int(tables["ncross"][0]) == 12
    b = CudaBackend(tables)
    pc.check_forward_intermediates(b, name, N=8)
    pc.check_forward_intermediates(b, name, N=16, states=pc.touching_states(name, 16, seed=1))   # animals touching
    pc.check_reset(b, name, N=32)
    print(pc.check_physics_1_10_100(b, name, N=8))
    pc.check_unwrapped_step(b, name, N=8, T=8)
    pc.check_nan_guard(b, name)
    bt = CudaBackend(common.setup(name, 12)[3])
    print(pc.check_teacher_forced(bt, name, N=16, T=30, episode_length=12))


def test_poisoned_scratch_and_scheduling_invariance():
    """compute-sanitizer is closed on this pool (profiles/r2a_sanitizer.txt); what stands in for initcheck / racecheck:
      * `poison`: every program first fills its shared-memory slice with NaN -- a read of anything it did not write itself would
        surface in the outputs; they must be bit-identical to the unpoisoned run;
      * scheduling: warps per CTA (which warps share an SM / a barrier), barrier placement (BT_SYNC), where the constant records are
        read from (staged in shared memory or global) and the position of an environment in the batch change every inter-warp
        timing; outputs must not change by a single bit."""
    import os
    from backends import CudaBackend
    from brax_tracking_b200 import model as model_mod
    for name, n in (("rodent", 96), ("fly_free", 64), ("rodent_pair", 24)):
        m, cfg, clip, tables = common.setup(name, 12)
        keys = common.jax_keys(n, seed=41)
        acts = common.actions(6, n, m.nu, seed=42, scale=0.5)

        def rollout(tb, env=None):
            old = {k: os.environ.get(k) for k in (env or {})}
            os.environ.update(env or {})
            try:
                b = CudaBackend(tb)
            finally:
                for k, v in old.items():
                    os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
            st, out = b.reset(keys)
            first = {k: v.copy() for k, v in st.items()}
            fo, fi = out["obs"].copy(), out["info_i"].copy()
            for t in range(acts.shape[0]):
                b.step(st, out, first, fo, fi, acts[t])
            return st, out

        ref_st, ref_out = rollout(tables)
        variants = {"poison": (dict(tables, poison=np.array([1], np.int32)), None),
                    "sync=1": (dict(tables, sync_mode=np.array([1], np.int32)), None),
                    "sync=0": (dict(tables, sync_mode=np.array([0], np.int32)), None),
                    "sync=64": (dict(tables, sync_mode=np.array([64], np.int32)), None),
                    # every extra alignment point, among the warps of equal parity (incl. the barrier-reduction of the CG loop)
                    "sync=64+1024+2+4+8+16": (dict(tables, sync_mode=np.array([64 + 1024 + 30], np.int32)), None),
                    "sync=1+2+4+8+16": (dict(tables, sync_mode=np.array([31], np.int32)), None),
                    # constant records read from global memory instead of the copy staged in shared memory (BtEnv::crec)
                    "stage=0": (dict(tables, sh_stage_floats=np.array([0], np.int32)), None),
                    "warps=3": (tables, {"BT_WARPS": "3"}), "warps=1": (tables, {"BT_WARPS": "1"})}
        for label, (tb, env) in variants.items():
            st, out = rollout(tb, env)
            for k in ref_st:
                assert np.array_equal(ref_st[k], st[k]), (name, label, k)
            for k in ref_out:
                assert np.array_equal(ref_out[k], out[k], equal_nan=True), (name, label, k)


PPO_KW = dict(learning_rate=3e-4, entropy_cost=1e-3, discounting=0.99, unroll_length=4, batch_size=256, num_minibatches=4,
              num_updates_per_batch=2, normalize_observations=True, num_eval_envs=16)


def test_ppo_loop_runs_on_the_fused_step(tmp_path):
    """configs[4] in miniature: PPO training steps on 256 rodent envs; params move, training + Evaluator metrics finite
    (custom_ppo.py:442-449,484-489), the initial eval is reported at step 0, eval rollouts are deterministic when asked."""
    import torch
    from brax_tracking_b200 import envs, ppo
    m, cfg, clip, _ = common.setup("rodent")
    env = envs.RodentSingleClip(clip, mj_model=m)
    seen = []
    S = 256 * 4 * 4
    mk, params, metrics = ppo.train(env, num_timesteps=2 * S, episode_length=cfg["episode_length"], num_envs=256, num_evals=3,
                                    deterministic_eval=True, progress_fn=lambda s, mt: seen.append((s, mt)), **PPO_KW)
    assert [s for s, _ in seen] == [0, S, 2 * S]                              # initial eval + one per epoch
    assert "eval/episode_reward" in seen[0][1] and "training/sps" not in seen[0][1]
    for k in ("eval/episode_reward", "eval/episode_pos_reward", "eval/avg_episode_length", "eval/sps", "training/sps", "training/v_loss"):
        assert np.isfinite(metrics[k]), k
    assert 1 <= metrics["eval/avg_episode_length"] <= cfg["episode_length"]
    assert all(np.isfinite(v) for v in metrics.values()), metrics
    act = mk(deterministic=True)
    a, raw, logits = act(torch.zeros(3, env.observation_size, device="cuda"))
    assert a.shape == (3, env.action_size) and torch.isfinite(a).all() and float(a.abs().max()) <= 1.0
    assert float(params[0]["count"]) == 2 * S
    # the CUDA-graph replays (unroll, minibatch update) compute what the eager loop computes: same seed, same draws, ONE
    # training step (the rollout runs on the initial policy in both; later steps amplify rounding through the contacts)
    one = {}
    for use_graph in (True, False):
        e = envs.RodentSingleClip(clip, mj_model=m)
        so = {}
        one[use_graph] = ppo.train(e, num_timesteps=S, episode_length=cfg["episode_length"], num_envs=256, num_evals=2, run_evals=False,
                                   use_cuda_graph=use_graph, state_out=so, **PPO_KW)[2]
        one[(use_graph, "p")] = so["training_state"].optimizer.p.clone()
        one[(use_graph, "n")] = float(so["training_state"].optimizer.step_count)
    g_, e_ = one[True], one[False]
    assert abs(g_["training/mean_step_reward"] - e_["training/mean_step_reward"]) < 1e-4 * abs(e_["training/mean_step_reward"]), (g_, e_)
    assert abs(g_["training/v_loss"] - e_["training/v_loss"]) < 0.05 * abs(e_["training/v_loss"]), (g_, e_)
    # graph warm-up no longer applies extra Adam steps to the first minibatch: the first Adam steps move every weight by
    # ~lr, so three extra updates would show as ~1e-3 differences; rounding differences between replay and eager stay far below
    dp = (one[(True, "p")] - one[(False, "p")]).abs()
    assert float(dp.median()) < 1e-5 and one[(True, "n")] == one[(False, "n")] == 8, (float(dp.median()), float(dp.max()))
    # evaluation rollout from frame 0 (main.py:136-258) + device FK for clip preprocessing (preprocess.py:144-204)
    tr = ppo.evaluate_rollout(env, act, common.jax_keys(4, seed=2), num_steps=10)
    assert tr["reward"].shape == (10, 4) and np.isfinite(tr["reward"]).all()
    assert (tr["cur_frame"][1] == 1).all() and (tr["cur_frame"][-1] == 5).all()      # two control steps per mocap frame
    from brax_tracking_b200 import mjcf, preprocess
    q = np.tile(m.qpos0, (6, 1)) + 0.05 * np.random.default_rng(0).standard_normal((6, m.nq))
    c_dev = preprocess.process_clip(q, m, kinematics=preprocess.device_kinematics(env._native))
    c_host = preprocess.process_clip(q, m)
    np.testing.assert_allclose(c_dev.body_positions, c_host.body_positions, atol=2e-6)


def test_checkpoint_save_restore_continue(tmp_path):
    """main.py:136-139,332-333 / custom_ppo.py:411-423 made symmetric: a run that saves after its first epoch, restored into a
    fresh process state and continued, ends where the uninterrupted run ends (parameters, optimiser moments, normaliser, env
    state, step count)."""
    import torch
    from brax_tracking_b200 import envs, ppo
    m, cfg, clip, _ = common.setup("rodent")
    S = 256 * 4 * 4
    kw = dict(episode_length=cfg["episode_length"], num_envs=256, num_evals=3, run_evals=False, **PPO_KW)
    d_a, d_b = str(tmp_path / "a"), str(tmp_path / "b")
    ppo.train(envs.RodentSingleClip(clip, mj_model=m), num_timesteps=2 * S, checkpoint_dir=d_a, **kw)
    import os
    assert sorted(os.listdir(d_a)) == [f"{S}.pt", f"{2 * S}.pt"]
    ppo.train(envs.RodentSingleClip(clip, mj_model=m), num_timesteps=2 * S, checkpoint_dir=d_b, restore_checkpoint_path=os.path.join(d_a, f"{S}.pt"), **kw)
    assert os.listdir(d_b) == [f"{2 * S}.pt"]                                 # continued at S, not restarted at 0
    A = torch.load(os.path.join(d_a, f"{2 * S}.pt"), weights_only=True)
    Bc = torch.load(os.path.join(d_b, f"{2 * S}.pt"), weights_only=True)
    assert A["env_steps"] == Bc["env_steps"] == 2 * S and float(A["optimizer"]["step"]) == float(Bc["optimizer"]["step"]) == 16
    assert torch.equal(A["normalizer"]["count"], Bc["normalizer"]["count"])
    for grp in ("policy", "value"):
        for k in A[grp]:
            torch.testing.assert_close(A[grp][k], Bc[grp][k], rtol=0, atol=2e-6, msg=f"{grp}.{k}")
    torch.testing.assert_close(A["optimizer"]["m"], Bc["optimizer"]["m"], rtol=0, atol=1e-6)
    torch.testing.assert_close(A["normalizer"]["mean"], Bc["normalizer"]["mean"], rtol=0, atol=1e-6)
    assert torch.equal(A["env_raw"]["info_i"], Bc["env_raw"]["info_i"])      # same frames / counters in every environment
    torch.testing.assert_close(A["env_state"]["qpos"], Bc["env_state"]["qpos"], rtol=0, atol=1e-4)


def test_fused_tanh_normal_terms_match_the_torch_formulas():
    """csrc/bt_ppo.cu (log-prob + entropy term of brax's NormalTanhDistribution, forward and backward) against the
    elementwise torch restatement that test_ppo.py checks on the CPU; fp32 tolerance 1e-4 relative."""
    import torch
    from brax_tracking_b200 import ppo
    torch.manual_seed(0)
    B, T, A = 37, 5, 38
    logits = (torch.randn(B, T, 2 * A, device="cuda") * 1.5).requires_grad_(True)
    raw = torch.randn(B, T, A, device="cuda") * 2
    noise = torch.randn(T, B, A, device="cuda")                              # time-major, as the learner draws it
    wl, we = torch.randn(T, B, device="cuda"), torch.randn(T, B, device="cuda")
    lp, ent = ppo._TanhNormalTerms.apply(logits, raw, noise.transpose(0, 1))
    ((lp * wl).sum() + (ent * we).sum()).backward()
    g_fused, logits.grad = logits.grad.clone(), None
    lt = logits.transpose(0, 1)
    lp_ref, ent_ref = ppo.NormalTanh.log_prob(lt, raw.transpose(0, 1)), ppo.NormalTanh.entropy(lt, noise)
    ((lp_ref * wl).sum() + (ent_ref * we).sum()).backward()
    assert lp.shape == (T, B) and ent.shape == (T, B)
    torch.testing.assert_close(lp, lp_ref, rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(ent, ent_ref, rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(g_fused, logits.grad, rtol=1e-4, atol=1e-4)


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()
