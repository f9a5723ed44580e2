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


def test_physics_1_10_100(rodent_cuda):
    print(pc.check_physics_1_10_100(rodent_cuda, "rodent", N=8))


@pytest.mark.parametrize("name", ["fly_free", "fly_tethered"])
def test_fly_elliptic_cone(name):
    """configs[2]: fruit-fly imitation env, 8192-env kernel variant (2 dof slots, 1 contact slot per lane)."""
    from backends import CudaBackend
    b = CudaBackend(common.setup(name)[3])
    pc.check_forward_intermediates(b, name, N=16)
    pc.check_reset(b, name, N=64)
    pc.check_physics_1_10_100(b, name, N=8)
    bt = CudaBackend(common.setup(name, 12)[3])
    print(pc.check_teacher_forced(bt, name, N=16, T=40, episode_length=12))


@pytest.mark.parametrize("name", ["rodent", "fly_free", "fly_tethered"])
def test_cuda_matches_golden(name):
    from backends import CudaBackend
    from test_golden import run_against_golden
    run_against_golden(CudaBackend, name)


def test_two_rodent_stress_model():
    """configs[3]: 4096-env kernel variant (5 dof slots, 4 contact slots per lane)."""
    from backends import CudaBackend
    name = "rodent_pair"
    b = CudaBackend(common.setup(name)[3])
    pc.check_forward_intermediates(b, name, N=8)
    pc.check_reset(b, name, N=32)
    pc.check_physics_1_10_100(b, name, N=4)
    bt = CudaBackend(common.setup(name, 12)[3])
    print(pc.check_teacher_forced(bt, name, N=8, T=30, episode_length=12))


def test_ppo_loop_runs_on_the_fused_step(tmp_path):
    """configs[4] in miniature: two PPO training steps on 256 rodent envs; params move, metrics finite, checkpoint round-trips."""
    import torch
    from brax_tracking_b200 import envs, ppo
    m, cfg, clip, _ = common.setup("rodent")
    env = envs.RodentSingleClip(clip, mj_model=m)
    seen = []
    mk, params, metrics = ppo.train(env, num_timesteps=2 * 256 * 4 * 4, episode_length=cfg["episode_length"], num_envs=256, num_evals=2,
                                    learning_rate=3e-4, entropy_cost=1e-3, discounting=0.99, unroll_length=4, batch_size=256,
                                    num_minibatches=4, num_updates_per_batch=2, normalize_observations=True,
                                    progress_fn=lambda s, mt: seen.append((s, mt)))
    assert seen and seen[-1][0] >= 2 * 256 * 4 * 4
    assert all(np.isfinite(v) for v in metrics.values()), metrics
    act = mk(deterministic=True)
    a, raw, logits = act(torch.zeros(3, env.observation_size, device="cuda"))
    assert a.shape == (3, env.action_size) and torch.isfinite(a).all() and float(a.abs().max()) <= 1.0
    assert float(params[0]["count"]) == 2 * 256 * 4 * 4
    # evaluation rollout from frame 0 (main.py:136-258) + device FK for clip preprocessing (preprocess.py:144-204)
    tr = ppo.evaluate_rollout(env, act, common.jax_keys(4, seed=2), num_steps=10)
    assert tr["reward"].shape == (10, 4) and np.isfinite(tr["reward"]).all()
    assert (tr["cur_frame"][1] == 1).all() and (tr["cur_frame"][-1] == 5).all()      # two control steps per mocap frame
    from brax_tracking_b200 import mjcf, preprocess
    q = np.tile(m.qpos0, (6, 1)) + 0.05 * np.random.default_rng(0).standard_normal((6, m.nq))
    c_dev = preprocess.process_clip(q, m, kinematics=preprocess.device_kinematics(env._native))
    c_host = preprocess.process_clip(q, m)
    np.testing.assert_allclose(c_dev.body_positions, c_host.body_positions, atol=2e-6)


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()
