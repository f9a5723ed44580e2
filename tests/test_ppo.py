"""PPO learner pieces against small numpy restatements of the Brax formulas
(SURVEY Appendix B.5); the GPU smoke run of the whole loop is in tests/test_gpu_parity.py."""
import math

import numpy as np
import torch

from brax_tracking_b200 import ppo


def test_compute_gae_matches_numpy_restatement():
    rng = np.random.default_rng(0)
    T, B = 7, 5
    trunc = (rng.random((T, B)) < 0.15).astype(np.float32)
    term = ((rng.random((T, B)) < 0.15) * (1 - trunc)).astype(np.float32)
    rew = rng.standard_normal((T, B)).astype(np.float32)
    val = rng.standard_normal((T, B)).astype(np.float32)
    boot = rng.standard_normal(B).astype(np.float32)
    lam, disc = 0.95, 0.99
    vs, adv = ppo.compute_gae(*(torch.from_numpy(x) for x in (trunc, term, rew, val, boot)), lam, disc)
    tm = 1 - trunc
    vtp1 = np.concatenate([val[1:], boot[None]])
    deltas = (rew + disc * (1 - term) * vtp1 - val) * tm
    acc = np.zeros(B, np.float32); out = np.zeros((T, B), np.float32)
    for t in range(T - 1, -1, -1):
        acc = deltas[t] + disc * (1 - term[t]) * tm[t] * lam * acc
        out[t] = acc
    vs_ref = out + val
    adv_ref = (rew + disc * (1 - term) * np.concatenate([vs_ref[1:], boot[None]]) - val) * tm
    np.testing.assert_allclose(vs.numpy(), vs_ref, atol=1e-5)
    np.testing.assert_allclose(adv.numpy(), adv_ref, atol=1e-5)


def test_normal_tanh_log_prob_and_entropy():
    torch.manual_seed(0)
    logits = torch.randn(4, 6)
    noise = torch.randn(4, 3)
    raw = ppo.NormalTanh.sample_raw(logits, noise)
    loc, scale = logits[:, :3], torch.nn.functional.softplus(logits[:, 3:]) + 0.001
    a = torch.tanh(raw)
    normal_lp = torch.distributions.Normal(loc, scale).log_prob(raw)
    want = (normal_lp - torch.log(1 - a ** 2 + 1e-12)).sum(-1)           # change of variables for tanh
    np.testing.assert_allclose(ppo.NormalTanh.log_prob(logits, raw).numpy(), want.numpy(), atol=1e-4)
    ent = ppo.NormalTanh.entropy(logits, noise)
    want_e = (torch.distributions.Normal(loc, scale).entropy() + torch.log(1 - a ** 2 + 1e-12)).sum(-1)
    np.testing.assert_allclose(ent.numpy(), want_e.numpy(), atol=1e-4)


def test_running_statistics_matches_numpy():
    rs = ppo.RunningStatistics(3, torch.device("cpu"))
    rng = np.random.default_rng(1)
    chunks = [rng.standard_normal((11, 4, 3)).astype(np.float32) * 3 + 1 for _ in range(3)]
    for c in chunks:
        rs.update(torch.from_numpy(c))
    allx = np.concatenate([c.reshape(-1, 3) for c in chunks])
    np.testing.assert_allclose(rs.mean.numpy(), allx.mean(0), atol=1e-5)
    np.testing.assert_allclose(rs.std.numpy(), allx.std(0), rtol=1e-4)
    assert float(rs.count) == allx.shape[0]


def test_ppo_loss_is_finite_and_has_gradients():
    torch.manual_seed(0)
    B, T, O, nu = 6, 5, 9, 2
    pol, val = ppo.MLP([O, 16, 2 * nu]), ppo.MLP([O, 16, 1])
    w = pol.layers[0].weight
    assert abs(float(w.abs().max())) <= math.sqrt(3.0 / O) + 1e-6 and float(pol.layers[0].bias.abs().max()) == 0.0
    data = dict(observation=torch.randn(B, T, O), next_observation=torch.randn(B, T, O), raw_action=torch.randn(B, T, nu),
                log_prob=torch.randn(B, T) - 2, reward=torch.randn(B, T), discount=torch.ones(B, T), truncation=torch.zeros(B, T))
    loss, m = ppo.compute_ppo_loss(pol, val, lambda x: x, data, torch.randn(T, B, nu))
    loss.backward()
    assert torch.isfinite(loss) and all(p.grad is not None and torch.isfinite(p.grad).all() for p in pol.parameters())
    assert set(m) == {"total_loss", "policy_loss", "v_loss", "entropy_loss"}


def test_linear_with_matvec_bias_gradient_matches_autograd():
    torch.manual_seed(1)
    net = ppo.MLP([9, 16, 4], in_align=4)
    x = torch.randn(5, 7, 9)
    tgt = torch.randn(5, 7, 4)
    _check_linear_grads(net, x, tgt, 1e-5)
    net2 = ppo.MLP([9, 16, 8], in_align=4)                                   # enough rows for the chunked weight-gradient path
    _check_linear_grads(net2, torch.randn(16384, 9), torch.randn(16384, 8), 2e-4)


def _check_linear_grads(net, x, tgt, tol):
    ((net(x) - tgt) ** 2).sum().backward()
    got = [p.grad.clone() for p in net.parameters()]
    for p in net.parameters():
        p.grad = None
    h = torch.nn.functional.pad(x, (0, net.in_padded - x.shape[-1]))
    for i, l in enumerate(net.layers):                                       # the same network through nn.Linear's own backward
        h = l(h)
        if i + 1 < len(net.layers):
            h = torch.nn.functional.silu(h)
    ((h - tgt) ** 2).sum().backward()
    for g, p in zip(got, net.parameters()):
        np.testing.assert_allclose(g.numpy(), p.grad.numpy(), rtol=tol, atol=tol * float(p.grad.abs().max()))


def test_flat_adam_matches_torch_adam_and_keeps_parameters_as_views():
    """FlatAdam (optax.adam semantics, custom_ppo.py:233) on the CPU path against torch.optim.Adam; parameters and gradients are
    views of the two flat buffers, so a module's load_state_dict / backward keep writing into them."""
    torch.manual_seed(0)
    net_a, net_b = ppo.MLP([5, 8, 3]), ppo.MLP([5, 8, 3])
    net_b.load_state_dict(net_a.state_dict())
    params = list(net_a.parameters())
    flat_p, flat_g = ppo._bind_flat(params)
    opt_a = ppo.FlatAdam(flat_p, flat_g, 3e-3)
    opt_b = torch.optim.Adam(net_b.parameters(), lr=3e-3, eps=1e-8)
    x, y = torch.randn(16, 5), torch.randn(16, 3)
    for _ in range(5):
        flat_g.zero_()
        (2.0 * ((net_a(x) - y) ** 2).mean()).backward()      # "summed over 2 ranks" gradient ...
        opt_a.step(0.5)                                       # ... averaged inside the update
        opt_b.zero_grad()
        ((net_b(x) - y) ** 2).mean().backward()
        opt_b.step()
    for pa, pb in zip(net_a.parameters(), net_b.parameters()):
        np.testing.assert_allclose(pa.detach().numpy(), pb.detach().numpy(), atol=2e-6)
        assert pa.data.data_ptr() >= flat_p.data_ptr() and pa.data.data_ptr() < flat_p.data_ptr() + 4 * flat_p.numel()
    assert float(opt_a.step_count) == 5


def test_packed_running_statistics_update_equals_the_three_psum_form():
    """One packed all-reduce (sums about the old mean) == brax's count / mean-update / variance-update psums."""
    rng = np.random.default_rng(3)
    rs = ppo.RunningStatistics(4, torch.device("cpu"))
    mean = np.zeros(4); sv = np.zeros(4); count = 0.0
    for _ in range(4):
        b = (rng.standard_normal((50, 4)) * [1, 5, 0.1, 2] + [0, 10, -3, 100]).astype(np.float32)
        rs.update(torch.from_numpy(b))
        n = count + b.shape[0]                                # brax running_statistics.update
        d_old = b - mean
        mean = mean + d_old.sum(0) / n
        sv = sv + (d_old * (b - mean)).sum(0)
        count = n
    np.testing.assert_allclose(rs.mean.numpy(), mean, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(rs.summed_variance.numpy(), sv, rtol=1e-4)
