"""N > 1 path on CPU: two `gloo` ranks each step their own env shard (host emulation of the kernels); gathering the
shards reproduces the single-process result bit for bit -- the step has no cross-environment communication
(custom_ppo.py:199,213-223)."""
import os
import sys

import numpy as np
import pytest

import common
from brax_tracking_b200 import parallel, prng

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_keys_partition_the_split():
    key = prng.PRNGKey(3)
    full = prng.split(key, 16)
    got = np.concatenate([parallel.shard_keys(key, 16, r, 4) for r in range(4)])
    assert np.array_equal(full, got)
    assert np.array_equal(full, common.jax_keys(16, seed=3))
    with pytest.raises(ValueError):
        parallel.shard_bounds(10, 0, 4)


def _flat_grad_of_mean_loss(rank, world):
    """Gradient of a mean loss over rank `rank`'s contiguous share of a fixed batch, through ppo._bind_flat
    (every parameter / .grad a view of one flat buffer each; autograd accumulates into the gradient views in place)."""
    import torch
    from brax_tracking_b200 import ppo
    torch.manual_seed(5)
    net = ppo.MLP([9, 8, 3], in_align=4)
    x, y = torch.randn(12, 9), torch.randn(12, 3)
    params = list(net.parameters())
    _, flat = ppo._bind_flat(params)
    ptrs = [p.grad.data_ptr() for p in params]
    lo, hi = parallel.shard_bounds(12, rank, world)
    for _ in range(2):                                                       # a second pass reuses the same views
        flat.zero_()
        ((net(x[lo:hi]) - y[lo:hi]) ** 2).mean().backward()
    assert [p.grad.data_ptr() for p in params] == ptrs and float(flat.abs().sum()) > 0
    return flat


def _worker(rank, world, port, n_envs, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    from backends import EmuBackend
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m, cfg, clip, tables = common.setup("rodent")
    b = EmuBackend(tables)
    keys = parallel.shard_keys(prng.PRNGKey(0), n_envs, rank, world)
    lo, hi = parallel.shard_bounds(n_envs, rank, world)
    acts = common.actions(3, n_envs, m.nu, seed=9, scale=0.3)[:, lo:hi]
    st, out = b.reset(keys)
    first = {k: v.copy() for k, v in st.items()}
    fo, fi = out["obs"].copy(), out["info_i"].copy()
    for t in range(3):
        b.step(st, out, first, fo, fi, acts[t])
    mine = torch.from_numpy(np.concatenate([st["qpos"], out["obs"], out["reward"][:, None], out["done"][:, None]], 1))
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    # the only collective of the whole job belongs to the learner: emulate the PPO gradient mean (custom_ppo.py:246-257)
    # with the learner's own flat gradient buffer: each rank back-propagates its half of a batch, one in-place all-reduce
    grad = _flat_grad_of_mean_loss(rank, world)
    dist.all_reduce(grad)
    grad /= world
    # ... and the running statistics of the observations (custom_ppo.py:323-327): ONE packed all-reduce per update
    from brax_tracking_b200 import ppo
    rs = ppo.RunningStatistics(5, torch.device("cpu"))
    gen = torch.Generator().manual_seed(11)
    allx = torch.randn(3, 12, 5, generator=gen) * torch.tensor([1.0, 4.0, 0.2, 2.0, 9.0]) + torch.tensor([0.0, 3.0, -1.0, 50.0, 0.5])
    lo2, hi2 = parallel.shard_bounds(12, rank, world)
    for chunk in allx:
        rs.update(chunk[lo2:hi2], world)
    if rank == 0:
        q.put((torch.cat(gathered).numpy(), grad.numpy(), rs.mean.numpy(), rs.std.numpy(), float(rs.count)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process():
    import torch.multiprocessing as mp
    from backends import EmuBackend
    n_envs, world = 8, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_envs, q)) for r in range(world)]
    for p in procs:
        p.start()
    import queue
    res = None
    for _ in range(300):                                                     # a worker that died must fail the test at once
        try:
            res = q.get(timeout=1)
            break
        except queue.Empty:
            assert all(p.is_alive() or p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert res is not None, "the ranks never reported"
    got, grad, rs_mean, rs_std, rs_count = res
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m, cfg, clip, tables = common.setup("rodent")
    b = EmuBackend(tables)
    st, out = b.reset(common.jax_keys(n_envs))
    first = {k: v.copy() for k, v in st.items()}
    fo, fi = out["obs"].copy(), out["info_i"].copy()
    acts = common.actions(3, n_envs, m.nu, seed=9, scale=0.3)
    for t in range(3):
        b.step(st, out, first, fo, fi, acts[t])
    want = np.concatenate([st["qpos"], out["obs"], out["reward"][:, None], out["done"][:, None]], 1)
    assert np.array_equal(got, want)
    np.testing.assert_allclose(grad, _flat_grad_of_mean_loss(0, 1).numpy(), atol=1e-6)   # pmean of the shard gradients
    import torch
    gen = torch.Generator().manual_seed(11)
    allx = (torch.randn(3, 12, 5, generator=gen) * torch.tensor([1.0, 4.0, 0.2, 2.0, 9.0]) + torch.tensor([0.0, 3.0, -1.0, 50.0, 0.5])).numpy()
    flat = allx.reshape(-1, 5)
    assert rs_count == flat.shape[0]
    np.testing.assert_allclose(rs_mean, flat.mean(0), rtol=1e-5, atol=1e-5)          # statistics of the GLOBAL batch on every rank
    np.testing.assert_allclose(rs_std, flat.std(0), rtol=1e-4)
