"""PPO rollout + learner loop around the fused step.

Mirrors ``train()`` of /root/reference/custom_brax/custom_ppo.py:65-506 (a copy of Brax PPO that swaps in
``custom_wrappers.wrap``) and the Brax pieces it calls: ``acting.generate_unroll``,
``running_statistics``, ``NormalTanhDistribution``, ``compute_gae`` / ``compute_ppo_loss``, ``optax.adam``.
Same argument names and meaning as the reference; differences that come with the platform:

  * one process per GPU (torchrun) instead of ``jax.pmap``; the env shard of a rank is
    ``split(key_env, num_envs)[rank * n : (rank + 1) * n]`` (custom_ppo.py:213-223);
  * ``lax.pmean`` of the gradients (custom_ppo.py:246-248) = ONE NCCL all-reduce of a flat, pre-packed gradient buffer
    per minibatch; the running-statistics sums (custom_ppo.py:323-327) are all-reduced once per training step;
  * the policy / value MLPs run in PyTorch (cuBLAS) -- the north star keeps them out of the hand-written kernels.
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass
from typing import Callable, Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import envs as envs_mod, native, parallel, prng


# ---------------------------------------------------------------------------------------------- networks
class _Linear(torch.autograd.Function):
    """y = x W' + b with the bias gradient taken as a matrix-vector product (ones' dY) instead of a column reduction:
    over the learner's [131 k, 256] activations the reduction kernel runs at a fifth of the memory bandwidth."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        return F.linear(x, w, b)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gy2, x2 = gy.reshape(-1, gy.shape[-1]), x.reshape(-1, x.shape[-1])
        gx = (gy2 @ w).view(x.shape) if ctx.needs_input_grad[0] else None
        n = gy2.shape[0]
        if n >= 16384 and n % 16 == 0 and gy2.shape[1] >= 8 and gy2.is_contiguous() and x2.is_contiguous():
            # weight gradient over K = n rows as 16 partial products + a sum: the batched kernel fills the GPU, the plain
            # `nt` GEMM of this shape runs 2.5x slower (tools/dw_gemm_probe.py: 95 vs 247 us for 256 x 640 at n = 131 k)
            gw = torch.bmm(gy2.view(16, n // 16, -1).transpose(1, 2), x2.view(16, n // 16, -1)).sum(0)
        else:
            gw = gy2.t() @ x2
        gb = torch.mv(gy2.t(), torch.ones(gy2.shape[0], dtype=gy2.dtype, device=gy2.device))
        return gx, gw, gb


class MLP(nn.Module):
    """brax.training.networks.MLP: swish activations, LeCun-uniform kernels, zero biases, linear output layer.

    ``in_align``: the first layer's fan-in is stored padded to a multiple of it (zero columns: they see zero inputs, get zero
    gradients and stay zero), so that the widest GEMMs of the learner (617 -> 256 for the rodent) run on 16-byte aligned
    rows; inputs of the logical width are padded on the fly, inputs that are already padded are used as they are."""

    def __init__(self, sizes: Sequence[int], in_align: int = 1):
        super().__init__()
        self.in_features = sizes[0]
        self.in_padded = -(-sizes[0] // in_align) * in_align
        dims = [self.in_padded, *sizes[1:]]
        self.layers = nn.ModuleList([nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])])
        for i, l in enumerate(self.layers):
            fan_in = sizes[0] if i == 0 else l.in_features
            bound = math.sqrt(3.0 / fan_in)
            nn.init.uniform_(l.weight, -bound, bound)
            nn.init.zeros_(l.bias)
        if self.in_padded != self.in_features:
            with torch.no_grad():
                self.layers[0].weight[:, self.in_features:].zero_()

    def forward(self, x):
        if x.shape[-1] != self.in_padded:
            x = F.pad(x, (0, self.in_padded - x.shape[-1]))
        for i, l in enumerate(self.layers):
            x = _Linear.apply(x, l.weight, l.bias) if torch.is_grad_enabled() and l.weight.requires_grad else l(x)
            if i + 1 < len(self.layers):
                x = F.silu(x)
        return x


class RunningStatistics:
    """brax.training.acme.running_statistics: count / mean / summed_variance / std (clipped to [1e-6, 1e6])."""

    def __init__(self, size: int, device):
        self.count = torch.zeros((), dtype=torch.float64, device=device)
        self.mean = torch.zeros(size, dtype=torch.float32, device=device)
        self.summed_variance = torch.zeros(size, dtype=torch.float32, device=device)
        self.std = torch.ones(size, dtype=torch.float32, device=device)

    def update(self, batch: torch.Tensor, world: int = 1):
        """Parallel (Chan) update with the batch of every rank (pmap_axis_name='i' in the reference, custom_ppo.py:323-327).
        The three psums of the reference (count, sum of deviations, sum of squared deviations about the NEW mean) are taken
        about the OLD mean instead, which every rank already has, so that they travel as ONE packed all-reduce of 2 D + 1 floats:
        sum (b - mean')(b - mean) = S2 - (S1 / N') S1 with S1 = sum (b - mean), S2 = sum (b - mean)^2."""
        b = batch.reshape(-1, batch.shape[-1]).to(torch.float32)
        D = b.shape[1]
        diff = b - self.mean
        packed = torch.empty(2 * D + 1, dtype=torch.float64, device=b.device)
        packed[:D] = diff.sum(0, dtype=torch.float64)
        packed[D:2 * D] = (diff * diff).sum(0, dtype=torch.float64)
        packed[2 * D] = float(b.shape[0])
        if world > 1:
            dist.all_reduce(packed)
        s1, s2, n = packed[:D], packed[D:2 * D], packed[2 * D]
        new_count = self.count + n
        delta = s1 / new_count
        # in place: captured CUDA graphs (rollout policy) keep reading these tensors
        self.summed_variance.add_((s2 - delta * s1).to(torch.float32))
        self.mean.add_(delta.to(torch.float32))
        self.count.copy_(new_count)
        self.std.copy_(torch.sqrt(torch.clamp(self.summed_variance / new_count.to(torch.float32), min=0.0)).clamp(1e-6, 1e6))

    def normalize(self, x):
        return (x - self.mean) / self.std

    def state_dict(self):
        return dict(count=self.count, mean=self.mean, summed_variance=self.summed_variance, std=self.std)

    def load_state_dict(self, d):
        for k in ("count", "mean", "summed_variance", "std"):
            getattr(self, k).copy_(d[k])


class NormalTanh:
    """brax NormalTanhDistribution (min_std = 0.001): logits -> (loc, softplus(scale) + min_std), tanh bijector."""
    MIN_STD = 0.001

    @staticmethod
    def params(logits):
        loc, scale = torch.chunk(logits, 2, dim=-1)
        return loc, F.softplus(scale) + NormalTanh.MIN_STD

    @staticmethod
    def log_det_jac(x):
        return 2.0 * (math.log(2.0) - x - F.softplus(-2.0 * x))

    @staticmethod
    def sample_raw(logits, noise):
        loc, scale = NormalTanh.params(logits)
        return loc + scale * noise

    @staticmethod
    def log_prob(logits, raw):
        loc, scale = NormalTanh.params(logits)
        lp = -0.5 * ((raw - loc) / scale) ** 2 - torch.log(scale) - 0.5 * math.log(2.0 * math.pi)
        return (lp - NormalTanh.log_det_jac(raw)).sum(-1)

    @staticmethod
    def entropy(logits, noise):
        loc, scale = NormalTanh.params(logits)
        ent = 0.5 + 0.5 * math.log(2.0 * math.pi) + torch.log(scale)
        return (ent + NormalTanh.log_det_jac(loc + scale * noise)).sum(-1)


class _TanhNormalTerms(torch.autograd.Function):
    """(log_prob(raw), sampled-entropy term) of NormalTanh per row, time-major [T, B], from batch-major logits [B, T, 2A]:
    one fused CUDA pass forward and one backward (csrc/bt_ppo.cu) instead of ~60 elementwise launches."""

    @staticmethod
    def forward(ctx, logits, raw, noise):
        ctx.save_for_backward(logits, raw, noise)
        return native.ppo_tanh_normal(logits, raw, noise)

    @staticmethod
    def backward(ctx, glp, gent):
        logits, raw, noise = ctx.saved_tensors
        return native.ppo_tanh_normal(logits, raw, noise, grads=(glp.contiguous(), gent.contiguous())), None, None


# ---------------------------------------------------------------------------------------------- losses
def compute_gae(truncation, termination, rewards, values, bootstrap_value, lambda_: float, discount: float):
    """brax.training.agents.ppo.losses.compute_gae; time-major [T, B]."""
    with torch.no_grad():                                                    # targets only: nothing here is differentiated
        trunc_mask = 1.0 - truncation
        values_tp1 = torch.cat([values[1:], bootstrap_value[None]], 0)
        deltas = (rewards + discount * (1.0 - termination) * values_tp1 - values) * trunc_mask
        coef = discount * (1.0 - termination) * trunc_mask * lambda_         # one fused multiply-add per step of the scan
        acc = torch.zeros_like(bootstrap_value)
        out = []
        for t in range(values.shape[0] - 1, -1, -1):
            acc = torch.addcmul(deltas[t], coef[t], acc)
            out.append(acc)
        vs = torch.stack(out[::-1], 0) + values
        vs_tp1 = torch.cat([vs[1:], bootstrap_value[None]], 0)
        adv = (rewards + discount * (1.0 - termination) * vs_tp1 - values) * trunc_mask
    return vs.detach(), adv.detach()


def compute_ppo_loss(policy, value, normalizer, data: Dict[str, torch.Tensor], noise, entropy_cost=1e-4, discounting=0.9,
                     reward_scaling=1.0, gae_lambda=0.95, clipping_epsilon=0.3, normalize_advantage=True):
    """brax compute_ppo_loss; ``data`` is batch-major [B, T, ...] as stored by the rollout."""
    tm = {k: v.transpose(0, 1) for k, v in data.items() if "observation" not in k}  # time first
    # the networks act row by row: they are applied to the observation rows as stored (batch-major, contiguous: no transposed
    # copy of the widest tensor of the update) and only their narrow outputs are viewed time-first
    obs = normalizer(data["observation"])
    logits_bm = policy(obs)
    logits = logits_bm.transpose(0, 1)
    baseline = value(obs).squeeze(-1).transpose(0, 1)
    bootstrap = value(normalizer(data["next_observation"][:, -1])).squeeze(-1)
    rewards = tm["reward"] * reward_scaling
    truncation = tm["truncation"]
    termination = (1.0 - tm["discount"]) * (1.0 - truncation)
    fused = logits_bm.is_cuda and logits_bm.dtype == torch.float32 and logits_bm.dim() == 3
    if fused:
        target_lp, entropy_rows = _TanhNormalTerms.apply(logits_bm.contiguous(), data["raw_action"], noise.transpose(0, 1))
    else:
        target_lp, entropy_rows = NormalTanh.log_prob(logits, tm["raw_action"]), NormalTanh.entropy(logits, noise)
    vs, adv = compute_gae(truncation, termination, rewards, baseline, bootstrap, gae_lambda, discounting)
    if normalize_advantage:
        adv = (adv - adv.mean()) / (adv.std(unbiased=False) + 1e-8)
    rho = torch.exp(target_lp - tm["log_prob"])
    policy_loss = -torch.minimum(rho * adv, rho.clamp(1 - clipping_epsilon, 1 + clipping_epsilon) * adv).mean()
    v_err = vs - baseline
    v_loss = (v_err * v_err).mean() * 0.5 * 0.5
    entropy = entropy_rows.mean()
    total = policy_loss + v_loss - entropy_cost * entropy
    return total, dict(total_loss=total.detach(), policy_loss=policy_loss.detach(), v_loss=v_loss.detach(), entropy_loss=(-entropy_cost * entropy).detach())


# ---------------------------------------------------------------------------------------------- training
@dataclass
class TrainingState:
    """custom_ppo.py:41-48"""
    policy: MLP
    value: MLP
    optimizer: torch.optim.Optimizer
    normalizer: RunningStatistics
    env_steps: int = 0


def _bind_flat(params) -> Tuple[torch.Tensor, torch.Tensor]:
    """Every parameter and every ``.grad`` becomes a view into ONE flat buffer each (autograd accumulates into an existing
    ``.grad`` in place), so that lax.pmean(grads, 'i') is a single NCCL all-reduce with no gather / scatter copies around it and
    optax.adam is a single fused pass over the buffers."""
    n = sum(p.numel() for p in params)
    flat_p = torch.empty(n, dtype=params[0].dtype, device=params[0].device)
    flat_g = torch.zeros(n, dtype=params[0].dtype, device=params[0].device)
    off = 0
    for p in params:
        k = p.numel()
        flat_p[off:off + k].copy_(p.data.reshape(-1))
        p.data = flat_p[off:off + k].view_as(p)
        p.grad = flat_g[off:off + k].view_as(p)
        off += k
    return flat_p, flat_g


class FlatAdam:
    """optax.adam(learning_rate) (custom_ppo.py:233) over the flat parameter buffer: one fused kernel per update
    (csrc/bt_ppo.cu::k_flat_adam) that also applies the 1 / world of the gradient mean; the update count lives on the device so the
    update is capturable in a CUDA graph.  CPU tensors (unit tests) take the same formulas through torch ops."""

    def __init__(self, flat_p, flat_g, lr, b1=0.9, b2=0.999, eps=1e-8):
        self.p, self.g, self.lr, self.b1, self.b2, self.eps = flat_p, flat_g, float(lr), b1, b2, eps
        self.m, self.v = torch.zeros_like(flat_p), torch.zeros_like(flat_p)
        self.step_count = torch.zeros((), dtype=torch.float32, device=flat_p.device)

    def step(self, gscale: float = 1.0):
        if self.p.is_cuda:
            native.ppo_flat_adam(self.p, self.g, self.m, self.v, self.step_count, self.lr, self.b1, self.b2, self.eps, gscale)
        else:
            t = self.step_count + 1.0
            g = self.g * gscale
            self.m.mul_(self.b1).add_(g, alpha=1 - self.b1)
            self.v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            self.p.sub_(self.lr * (self.m / (1 - self.b1 ** t)) / (torch.sqrt(self.v / (1 - self.b2 ** t)) + self.eps))
        self.step_count.add_(1.0)

    def state_dict(self):
        return dict(m=self.m, v=self.v, step=self.step_count)

    def load_state_dict(self, d):
        self.m.copy_(d["m"]); self.v.copy_(d["v"]); self.step_count.copy_(d["step"])


class Evaluator:
    """brax.training.acting.Evaluator over brax's EvalWrapper (custom_ppo.py:442-449,484-489): ``num_eval_envs`` environments run
    ONE episode each (``episode_length // action_repeat`` steps of the wrapped env); per environment the reward and every metric
    are summed while its first episode is active.  Returns ``eval/episode_<name>`` (mean over environments; ``reward`` included),
    ``eval/avg_episode_length``, ``eval/epoch_eval_time``, ``eval/sps``, ``eval/walltime`` merged with the training metrics."""

    def __init__(self, eval_env, make_policy, num_eval_envs: int, episode_length: int, action_repeat: int, key, deterministic: bool = False):
        self._env, self._make_policy, self._n = eval_env, make_policy, int(num_eval_envs)
        self._steps = episode_length // action_repeat
        self._key, self._det, self._walltime = np.asarray(key, dtype=np.uint32), deterministic, 0.0

    def run_evaluation(self, training_metrics: Dict[str, float]) -> Dict[str, float]:
        self._key, unroll_key = prng.split(self._key, 2)
        env, t0 = self._env, time.time()
        state = env.reset(prng.split(unroll_key, self._n))
        act = self._make_policy(deterministic=self._det)
        sums = {k: torch.zeros_like(state.reward) for k in ("reward", *state.metrics)}
        active = torch.ones_like(state.reward)
        ep_steps = torch.zeros_like(state.reward)
        for _ in range(self._steps):
            state = env.step(state, act(state.obs)[0].contiguous())
            ep_steps = torch.where(active > 0, state.info["steps"], ep_steps)
            sums["reward"] += state.reward * active
            for k, v in state.metrics.items():
                sums[k] += v * active
            active = active * (1.0 - state.done)
        out = {f"eval/episode_{k}": float(v.mean()) for k, v in sums.items()}
        out.update({f"eval/episode_{k}_std": float(v.std(unbiased=False)) for k, v in sums.items()})
        dt = time.time() - t0
        self._walltime += dt
        out.update({"eval/avg_episode_length": float(ep_steps.mean()), "eval/epoch_eval_time": dt,
                    "eval/sps": self._steps * self._n / dt, "eval/walltime": self._walltime, **training_metrics})
        return out


def train(environment, num_timesteps: int, episode_length: int, action_repeat: int = 1, num_envs: int = 1, num_eval_envs: int = 128,
          learning_rate: float = 1e-4, entropy_cost: float = 1e-4, discounting: float = 0.9, seed: int = 0, unroll_length: int = 10,
          batch_size: int = 32, num_minibatches: int = 16, num_updates_per_batch: int = 2, num_evals: int = 1,
          num_resets_per_eval: int = 0, normalize_observations: bool = False, reward_scaling: float = 1.0,
          clipping_epsilon: float = 0.3, gae_lambda: float = 0.95, deterministic_eval: bool = False,
          policy_hidden_layer_sizes: Sequence[int] = (256, 256), value_hidden_layer_sizes: Sequence[int] = (256, 256),
          progress_fn: Callable[[int, Dict], None] = lambda *a: None, normalize_advantage: bool = True, eval_env=None,
          policy_params_fn: Callable[..., None] = lambda *a: None, restore_checkpoint_path: Optional[str] = None,
          checkpoint_dir: Optional[str] = None, run_evals: bool = True, matmul_precision: str = "tf32", use_cuda_graph: bool = True,
          state_out: Optional[dict] = None):
    """PPO training on the fused B200 step.  Returns (make_policy, params, metrics) like the reference (custom_ppo.py:65-506).

    Beyond the reference's arguments: ``checkpoint_dir`` -- rank 0 writes ``<dir>/<env_steps>.pt`` (``save_checkpoint``: normaliser,
    policy, value, optimiser, step count, sampling-generator state, environment state) after every evaluation and at the end
    (the reference saves (normaliser, policy) from ``policy_params_fn``, main.py:136-139,332-333); ``restore_checkpoint_path``
    resumes from such a file, continuing at its ``env_steps``; ``run_evals=False`` skips the Evaluator (benchmarks); ``state_out``: a
    dict that receives the final ``TrainingState`` and env ``State`` (tools, tests);
    ``matmul_precision``: "tf32" (default; what XLA's DEFAULT precision gives the reference's f32 MLPs on Ampere and later GPUs)
    or "highest" (plain fp32 matmuls)."""
    torch.backends.cuda.matmul.allow_tf32 = matmul_precision == "tf32"
    torch.backends.cudnn.allow_tf32 = matmul_precision == "tf32"
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    assert batch_size * num_minibatches % num_envs == 0                      # custom_ppo.py:152
    assert num_envs % world == 0                                             # custom_ppo.py:199
    env_step_per_training_step = batch_size * unroll_length * num_minibatches * action_repeat
    num_evals_after_init = max(num_evals - 1, 1)
    # custom_ppo.py:176-183: training steps per epoch, with the extra resets dividing the epoch
    num_training_steps_per_epoch = int(np.ceil(num_timesteps / (num_evals_after_init * env_step_per_training_step
                                                                * max(num_resets_per_eval, 1))))

    env = envs_mod.wrap(environment, episode_length=episode_length, action_repeat=action_repeat)
    device = environment._native._dev()
    n_local = num_envs // world
    # keys: PRNGKey(seed) -> (global, local); local -> (local, key_env, eval_key)   (custom_ppo.py:189-196)
    key = prng.PRNGKey(seed)
    global_key, local_key = prng.split(key, 2)
    local_key, key_env, eval_key = prng.split(local_key, 3)
    key_envs = prng.split(key_env, num_envs)                                 # custom_ppo.py:221
    lo, hi = parallel.shard_bounds(num_envs, rank, world)
    state = env.reset(key_envs[lo:hi])
    gen = torch.Generator(device=device)
    gen.manual_seed(int(global_key[1]) + 7919 * rank)

    obs_size, nu = env.observation_size, env.action_size
    torch.manual_seed(int(global_key[0]) % (2 ** 31))                        # networks identical on every rank
    policy = MLP([obs_size, *policy_hidden_layer_sizes, 2 * nu], in_align=32).to(device)
    value = MLP([obs_size, *value_hidden_layer_sizes, 1], in_align=32).to(device)
    obs_pad = policy.in_padded
    params = list(policy.parameters()) + list(value.parameters())
    flat_param, flat_grad = _bind_flat(params)
    grad_ptrs = [p.grad.data_ptr() for p in params]
    opt = FlatAdam(flat_param, flat_grad, learning_rate, eps=1e-8)
    ts = TrainingState(policy, value, opt, RunningStatistics(obs_size, device))
    if restore_checkpoint_path is not None:
        blob = load_checkpoint(ts, restore_checkpoint_path)
        if "rng_state" in blob and world == 1:
            gen.set_state(blob["rng_state"].cpu())
        if "env_state" in blob and blob["env_state"]["qpos"].shape == state.pipeline_state["qpos"].shape:
            restore_env_state(state, blob)
    norm = ts.normalizer.normalize if normalize_observations else (lambda x: x)

    def make_policy(deterministic: bool = False):
        def act(obs):
            with torch.no_grad():
                logits = policy(norm(obs))
                loc, scale = NormalTanh.params(logits)
                raw = loc if deterministic else loc + scale * torch.randn(loc.shape, device=loc.device, generator=gen)
                return torch.tanh(raw), raw, logits
        return act

    n_unrolls = batch_size * num_minibatches // num_envs
    T = unroll_length
    # rollout buffers, time-major per unroll; only the LAST next_observation of an unroll is ever used (bootstrap value,
    # brax compute_ppo_loss: data.next_observation[-1]), so that is all that is kept
    buf = dict(observation=torch.empty(n_unrolls, T, n_local, obs_size, device=device),
               next_observation=torch.empty(n_unrolls, 1, n_local, obs_size, device=device),
               raw_action=torch.empty(n_unrolls, T, n_local, nu, device=device),
               log_prob=torch.empty(n_unrolls, T, n_local, device=device), reward=torch.empty(n_unrolls, T, n_local, device=device),
               discount=torch.empty(n_unrolls, T, n_local, device=device), truncation=torch.empty(n_unrolls, T, n_local, device=device))
    # learner-side view of the same data, batch-major [n_unrolls * n_local, T, ...] (custom_ppo.py:316-320), in static
    # storage so that the minibatch update can be replayed as one CUDA graph
    B = n_unrolls * n_local
    # (observation rows are stored at the padded width of the networks' first layer, pad columns zero)
    data = {k: torch.zeros(B, v.shape[1], *(v.shape[3:] if "observation" not in k else (obs_pad,)), device=device)
            for k, v in buf.items()}
    mb_size = B // num_minibatches
    mb_idx = torch.zeros(mb_size, dtype=torch.long, device=device)
    mb_noise = torch.zeros(T, mb_size, nu, device=device)
    mb_data = {k: torch.empty(mb_size, *v.shape[1:], device=device) for k, v in data.items()}
    ident = lambda x: x
    graph = {"g": None, "lm": None, "tried": False}

    def minibatch_update():
        """One SGD step on the minibatch selected by mb_idx / mb_noise (custom_ppo.py:250-284)."""
        for k, v in data.items():
            torch.index_select(v, 0, mb_idx, out=mb_data[k])
        # observations were normalised once for the whole batch (the statistics are fixed during the SGD epochs)
        loss, lm = compute_ppo_loss(policy, value, ident, mb_data, mb_noise, entropy_cost, discounting, reward_scaling, gae_lambda,
                                    clipping_epsilon, normalize_advantage)
        flat_grad.zero_()
        loss.backward()
        if world > 1:
            dist.all_reduce(flat_grad)       # lax.pmean(grads, 'i'): ONE all-reduce (sum), in place; the 1 / world rides in the Adam pass
        opt.step(1.0 / world)
        return lm

    def run_minibatch():
        if not use_cuda_graph or device.type != "cuda":
            return minibatch_update()
        if graph["g"] is None and not graph["tried"]:
            graph["tried"] = True
            ok = True
            try:
                # warm-up on a side stream (cuBLAS handles, autograd buffers) with the parameters and the optimiser state put back
                # afterwards: only the replay applies this minibatch's update, as in the eager schedule
                snap = [t.clone() for t in (flat_param, opt.m, opt.v, opt.step_count)]
                side = torch.cuda.Stream(device)
                side.wait_stream(torch.cuda.current_stream(device))
                with torch.cuda.stream(side):
                    for _ in range(3):
                        minibatch_update()
                torch.cuda.current_stream(device).wait_stream(side)
                for t, c in zip((flat_param, opt.m, opt.v, opt.step_count), snap):
                    t.copy_(c)
                assert [p.grad.data_ptr() for p in params] == grad_ptrs, "autograd replaced a flat gradient view"
                g = torch.cuda.CUDAGraph()
                # the NCCL all-reduce is captured with the update (world > 1); thread-local capture mode because the process
                # group's watchdog thread queries events while this thread captures
                with torch.cuda.graph(g, capture_error_mode="thread_local" if world > 1 else "global"):
                    graph["lm"] = minibatch_update()
                graph["g"] = g
            except Exception as e:  # pragma: no cover - capture is an optimisation, never a requirement
                print(f"[ppo] CUDA graph capture of the minibatch update failed ({e!r}); running eagerly", flush=True)
                ok = False
            if world > 1:   # every rank replays or every rank runs eagerly: the all-reduce counts must match
                flag = torch.tensor([1.0 if ok else 0.0], device=device)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                ok = bool(flag.item() > 0)
            if not ok:
                graph["g"] = None
        if graph["g"] is None:
            return minibatch_update()
        graph["g"].replay()
        return graph["lm"]

    # one unroll = T x (policy inference, sampling, fused env step, 7 transition records).  The env step is in place on static
    # buffers and the C ABI only enqueues a kernel on the current stream, so the whole unroll replays as ONE CUDA graph
    # writing a static [T, n_local, ...] staging record; only the sampling noise is drawn outside it.
    ubuf = {k: torch.empty_like(v[0]) for k, v in buf.items()}
    roll_noise = torch.zeros(T, n_local, nu, device=device)
    rgraph = {"g": None, "tried": False}

    def unroll_into(state, dst):
        with torch.no_grad():
            for t in range(T):
                dst["observation"][t].copy_(state.obs)
                logits = policy(norm(state.obs))
                loc, scale = NormalTanh.params(logits)
                raw = loc + scale * roll_noise[t]
                dst["raw_action"][t].copy_(raw)
                dst["log_prob"][t].copy_(NormalTanh.log_prob(logits, raw))
                state = env.step(state, torch.tanh(raw).contiguous())
                if t == T - 1:
                    dst["next_observation"][0].copy_(state.obs)
                dst["reward"][t].copy_(state.reward)
                dst["discount"][t].copy_(1.0 - state.done)
                dst["truncation"][t].copy_(state.info["truncation"])
        return state

    def run_unroll(state, u):
        roll_noise.normal_(generator=gen)
        dst_u = {k: v[u] for k, v in buf.items()}
        if not use_cuda_graph or device.type != "cuda":
            return unroll_into(state, dst_u)
        if rgraph["g"] is None:
            if rgraph["tried"]:
                return unroll_into(state, dst_u)
            rgraph["tried"] = True
            state = unroll_into(state, dst_u)                                 # first unroll eagerly: cuBLAS handles, workspaces
            try:
                torch.cuda.synchronize(device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, capture_error_mode="thread_local" if world > 1 else "global"):
                    unroll_into(state, ubuf)                                  # records only: the env state is not advanced
                rgraph["g"] = g
            except Exception as e:  # pragma: no cover - capture is an optimisation, never a requirement
                print(f"[ppo] CUDA graph capture of the unroll failed ({e!r}); running eagerly", flush=True)
            return state
        rgraph["g"].replay()
        for k, v in dst_u.items():
            v.copy_(ubuf[k])
        return state

    def training_step(state):
        # ---- rollout: acting.generate_unroll x n_unrolls (custom_ppo.py:296-314)
        for u in range(n_unrolls):
            state = run_unroll(state, u)
        # [n_unrolls, T, n, ...] -> [n_unrolls * n, T, ...]  (custom_ppo.py:316-320)
        for k, v in buf.items():
            dst = data[k].view(n_unrolls, n_local, *data[k].shape[1:])
            (dst[..., :obs_size] if "observation" in k else dst).copy_(v.transpose(1, 2))
        if normalize_observations:
            ts.normalizer.update(buf["observation"], world)                  # custom_ppo.py:323-327
            for k in ("observation", "next_observation"):
                data[k][..., :obs_size].sub_(ts.normalizer.mean).div_(ts.normalizer.std)
        # ---- SGD: num_updates_per_batch x num_minibatches (custom_ppo.py:250-284,329-334)
        lm = None
        for _ in range(num_updates_per_batch):
            perm = torch.randperm(B, device=device, generator=gen)
            for mb in perm.reshape(num_minibatches, -1):
                mb_idx.copy_(mb)
                mb_noise.normal_(generator=gen)
                lm = run_minibatch()
        ts.env_steps += env_step_per_training_step
        return state, lm

    def checkpoint(state):
        if checkpoint_dir is not None and rank == 0:
            import os
            os.makedirs(checkpoint_dir, exist_ok=True)
            save_checkpoint(ts, os.path.join(checkpoint_dir, f"{ts.env_steps}.pt"), env_state=state, rng_state=gen.get_state())

    evaluator = None
    if run_evals and rank == 0:
        # custom_ppo.py:430-449: the eval env is the same env under the same wrappers (brax adds its EvalWrapper inside Evaluator)
        ev_base = eval_env if eval_env is not None else environment
        ev_wrapped = ev_base if isinstance(ev_base, envs_mod.AutoResetWrapperTracking) else envs_mod.wrap(ev_base, episode_length=episode_length, action_repeat=action_repeat)
        evaluator = Evaluator(ev_wrapped, make_policy, num_eval_envs, episode_length, action_repeat, eval_key, deterministic_eval)

    metrics: Dict[str, float] = {}
    if evaluator is not None and num_evals > 1 and ts.env_steps == 0:       # initial eval (custom_ppo.py:452-459)
        metrics = evaluator.run_evaluation({})
        progress_fn(0, metrics)
    t_start = time.time()
    ev0, ev1 = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) if device.type == "cuda" else (None, None)
    while ts.env_steps < num_timesteps:                                      # (a restored run continues at its env_steps)
        t0 = time.time()
        lm, n_steps = None, 0
        if ev0 is not None:
            ev0.record()
        for _ in range(max(num_resets_per_eval, 1)):
            for _ in range(num_training_steps_per_epoch):
                state, lm = training_step(state)
                n_steps += 1
            if num_resets_per_eval > 0:                                       # custom_ppo.py:476-480: fresh keys, host-side reset
                key_envs = np.stack([prng.split(k, 2)[0] for k in key_envs])
                state = env.reset(key_envs[lo:hi])
        if ev1 is not None:
            ev1.record()
            torch.cuda.synchronize(device)
            dev_s = ev0.elapsed_time(ev1) / 1e3                               # device time of the epoch ...
            if world > 1:
                tt = torch.tensor([dev_s], dtype=torch.float64, device=device)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)                     # ... max over ranks
                dev_s = float(tt[0])
        else:
            dev_s = time.time() - t0
        dt = time.time() - t0
        sps = n_steps * env_step_per_training_step / dt                       # custom_ppo.py:373-377
        training_metrics = {"training/sps": sps, "training/walltime": time.time() - t_start, "training/device_sps": n_steps * env_step_per_training_step / dev_s,
                            "training/mean_step_reward": float(buf["reward"].mean()), **{f"training/{k}": float(v) for k, v in lm.items()}}
        metrics = evaluator.run_evaluation(training_metrics) if evaluator is not None else training_metrics
        if world > 1:
            dist.barrier()                                                    # the other ranks wait for rank 0's evaluation
        checkpoint(state)
        if rank == 0:
            progress_fn(ts.env_steps, metrics)
            policy_params_fn(ts.env_steps, make_policy, (ts.normalizer.state_dict(), policy.state_dict()))
    if state_out is not None:
        state_out.update(training_state=ts, env_state=state)
    return make_policy, (ts.normalizer.state_dict(), policy.state_dict()), metrics


# ---------------------------------------------------------------------------------------------- checkpoints
_ENV_RAW_KEYS = ("obs", "reward", "done", "metrics", "info_f", "info_i", "first_obs", "first_info_i", "clip_idx")


def save_checkpoint(ts: TrainingState, path: str, env_state=None, rng_state=None) -> None:
    """Symmetric save / restore of (normalizer, policy, value, optimizer, env_steps[, sampling-generator state, env state]) -- the
    reference saves only (normalizer, policy) through ``policy_params_fn`` (main.py:136-139,332-333) and restores through a
    different format (custom_ppo.py:411-423).  Plain tensors / numbers only: loads with ``weights_only=True``."""
    blob = dict(normalizer=ts.normalizer.state_dict(), policy=ts.policy.state_dict(), value=ts.value.state_dict(),
                optimizer=ts.optimizer.state_dict(), env_steps=int(ts.env_steps))
    if rng_state is not None:
        blob["rng_state"] = rng_state.clone()
    if env_state is not None:
        blob["env_state"] = {k: v.clone() for k, v in env_state.pipeline_state.items()}
        blob["env_first"] = {k: v.clone() for k, v in env_state._raw["first"].items()} if "first" in env_state._raw else {}
        blob["env_raw"] = {k: env_state._raw[k].clone() for k in _ENV_RAW_KEYS if k in env_state._raw}
    tmp = path + ".tmp"
    torch.save(blob, tmp)
    import os
    os.replace(tmp, path)                                                     # a killed run never leaves a truncated checkpoint


def load_checkpoint(ts: TrainingState, path: str) -> dict:
    blob = torch.load(path, map_location=ts.normalizer.mean.device, weights_only=True)
    ts.normalizer.load_state_dict(blob["normalizer"])
    ts.policy.load_state_dict(blob["policy"])                                 # (in place: the parameters stay views of the flat buffer)
    ts.value.load_state_dict(blob["value"])
    ts.optimizer.load_state_dict(blob["optimizer"])
    ts.env_steps = int(blob["env_steps"])
    return blob


def restore_env_state(state, blob) -> None:
    """Puts a checkpointed environment state back into a freshly reset ``State`` of the same batch size, in place."""
    for k, v in blob["env_state"].items():
        state.pipeline_state[k].copy_(v)
    for k, v in blob.get("env_first", {}).items():
        state._raw["first"][k].copy_(v)
    for k, v in blob["env_raw"].items():
        state._raw[k].copy_(v)


# ---------------------------------------------------------------------------------------------- evaluation rollout
def evaluate_rollout(environment, policy_fn, rng_keys, num_steps: Optional[int] = None) -> Dict[str, np.ndarray]:
    """The authors' de-facto regression signal (/root/reference/main.py:136-258): a deterministic-policy rollout of
    ``clip_length`` frames from frame 0 through ``RenderRolloutWrapperTracking`` with per-step traces of the reward terms,
    the tracking distances and the torso height.  ``policy_fn(obs) -> (action, ...)`` as returned by ``make_policy``.
    Returns {name: [num_steps, n_envs]} on the host (rendering itself is out of scope: it needs MuJoCo's GL renderer)."""
    env = envs_mod.RenderRolloutWrapperTracking(environment)
    state = env.reset(rng_keys)
    if num_steps is None:
        num_steps = int(250 * environment._steps_for_cur_frame)              # main.py:147
    names = ["pos_reward", "quat_reward", "joint_reward", "bodypos_reward", "endeff_reward"]
    trace = {k: [] for k in names + ["reward", "summed_pos_distance", "joint_distance", "torso_height", "done", "cur_frame"]}
    tz = 3 * (environment._thorax_idx % environment.sys.nbody) + 2
    for _ in range(num_steps):
        action = policy_fn(state.obs)[0]
        state = env.step(state, action.contiguous())
        for k in names:
            trace[k].append(state.metrics[k].clone())
        trace["reward"].append(state.reward.clone())
        trace["summed_pos_distance"].append(state.info["summed_pos_distance"].clone())
        trace["joint_distance"].append(state.info["joint_distance"].clone())
        trace["torso_height"].append(state.pipeline_state["xpos"][:, tz].clone())
        trace["done"].append(state.done.clone())
        trace["cur_frame"].append(state.info["cur_frame"].clone())
    return {k: torch.stack(v).cpu().numpy() for k, v in trace.items()}
