"""Kernel-time breakdown of one PPO training step (rollout + minibatch updates) with torch.profiler.
usage: python tools/profile_learner.py [num_envs]   (needs a GPU)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from brax_tracking_b200 import ppo, presets

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
env = presets.make_env("rodent", device=0)
kw = dict(episode_length=env.episode_length, num_envs=n, num_evals=2, learning_rate=3e-4, entropy_cost=1e-3, discounting=0.99,
          unroll_length=16, batch_size=n, num_minibatches=32, num_updates_per_batch=2, normalize_observations=True)
eager = len(sys.argv) > 2 and sys.argv[2] == "ops"   # "ops": eager loop, so that kernels are attributed to their aten ops
with profile(activities=[ProfilerActivity.CUDA] + ([ProfilerActivity.CPU] if eager else [])) as prof:
    ppo.train(env, num_timesteps=(1 if eager else 2) * n * 32 * 16, use_cuda_graph=not eager, **kw)
print(prof.key_averages().table(sort_by="self_cuda_time_total" if eager else "cuda_time_total", row_limit=45, max_name_column_width=70))
