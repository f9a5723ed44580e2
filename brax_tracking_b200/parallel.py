"""Environment sharding across the GPUs of one box (one process per GPU).

Reference: /root/reference/custom_brax/custom_ppo.py:199,213-223 -- ``num_envs`` must divide by the device count,
``key_envs = split(key_env, num_envs // process_count)`` is reshaped to ``(local_devices, envs_per_device, 2)`` and
each device resets / steps its own block.  The step has no cross-environment reduction, so shards never communicate;
NCCL is used only by the PPO learner (gradient mean all-reduce, custom_ppo.py:246-257).
"""
from __future__ import annotations

import numpy as np

from . import prng


def shard_bounds(num_envs: int, rank: int, world: int):
    if num_envs % world != 0:
        raise ValueError(f"num_envs ({num_envs}) must be divisible by the number of ranks ({world})")  # custom_ppo.py:199
    per = num_envs // world
    return rank * per, (rank + 1) * per


def shard_keys(key, num_envs: int, rank: int, world: int) -> np.ndarray:
    """This rank's block of ``jax.random.split(key, num_envs)`` ([envs_per_rank, 2] uint32)."""
    lo, hi = shard_bounds(num_envs, rank, world)
    return prng.split(key, num_envs)[lo:hi]
