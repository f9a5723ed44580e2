"""Shared fixtures: compiled models, synthetic clips, packed tables, oracles (TEST INFRASTRUCTURE)."""
from __future__ import annotations

import functools

import numpy as np

import env_oracle
import oracle
from brax_tracking_b200 import configs, model, presets

ENV_ARGS = presets.ENV_ARGS


@functools.lru_cache(maxsize=None)
def setup(name: str, episode_length: int = None):
    """-> (mjcf.Model, cfg, clip dict, packed tables); `episode_length` overrides main.py:86 (tests of the truncation path)"""
    m, args, clip = presets.load(name)
    cfg = configs.resolve(m, args)
    if episode_length is not None:
        cfg["episode_length"] = int(episode_length)
    return m, cfg, clip, model.pack(m, cfg, clip)


@functools.lru_cache(maxsize=None)
def oracles(name: str, dtype=np.float64, episode_length: int = None):
    m, cfg, clip, _ = setup(name, episode_length)
    o = oracle.Oracle(m, dtype)
    return o, env_oracle.EnvOracle(o, clip, cfg, dtype=np.float32)


def jax_keys(n: int, seed: int = 0) -> np.ndarray:
    """jax.random.split(jax.random.PRNGKey(seed), n) -> [n, 2] uint32 (custom_ppo.py:221)."""
    return env_oracle.split((np.uint32(0), np.uint32(seed)), n)


def actions(n_steps: int, n: int, nu: int, seed: int = 1, scale: float = 1.0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return (scale * np.tanh(rng.standard_normal((n_steps, n, nu)))).astype(np.float32)


def region(tables, scratch, name, size):
    o = int(tables["o_" + name][0])
    return scratch[..., o:o + size]
