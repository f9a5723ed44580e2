"""Host-side JAX PRNG (legacy threefry2x32) for key plumbing: ``PRNGKey`` / ``split`` as used by
/root/reference/custom_brax/custom_ppo.py:189-196,221.  The per-environment draws of ``reset`` (split(rng, 4),
randint, uniform; /root/reference/envs/fruitfly.py:451-475) run on the device (csrc/bt_math.h, bt_programs.h)."""
from __future__ import annotations

import numpy as np

_U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def PRNGKey(seed: int) -> np.ndarray:
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.array([seed >> 32, seed & 0xFFFFFFFF], dtype=_U32)


def threefry2x32(key, x0, x1):
    k0, k1 = _U32(key[0]), _U32(key[1])
    ks = (k0, k1, _U32(k0 ^ k1 ^ _U32(0x1BD11BDA)))
    with np.errstate(over="ignore"):
        x0 = (np.asarray(x0, dtype=_U32) + ks[0]).astype(_U32)
        x1 = (np.asarray(x1, dtype=_U32) + ks[1]).astype(_U32)
        for g in range(5):
            for r in _ROT[g % 2]:
                x0 = (x0 + x1).astype(_U32)
                x1 = ((x1 << _U32(r)) | (x1 >> _U32(32 - r))).astype(_U32)
                x1 = (x1 ^ x0).astype(_U32)
            x0 = (x0 + ks[(g + 1) % 3]).astype(_U32)
            x1 = (x1 + ks[(g + 2) % 3] + _U32(g + 1)).astype(_U32)
    return x0, x1


def split(key, num: int = 2) -> np.ndarray:
    cnt = np.arange(2 * num, dtype=_U32)
    y0, y1 = threefry2x32(key, cnt[:num], cnt[num:])
    return np.concatenate([y0, y1]).reshape(num, 2)
