"""Known-answer tests that pin the threefry core of the oracle, of the host key plumbing and (through
tests/test_*_parity.py::check_reset) of the device reset.  Vectors: JAX's own unit tests / docs, as listed in SURVEY.md
Appendix D."""
import numpy as np

import env_oracle
from brax_tracking_b200 import prng

KAT = [((0x0, 0x0), (0x0, 0x0), (0x6B200159, 0x99BA4EFE)),
       ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
       ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]


def test_threefry2x32_known_answers():
    for key, ctr, want in KAT:
        for impl in (env_oracle.threefry2x32, prng.threefry2x32):
            x0, x1 = impl((np.uint32(key[0]), np.uint32(key[1])), np.array([ctr[0]], np.uint32), np.array([ctr[1]], np.uint32))
            assert (int(x0[0]), int(x1[0])) == want


def test_split_of_prngkey_zero():
    want = np.array([[4146024105, 967050713], [2718843009, 1272950319]], dtype=np.uint32)
    assert np.array_equal(env_oracle.split((np.uint32(0), np.uint32(0)), 2), want)
    assert np.array_equal(prng.split(prng.PRNGKey(0), 2), want)


def test_uniform_and_randint_ranges():
    k = (np.uint32(1), np.uint32(2))
    u = env_oracle.uniform(k, 1001, -1e-3, 1e-3)
    assert u.dtype == np.float32 and u.min() >= -1e-3 and u.max() < 1e-3 and len(np.unique(u)) > 900
    r = [env_oracle.randint((np.uint32(i), np.uint32(7)), 0, 44) for i in range(300)]
    assert min(r) >= 0 and max(r) < 44 and len(set(r)) > 35
    # odd-length draws pad the counter array (jax _threefry_random_bits): counters (0..3 | 4..7) for n = 7 and n = 8
    b7, b8 = env_oracle.random_bits(k, 7), env_oracle.random_bits(k, 8)
    assert np.array_equal(b7, np.concatenate([b8[:4], b8[4:7]]))
