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
    # odd-length draws pad the counter array with a ZERO (jax/_src/prng.py::threefry_2x32): counters (0 1 2 3 | 4 5 6 0) for
    # n = 7 against (0 1 2 3 | 4 5 6 7) for n = 8 -- only the element paired with the pad differs
    b7, b8 = env_oracle.random_bits(k, 7), env_oracle.random_bits(k, 8)
    assert np.array_equal(b7[:3], b8[:3]) and np.array_equal(b7[4:], b8[4:7]) and b7[3] != b8[3]
    y0, _ = env_oracle.threefry2x32(k, np.array([3], np.uint32), np.array([0], np.uint32))
    assert b7[3] == y0[0]


def test_odd_length_draws_match_published_jax_values():
    """Values printed by real JAX (legacy threefry, as in the 2024 releases the reference ran on), from the JAX PRNG
    documentation: ``random.uniform(random.PRNGKey(0))`` = 0.41845703 and ``random.normal(random.PRNGKey(0), (1,))`` =
    [-0.20584226].  Both are ONE-element draws, i.e. the odd-length path: the single counter 0 is paired with the pad 0
    (pairing it with n = 1 instead gives 0.5995 / 0.2519...), so they pin the padding rule that every ``randint`` start
    frame and every odd-sized ``uniform`` (rodent nv = 73) goes through."""
    from scipy.special import erfinv
    key = (np.uint32(0), np.uint32(0))
    bits = env_oracle.random_bits(key, 1)
    assert int(bits[0]) == 0x6B200159                      # = threefry2x32((0, 0), (0, 0))[0], the first KAT above
    u = env_oracle.uniform(key, 1, 0.0, 1.0)
    assert abs(float(u[0]) - 0.41845703) < 5e-9
    # jax.random.normal: sqrt(2) * erfinv(uniform(key, minval=nextafter(-1, 0), maxval=1))
    lo = np.nextafter(np.float32(-1.0), np.float32(0.0))
    un = env_oracle.uniform(key, 1, lo, 1.0)
    assert abs(float(np.sqrt(2.0) * erfinv(np.float64(un[0]))) - (-0.20584226)) < 2e-7


def test_device_bit_generator_matches_the_oracle_for_odd_and_even_lengths():
    """csrc/bt_math.h::bt_random_bits (compiled for the host by tests/host_emu) element by element against the oracle."""
    import emu
    lib = emu.build()
    import ctypes as C
    l = C.CDLL(lib)
    l.emu_random_bits.restype = C.c_uint32
    for k0, k1 in ((0, 0), (1, 2), (0xDEADBEEF, 0x12345678)):
        for n in (1, 2, 3, 7, 8, 73, 74, 146):
            want = env_oracle.random_bits((np.uint32(k0), np.uint32(k1)), n)
            got = np.array([l.emu_random_bits(C.c_uint32(k0), C.c_uint32(k1), i, n) for i in range(n)], dtype=np.uint32)
            assert np.array_equal(got, want), (k0, k1, n)
