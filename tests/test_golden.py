"""Product programs against the COMMITTED golden vectors (tests/golden/*.npz, float64 oracle): CPU emulation here, the
CUDA build in tests/test_gpu_parity.py.  Free-running for 8 control steps (short enough that fp32 chaos stays below the
stated tolerance for the integer / flag outputs; qpos is compared after the first step only)."""
import os

import numpy as np
import pytest

import common
from backends import EmuBackend

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def run_against_golden(make_backend, name):
    g = np.load(os.path.join(GOLD, f"{name}_golden.npz"))
    ep = int(g["episode_length"])
    b = make_backend(common.setup(name, ep)[3])
    st, out = b.reset(g["keys"])
    np.testing.assert_array_equal(out["info_i"][:, 0], g["reset_cur_frame"])          # bit exact
    np.testing.assert_array_equal(st["qvel"], g["reset_qvel"])                         # bit exact (threefry uniform)
    np.testing.assert_allclose(st["qpos"], g["reset_qpos"], atol=1.2e-7)
    np.testing.assert_allclose(out["obs"], g["reset_obs"], atol=2e-5)
    first = {k: v.copy() for k, v in st.items()}
    first_obs, first_ii = out["obs"].copy(), out["info_i"].copy()
    for t in range(g["actions"].shape[0]):
        b.step(st, out, first, first_obs, first_ii, g["actions"][t])
        if t == 0:
            np.testing.assert_allclose(st["qpos"], g["step1_qpos"], atol=5e-5)          # fp32 tolerance, 1 control step
            np.testing.assert_allclose(st["qvel"], g["step1_qvel"], atol=2e-2)
            np.testing.assert_allclose(out["obs"], g["step1_obs"], atol=2e-2)
        np.testing.assert_array_equal(out["info_i"][:, 0], g["cur_frame"][t])           # frame index: bit exact
        np.testing.assert_array_equal(out["info_f"][:, 3], g["steps"][t])               # episode counter: bit exact
        np.testing.assert_array_equal(out["info_f"][:, 4], g["truncation"][t])
        np.testing.assert_array_equal(out["done"], g["done"][t])                        # done flags: bit exact
        if t < 3:
            np.testing.assert_allclose(out["reward"], g["reward"][t], atol=2e-2)


@pytest.mark.parametrize("name", ["rodent", "fly_free", "fly_tethered", "rodent_pair"])
def test_emulated_programs_match_golden(name):
    run_against_golden(EmuBackend, name)
