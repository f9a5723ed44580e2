"""CPU suite: the kernels' per-environment programs compiled for the host (one lane per environment) against the
oracle.  This pins the table logic, tree-sparse factorisation, matrix-free Jacobian, CG solver and the env layer
without a GPU; the warp-level build is checked by tests/test_gpu_parity.py on the B200."""
import pytest

import common
import parity_cases as pc
from backends import EmuBackend

# the fly never terminates within a short test (tethered: it cannot fall); a short episode forces the truncation /
# auto-reset path instead
FLY_EPISODE = 12


@pytest.fixture(scope="module")
def rodent_emu():
    return EmuBackend(common.setup("rodent")[3])


def test_forward_intermediates(rodent_emu):
    pc.check_forward_intermediates(rodent_emu, "rodent", N=4)


def test_reset(rodent_emu):
    pc.check_reset(rodent_emu, "rodent", N=16)


def test_teacher_forced_wrapped_step(rodent_emu):
    r = pc.check_teacher_forced(rodent_emu, "rodent", N=8, T=60)
    assert r["n_done"] > 0


def test_late_clip_window_clamps(rodent_emu):
    r = pc.check_late_clip(rodent_emu, "rodent", N=16, T=24)
    assert r["n_live"] > 0


def test_unwrapped_env_step_and_pipeline_init(rodent_emu):
    pc.check_unwrapped_step(rodent_emu, "rodent", N=4, T=8)


def test_nan_guard(rodent_emu):
    pc.check_nan_guard(rodent_emu, "rodent")


def test_physics_1_10_100(rodent_emu):
    pc.check_physics_1_10_100(rodent_emu, "rodent", N=4)


@pytest.mark.parametrize("name", ["fly_free", "fly_tethered"])
def test_fly_elliptic_cone(name):
    """configs[2]: fruit-fly imitation env (elliptic cones, fluid forces, claw-claw capsule contacts)."""
    b = EmuBackend(common.setup(name)[3])
    pc.check_forward_intermediates(b, name, N=4)
    pc.check_reset(b, name, N=8)
    pc.check_physics_1_10_100(b, name, N=16)      # medians over environments: 4 are too few for a chaotic free fall
    bt = EmuBackend(common.setup(name, FLY_EPISODE)[3])
    r = pc.check_teacher_forced(bt, name, N=8, T=30, episode_length=FLY_EPISODE)
    assert r["n_done"] > 0


def test_two_rodent_model_with_inter_animal_contacts():
    """configs[3]: assets/rodent_pair.xml -- two animals in one world (nv 146, 114 floor contacts + 12 opened inter-animal capsule
    pairs, two kinematic trees), both tracked (obs 1234); SURVEY Appendix C.3 (i), DESIGN.md "config 4"."""
    name = "rodent_pair"
    m, cfg, clip, tables = common.setup(name)
    assert (m.nq, m.nv, m.nu, m.na, m.nbody) == (148, 146, 60, 60, 133) and int(tables["obs_size"][0]) == 1234
    assert int(tables["ncon"][0]) == 126 and int(tables["ncross"][0]) == 12 and int(tables["n_animals"][0]) == 2
    b = EmuBackend(tables)
    pc.check_forward_intermediates(b, name, N=2)
    # the animals actually touching: active (penetrating) contacts between the two kinematic trees
    pc.check_forward_intermediates(b, name, N=4, states=pc.touching_states(name, 4, seed=1))
    pc.check_reset(b, name, N=4)
    pc.check_physics_1_10_100(b, name, N=2)
    bt = EmuBackend(common.setup(name, FLY_EPISODE)[3])
    r = pc.check_teacher_forced(bt, name, N=4, T=20, episode_length=FLY_EPISODE)
    assert r["n_done"] > 0


def test_multi_clip_gather():
    """RodentMultiClip: per-environment clip index drawn at reset, every clip gather offset by it."""
    print(pc.check_multi_clip(EmuBackend, N=12, T=8))
