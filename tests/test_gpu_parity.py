"""GPU parity tests proper: the sm_100a kernels, called through the C ABI, against the oracle on identical inputs."""
import pytest

import common
import parity_cases as pc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rodent_cuda():
    from backends import CudaBackend
    return CudaBackend(common.setup("rodent")[3])


def test_forward_intermediates(rodent_cuda):
    pc.check_forward_intermediates(rodent_cuda, "rodent", N=16)


def test_reset(rodent_cuda):
    pc.check_reset(rodent_cuda, "rodent", N=64)


def test_teacher_forced_wrapped_step_100(rodent_cuda):
    r = pc.check_teacher_forced(rodent_cuda, "rodent", N=16, T=100)
    print(r)


def test_physics_1_10_100(rodent_cuda):
    print(pc.check_physics_1_10_100(rodent_cuda, "rodent", N=8))


@pytest.mark.parametrize("name", ["fly_free", "fly_tethered"])
def test_fly_elliptic_cone(name):
    """configs[2]: fruit-fly imitation env, 8192-env kernel variant (2 dof slots, 1 contact slot per lane)."""
    from backends import CudaBackend
    b = CudaBackend(common.setup(name)[3])
    pc.check_forward_intermediates(b, name, N=16)
    pc.check_reset(b, name, N=64)
    pc.check_physics_1_10_100(b, name, N=8)
    bt = CudaBackend(common.setup(name, 12)[3])
    print(pc.check_teacher_forced(bt, name, N=16, T=40, episode_length=12))


@pytest.mark.parametrize("name", ["rodent", "fly_free", "fly_tethered"])
def test_cuda_matches_golden(name):
    from backends import CudaBackend
    from test_golden import run_against_golden
    run_against_golden(CudaBackend, name)


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()
