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


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()
