"""CPU: pin the oracle (oracle/*.py) against the golden vectors generated from the reference's own
test modules and numba implementation (tests/golden/make_golden.py). An oracle that drifts from the
reference fails here, before it is used to judge the CUDA path."""

import pytest
from adapters import make_ops
from kernel_cases import POISSON_CASES, STENCIL_CASES


@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("case", STENCIL_CASES, ids=lambda c: c.__name__)
def test_oracle_stencils_match_reference_goldens(case, precision):
    case(make_ops("oracle", precision), precision)


@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("case", POISSON_CASES, ids=lambda c: c.__name__)
def test_oracle_poisson_matches_reference_goldens(case, precision):
    case(make_ops("oracle", precision), precision)


# ---- immersed boundary: oracle vs the reference's own numba implementation ----------------------------
class _OracleComm:
    def __init__(self, dim, dx, shift, n, real_t, n_components, kernel_type):
        from oracle import ib

        self.ib, self.dx, self.shift, self.kernel_type = ib, dx, shift, kernel_type

    def local_eulerian_grid_support_of_lagrangian_grid_kernel(self, support, idx, pos):
        self.ib.local_eulerian_grid_support_of_lagrangian_grid(support, idx, pos, self.dx, self.shift)

    def interpolation_weights_kernel(self, weights, support):
        fn = self.ib.cosine_interpolation_weights if self.kernel_type == "cosine" else self.ib.peskin_interpolation_weights
        fn(weights, support, self.dx)

    def eulerian_to_lagrangian_grid_interpolation_kernel(self, lag, eul, weights, idx):
        self.ib.eulerian_to_lagrangian_grid_interpolation(lag, eul, weights, idx, self.dx)

    def lagrangian_to_eulerian_grid_interpolation_kernel(self, eul, lag, weights, idx):
        self.ib.lagrangian_to_eulerian_grid_interpolation(eul, lag, weights, idx)


@pytest.mark.parametrize("precision", ["single", "double"])
def test_oracle_ib_communicator_matches_reference(precision):
    from ib_cases import case_communicator

    case_communicator(_OracleComm, precision)


@pytest.mark.parametrize("precision", ["single", "double"])
def test_oracle_virtual_boundary_matches_reference(precision):
    from ib_cases import case_virtual_boundary
    from oracle import ib

    def factory(dim, k, c, dx, n, real_t):
        return ib.VirtualBoundaryForcing(k, c, dim, dx, n, real_t)

    case_virtual_boundary(factory, precision)
