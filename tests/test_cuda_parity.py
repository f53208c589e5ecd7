"""GPU: the CUDA path (reference-named factories -> ctypes -> C ABI -> sm_100a kernels) against the
golden vectors generated from the reference (tests/golden) — the same cases the oracle is pinned with
in test_oracle_golden.py. Tolerances: the reference's own atol = 1e3*eps, plus the north-star per-kernel
relative L2 bound (1e-5 fp32 / 1e-12 fp64) where the case checks it."""

import pytest
from adapters import make_ops
from kernel_cases import POISSON_CASES, STENCIL_CASES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("case", STENCIL_CASES, ids=lambda c: c.__name__)
def test_cuda_stencils_match_reference_goldens(case, precision):
    case(make_ops("cuda", precision), precision)


@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("case", POISSON_CASES, ids=lambda c: c.__name__)
def test_cuda_poisson_matches_reference_goldens(case, precision):
    case(make_ops("cuda", precision), precision)


@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("case", POISSON_CASES, ids=lambda c: c.__name__)
def test_cuda_poisson_generic_path_matches_reference_goldens(case, precision):
    from sopht_b200.numeric.eulerian_grid_ops.poisson_solvers import POISSON_FORCE_GENERIC

    case(make_ops("cuda", precision), precision, flags=POISSON_FORCE_GENERIC)
