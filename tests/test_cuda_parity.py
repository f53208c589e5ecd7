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


@pytest.mark.parametrize("grid", [(8, 8, 16), (16, 32, 64), (32, 16, 128), (64, 128, 256), (128, 128, 256)])
def test_cuda_poisson_pow2_path_vs_oracle(grid):
    """fp32 power-of-two fast path (hand-written pruned FFT pipeline) against the scipy.fft oracle, which is
    itself pinned to the reference's test restatement; scalar, vector and strided-view solves."""
    import numpy as np
    import torch
    from conftest import rel_l2

    from oracle import poisson as opoisson
    from sopht_b200.numeric.eulerian_grid_ops import UnboundedPoissonSolverPYFFTW3D

    rng = np.random.default_rng(7)
    solver = UnboundedPoissonSolverPYFFTW3D(*grid, x_range=1.0, real_t=np.float32)
    assert solver.path == "pow2"
    ref = opoisson.UnboundedPoissonSolver3D(*grid, x_range=1.0, real_t=np.float32, workers=8)
    rhs = rng.standard_normal((3, *grid)).astype(np.float32)
    want = np.zeros_like(rhs)
    ref.vector_field_solve(want, rhs)
    got = torch.zeros(3, *grid, device="cuda")
    solver.vector_field_solve(solution_vector_field=got, rhs_vector_field=torch.from_numpy(rhs).cuda())
    assert rel_l2(got.cpu().numpy(), want) < 1e-5
    one = torch.zeros(*grid, device="cuda")
    solver.solve(solution_field=one, rhs_field=torch.from_numpy(rhs[1]).cuda())
    assert rel_l2(one.cpu().numpy(), want[1]) < 1e-5
    # in-place solve (solution aliases rhs) and a component view of a larger tensor
    buf = torch.from_numpy(rhs).cuda()
    solver.solve(solution_field=buf[2], rhs_field=buf[2])
    assert rel_l2(buf[2].cpu().numpy(), want[2]) < 1e-5
    # x-strided view -> generic fallback inside the same handle
    wide = torch.zeros(*grid[:2], 2 * grid[2], device="cuda")
    wide[..., ::2] = torch.from_numpy(rhs[0]).cuda()
    out = torch.zeros_like(wide)
    solver.solve(solution_field=out[..., ::2], rhs_field=wide[..., ::2])
    assert rel_l2(out[..., ::2].cpu().numpy(), want[0]) < 1e-5
