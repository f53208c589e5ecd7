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


# (256, 512, 32) / (512, 256, 16) reach the radix-32 transform lengths (2ny / 2nz = 512, 1024: the row-mode z kernel of
# poisson_zrow.cuh), (8, 8, 2048) the three-pass L = 2048 row kernel, (1024, 8, 16) / (8, 1024, 16) the three-pass
# L = 2048 z and y column kernels that the 1024^3 slab runs use
@pytest.mark.parametrize("grid", [(8, 8, 16), (16, 32, 64), (32, 16, 128), (64, 128, 256), (128, 128, 256),
                                  (256, 512, 32), (512, 256, 16), (8, 8, 2048), (1024, 8, 16), (8, 1024, 16)])
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


# the z pass of the 2 nz = 1024 solve has three implementations (warp-quartet kernel: default; SOPHT_P2_ZQUAD=0: row-mode
# kernel; SOPHT_P2_ZROW=0: round-1 column kernel) and the chain has an opt-in programmatic-dependent-launch mode; the
# library reads these switches once per process, so every variant runs in its own interpreter
@pytest.mark.parametrize("env", [{}, {"SOPHT_P2_ZQUAD": "0"}, {"SOPHT_P2_ZROW": "0"}, {"SOPHT_PDL": "1"}],
                         ids=["zquad", "zrow", "zconv", "zquad-pdl"])
def test_cuda_poisson_z_pass_variants_vs_oracle(env):
    import os
    import subprocess
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    out = subprocess.run([sys.executable, os.path.join(here, "poisson_variant_check.py"), "512", "16", "32"],
                         env={**os.environ, **env}, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "POISSON VARIANT OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("grid", [(16, 16, 16), (9, 21, 70), (40, 24, 130)])
def test_cuda_fused_ns3d_passes_vs_oracle(precision, grid):
    """Fused simulator-level passes against the composition of the oracle's per-kernel restatements
    (ragged grids exercise partial tiles, chunk seams and the ghost ring)."""
    import ctypes

    import numpy as np
    import torch
    from conftest import REL_L2_TOL, real_t_of, rel_l2

    from oracle import stencils as ost
    from sopht_b200 import _lib

    real_t = real_t_of(precision)
    dc = _lib.dtype_code(real_t)
    rng = np.random.default_rng(11)
    w = rng.standard_normal((3, *grid)).astype(real_t)
    u = rng.standard_normal((3, *grid)).astype(real_t)
    p, q = real_t(0.37), real_t(0.11)
    lib, st = _lib.load(), _lib.current_stream()
    dev = lambda a: torch.from_numpy(a).cuda()  # noqa: E731
    fd = lambda t: ctypes.byref(_lib.field_desc(t, dc))  # noqa: E731

    # advect: w + p curl(u x w)
    b = np.zeros_like(w)
    ost.elementwise_cross_product(b, u, w)
    want = w.copy()
    ost.update_vorticity_from_velocity_forcing_3d(want, b, p)
    dw, du, dout = dev(w), dev(u), torch.full((3, *grid), 7.0, dtype=dev(w).dtype, device="cuda")
    _lib.check(lib.sopht_ns3d_advect_rotational(dc, fd(dout), fd(dw), fd(du), float(p), st))
    assert rel_l2(dout.cpu().numpy(), want) < REL_L2_TOL[precision]
    with pytest.raises(ValueError):  # aliasing is refused, not silently wrong
        _lib.check(lib.sopht_ns3d_advect_rotational(dc, fd(dw), fd(dw), fd(du), float(p), st))

    # diffuse (+ zeroing a third field)
    want = w.copy()
    ost.diffusion_timestep_euler_forward_vector(want, np.zeros(grid, real_t), q)
    dz = dev(u.copy())
    dout.fill_(7.0)
    _lib.check(lib.sopht_ns3d_diffuse(dc, fd(dout), fd(dw), float(q), fd(dz), st))
    assert rel_l2(dout.cpu().numpy(), want) < REL_L2_TOL[precision]
    assert float(dz.abs().max()) == 0.0

    # velocity from stream function, free stream, max reduction
    want = np.ones_like(w)
    ost.curl_3d(want, w, p)
    fsv = [0.5, -1.0, 2.0]
    ost.add_fixed_val_vector(want, want, fsv)
    vmax = torch.zeros(1, dtype=dw.dtype, device="cuda")
    _lib.check(lib.sopht_ns3d_velocity_from_stream_function(
        dc, fd(dout), fd(dw), float(p), _lib.double_array(fsv), ctypes.c_void_p(vmax.data_ptr()), st))
    assert rel_l2(dout.cpu().numpy(), want) < REL_L2_TOL[precision]
    assert float(vmax.item()) == pytest.approx(float(np.abs(want).sum(axis=0).max()), rel=1e-5)


# ---- ragged grids for the kernels the goldens only cover at 16^3 ----------------------
@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("grid", [(9, 21, 70), (17, 19, 23), (40, 24, 130)])
def test_cuda_stretching_divergence_brinkman_on_ragged_grids(precision, grid):
    """Vortex-stretching flux / Euler / SSP-RK3 steps, divergence, Brinkman penalisation (+ the fixed-value variant)
    and the level-set characteristic function on non-cubic, non-power-of-two grids and on strided component views,
    against the oracle (the reference's tests use 16^3 only)."""
    import numpy as np
    from conftest import REL_L2_TOL, real_t_of, rel_l2

    real_t = real_t_of(precision)
    tol = REL_L2_TOL[precision]
    rng = np.random.default_rng(23)
    cu, orc = make_ops("cuda", precision), make_ops("oracle", precision)
    w = rng.standard_normal((3, *grid)).astype(real_t)
    u = rng.standard_normal((3, *grid)).astype(real_t)
    p = real_t(0.21)

    def both(fn, *outs):
        got = [o.copy() for o in outs]
        want = [o.copy() for o in outs]
        fn(cu, *got)
        fn(orc, *want)
        for g_, w_ in zip(got, want):
            assert rel_l2(g_, w_) < tol

    both(lambda o, q: o.stretching_flux(q, w, u, p), np.full_like(w, 3.0))
    both(lambda o, ww, q: o.stretching_timestep(ww, u, q, p), w, np.full_like(w, 3.0))
    mid = np.zeros_like(w)
    both(lambda o, ww, q: o.stretching_timestep(ww, u, q, p, stepper="ssprk3", midstep=mid.copy()), w,
         np.full_like(w, 3.0))
    both(lambda o, d: o.divergence_3d(d, w, p), np.full(grid, 5.0, real_t))
    both(lambda o, d: o.divergence_3d(d, w, p, reset=False), np.full(grid, 5.0, real_t))
    chi = rng.random(grid).astype(real_t)
    both(lambda o, out: o.brinkmann(out, w[1], chi, u[2], 1e3), np.zeros(grid, real_t))  # strided component views
    both(lambda o, out: o.brinkmann(out, w, chi, u, 1e3, vector=True), np.zeros_like(w))
    ls = (rng.random(grid).astype(real_t) - real_t(0.5)) * real_t(0.2)
    both(lambda o, out: o.char_func(out, ls, 0.03), np.zeros(grid, real_t))
    g2 = grid[1:]
    f2, chi2 = rng.standard_normal(g2).astype(real_t), rng.random(g2).astype(real_t)
    both(lambda o, out: o.brinkmann_vs_fixed_val(out, f2, chi2, 1e3, 0.7), np.zeros(g2, real_t))


# ---- FFTPyFFTW{2,3}D plan objects and the scipy-named helper -------------------------------
# mirrors tests/test_numeric/test_eulerian_grid_ops/test_poisson_solver_{2,3}d/test_fft_kernel_{2,3}d.py of the reference
@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("shape", [(8, 8, 8), (6, 10, 18), (8, 8), (12, 20)])
def test_cuda_fft_plans_match_scipy(shape, precision):
    import numpy as np
    from conftest import real_t_of
    from scipy.fft import irfftn, rfftn

    from sopht_b200.numeric.eulerian_grid_ops import (
        FFTPyFFTW2D,
        FFTPyFFTW3D,
        fft_ifft_via_scipy_kernel_2d,
        fft_ifft_via_scipy_kernel_3d,
    )

    real_t = real_t_of(precision)
    tol = 1e-4 if precision == "single" else 1e-11  # get_test_tol of the reference scaled to these sizes
    rng = np.random.default_rng(42)
    field = rng.standard_normal(shape).astype(real_t)
    ref_f = rfftn(field)
    ref_i = irfftn(ref_f, s=shape)
    cls, helper = (FFTPyFFTW3D, fft_ifft_via_scipy_kernel_3d) if len(shape) == 3 else (FFTPyFFTW2D, fft_ifft_via_scipy_kernel_2d)
    plan = cls(*shape, num_threads=4, real_t=real_t)
    assert tuple(plan.field_pyfftw_buffer.shape) == shape
    assert tuple(plan.fourier_field_pyfftw_buffer.shape) == ref_f.shape
    fourier = np.zeros_like(ref_f)
    inv = np.zeros_like(ref_i)
    plan.fft_plan(input_array=field, output_array=fourier)
    np.testing.assert_allclose(fourier, ref_f, atol=tol * np.abs(ref_f).max())
    plan.ifft_plan(input_array=fourier.copy(), output_array=inv)  # a copy: complex-to-real plans destroy their input
    np.testing.assert_allclose(inv, ref_i, atol=tol)
    fourier2, inv2 = np.zeros_like(ref_f), np.zeros_like(ref_i)
    helper(fourier2, inv2, field, num_threads=4)
    np.testing.assert_allclose(fourier2, ref_f, atol=tol * np.abs(ref_f).max())
    np.testing.assert_allclose(inv2, field, atol=tol)
    # device tensors and the plan's own buffers
    import torch

    plan.field_pyfftw_buffer[...] = torch.from_numpy(field).cuda()
    plan.fft_plan()
    np.testing.assert_allclose(plan.fourier_field_pyfftw_buffer.cpu().numpy(), ref_f, atol=tol * np.abs(ref_f).max())
    plan.ifft_plan()
    np.testing.assert_allclose(plan.field_pyfftw_buffer.cpu().numpy(), field, atol=tol)
    with pytest.raises(ValueError):
        plan.fft_plan(input_array=torch.zeros(*shape[:-1], 2 * shape[-1], device="cuda",
                                              dtype=plan.field_pyfftw_buffer.dtype)[..., ::2])


@pytest.mark.parametrize("precision", ["single", "double"])
def test_cuda_poisson_exposes_greens_function_spectrum(precision):
    """fourier_greens_function_times_dx_cubed / _squared (UnboundedPoissonSolverPYFFTW3D.py:47-49, ...2D.py:42-44)."""
    import numpy as np
    from conftest import real_t_of, rel_l2

    from oracle import poisson as opoisson
    from sopht_b200.numeric.eulerian_grid_ops import UnboundedPoissonSolverPYFFTW2D, UnboundedPoissonSolverPYFFTW3D

    real_t = real_t_of(precision)
    tol = 1e-5 if precision == "single" else 1e-12
    s3 = UnboundedPoissonSolverPYFFTW3D(8, 16, 32, x_range=1.0, real_t=real_t)
    r3 = opoisson.UnboundedPoissonSolver3D(8, 16, 32, x_range=1.0, real_t=real_t)
    g3 = s3.fourier_greens_function_times_dx_cubed
    assert tuple(g3.shape) == (16, 32, 33) and g3.is_complex()
    assert rel_l2(g3.real.cpu().numpy(), np.real(r3.fourier_greens_function_times_dx_cubed)) < tol
    s2 = UnboundedPoissonSolverPYFFTW2D(12, 20, x_range=1.0, real_t=real_t)
    r2 = opoisson.UnboundedPoissonSolver2D(12, 20, x_range=1.0, real_t=real_t)
    g2 = s2.fourier_greens_function_times_dx_squared
    assert tuple(g2.shape) == (24, 21)
    assert rel_l2(g2.real.cpu().numpy(), np.real(r2.fourier_greens_function_times_dx_squared)) < tol
