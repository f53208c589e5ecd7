"""Periodic Poisson solver (extension for BASELINE config 4: periodic Taylor-Green vortex). The reference has nothing
periodic: parity is UNPINNED by the reference and rests on analytic answers - single Fourier modes,
the Taylor-Green stream function, and the exact-inverse property of the three-point symbol."""

import numpy as np
import pytest

TOL = {"float32": 1e-5, "float64": 1e-12}


def _rel_l2(a, b):
    return float(np.linalg.norm((np.asarray(a, dtype=np.float64) - b).ravel()) / np.linalg.norm(np.ravel(b)))


def _taylor_green(grid, x_range):
    """A Taylor-Green-type single mode cos cos cos on the periodic box (the vorticity of u = (sin x cos y cos z,
    -cos x sin y cos z, 0) has this shape; wavenumber
    2 pi / L per axis) and the psi_z with -lap psi_z = omega_z: psi = omega / |k|^2."""
    nz, ny, nx = grid
    dx = x_range / nx
    z, y, x = np.meshgrid((np.arange(nz) + 0.5) * dx, (np.arange(ny) + 0.5) * dx, (np.arange(nx) + 0.5) * dx,
                          indexing="ij")
    kx, ky, kz = 2 * np.pi / x_range, 2 * np.pi / (ny * dx), 2 * np.pi / (nz * dx)
    omega = -(kx + ky) * np.cos(kx * x) * np.cos(ky * y) * np.cos(kz * z)
    return omega, omega / (kx * kx + ky * ky + kz * kz)


def test_oracle_periodic_analytic():
    from oracle.poisson import PeriodicPoissonSolver

    grid, x_range = (12, 16, 20), 2.0
    omega, psi = _taylor_green(grid, x_range)
    sol = np.zeros(grid)
    PeriodicPoissonSolver(grid, x_range / grid[2], "spectral").solve(sol, omega + 3.0)  # the mean is dropped
    assert _rel_l2(sol, psi) < 1e-13
    rng = np.random.default_rng(2)
    rhs = rng.standard_normal(grid)
    rhs -= rhs.mean()
    dx = x_range / grid[2]
    PeriodicPoissonSolver(grid, dx, "three_point").solve(sol, rhs)
    lap = sum(np.roll(sol, 1, a) + np.roll(sol, -1, a) for a in range(3)) - 6 * sol
    assert _rel_l2(-lap / dx**2, rhs) < 1e-12 and abs(sol.mean()) < 1e-14


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("grid", [(12, 16, 20), (17, 19, 23), (32, 32, 64)])
def test_cuda_periodic_poisson_3d(dtype, grid):
    import torch

    import sopht_b200.numeric.eulerian_grid_ops as spne
    from oracle.poisson import PeriodicPoissonSolver

    real_t = np.float32 if dtype == "float32" else np.float64
    x_range = 2.0
    omega, psi = _taylor_green(grid, x_range)
    solver = spne.PeriodicPoissonSolver3D(*grid, x_range=x_range, real_t=real_t)
    pow2 = dtype == "float32" and all(n & (n - 1) == 0 for n in grid)
    assert solver.path == ("periodic_pow2_spectral" if pow2 else "periodic_fft_spectral")
    rhs = torch.from_numpy((omega + 3.0).astype(real_t)).cuda()
    sol = torch.zeros_like(rhs)
    solver.solve(solution_field=sol, rhs_field=rhs)
    assert _rel_l2(sol.cpu().numpy(), psi) < TOL[dtype]
    # random vector rhs against the numpy restatement, both symbols; three_point inverts the wrap-around Laplacian
    rng = np.random.default_rng(4)
    v = rng.standard_normal((3, *grid)).astype(real_t)
    dx = x_range / grid[2]
    for symbol in ("spectral", "three_point"):
        s = spne.PeriodicPoissonSolver3D(*grid, x_range=x_range, real_t=real_t, symbol=symbol)
        out = torch.zeros((3, *grid), dtype=rhs.dtype, device="cuda")
        s.vector_field_solve(solution_vector_field=out, rhs_vector_field=torch.from_numpy(v).cuda())
        ref = np.zeros(grid)
        oracle = PeriodicPoissonSolver(grid, dx, symbol)
        for c in range(3):
            oracle.solve(ref, v[c])
            assert _rel_l2(out[c].cpu().numpy(), ref) < TOL[dtype]
    if dtype == "float64":
        o = out[0]
        lap = sum(torch.roll(o, 1, a) + torch.roll(o, -1, a) for a in range(3)) - 6 * o
        want = torch.from_numpy(v[0] - v[0].mean()).cuda()
        assert float(torch.linalg.vector_norm(-lap / dx**2 - want) / torch.linalg.vector_norm(want)) < 1e-11
    with pytest.raises(ValueError, match="symbol"):
        spne.PeriodicPoissonSolver3D(*grid, symbol="chebyshev")


@pytest.mark.gpu
@pytest.mark.parametrize("grid", [(16, 16, 32), (32, 64, 128), (64, 16, 256), (16, 128, 64), (256, 32, 32),
                                  (128, 128, 256), (512, 16, 32), (16, 512, 32), (16, 16, 1024), (1024, 16, 32),
                                  (16, 2048, 32)])
def test_cuda_periodic_poisson_pow2_pipeline(grid):
    """fp32 power-of-two grids: the hand-written FFT pipeline (five in-place kernels, no cuFFT) against the numpy
    restatement and against the cuFFT-based path of the same library, both symbols, scalar / vector / strided-view
    solves; the grids reach every transform length family (two- and three-pass, radix 16 / 32)."""
    import os

    import torch

    import sopht_b200.numeric.eulerian_grid_ops as spne
    from oracle.poisson import PeriodicPoissonSolver

    x_range = 1.5
    dx = x_range / grid[2]
    rng = np.random.default_rng(12)
    v = rng.standard_normal((3, *grid)).astype(np.float32)
    for symbol in ("spectral", "three_point"):
        s = spne.PeriodicPoissonSolver3D(*grid, x_range=x_range, real_t=np.float32, symbol=symbol)
        assert s.path == "periodic_pow2_" + symbol
        out = torch.full((3, *grid), 7.0, device="cuda")
        s.vector_field_solve(solution_vector_field=out, rhs_vector_field=torch.from_numpy(v).cuda())
        oracle = PeriodicPoissonSolver(grid, dx, symbol)
        ref = np.zeros((3, *grid))
        for c in range(3):
            oracle.solve(ref[c], v[c].astype(np.float64))
        assert _rel_l2(out.cpu().numpy(), ref) < 1e-5
        # scalar solve into a z-halo-padded array (what the periodic simulator passes) and in place
        padded = torch.zeros(3, grid[0] + 2, grid[1], grid[2], device="cuda")
        padded[:, 1:-1] = torch.from_numpy(v).cuda()
        s.vector_field_solve(solution_vector_field=padded[:, 1:-1], rhs_vector_field=padded[:, 1:-1])
        assert _rel_l2(padded[:, 1:-1].cpu().numpy(), ref) < 1e-5
        assert float(padded[:, 0].abs().max()) == 0.0 and float(padded[:, -1].abs().max()) == 0.0
        one = torch.zeros(*grid, device="cuda")
        s.solve(solution_field=one, rhs_field=torch.from_numpy(v[1]).cuda())
        assert _rel_l2(one.cpu().numpy(), ref[1]) < 1e-5
    assert os.environ.get("SOPHT_PERIODIC_FORCE_CUFFT", "0") == "0"


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_cuda_periodic_poisson_2d(dtype):
    import torch

    import sopht_b200.numeric.eulerian_grid_ops as spne

    real_t = np.float32 if dtype == "float32" else np.float64
    ny, nx, x_range = 24, 40, 1.0
    dx = x_range / nx
    y, x = np.meshgrid((np.arange(ny) + 0.5) * dx, (np.arange(nx) + 0.5) * dx, indexing="ij")
    kx, ky = 2 * np.pi * 2 / x_range, 2 * np.pi * 3 / (ny * dx)
    rhs = np.sin(kx * x) * np.cos(ky * y)
    solver = spne.PeriodicPoissonSolver2D(ny, nx, x_range=x_range, real_t=real_t)
    sol = torch.zeros((ny, nx), dtype=getattr(torch, dtype), device="cuda")
    solver.solve(solution_field=sol, rhs_field=torch.from_numpy(rhs.astype(real_t)).cuda())
    assert _rel_l2(sol.cpu().numpy(), rhs / (kx * kx + ky * ky)) < TOL[dtype]


def _taylor_green_2d_vorticity(sim_position_field, x_range):
    """omega_z = 2 k sin(k x) sin(k y), k = 2 pi / L, of u = (sin kx cos ky, -cos kx sin ky, 0): an exact solution of
    the Navier-Stokes equations whose amplitude decays as exp(-2 nu k^2 t) (the nonlinear term vanishes)."""
    k = 2 * np.pi / x_range
    x, y = sim_position_field[0], sim_position_field[1]
    return 2 * k * np.sin(k * x) * np.sin(k * y), k


def test_oracle_periodic_flow_step_taylor_green_decay():
    """CPU: the halo-wrapped composition of the reference's sub-steps reproduces the analytic Taylor-Green decay."""
    from oracle import flow as oflow

    x_range, nu = 1.0, 5e-3
    sim = oflow.PeriodicNavierStokesFlowSimulator3D((8, 32, 32), x_range, nu, real_t=np.float64)
    w0, k = _taylor_green_2d_vorticity(sim.position_field, x_range)
    sim.vorticity_field[2] = w0
    sim.compute_velocity_from_vorticity()
    # u_x = sin(kx) cos(ky) up to the second-order error of the centred curl
    u_exact = np.sin(k * sim.position_field[0]) * np.cos(k * sim.position_field[1])
    assert _rel_l2(sim.velocity_field[0], u_exact) < 2e-2
    steps, dt = 40, 2e-3
    for _ in range(steps):
        sim.time_step(dt)
    decay = np.exp(-2 * nu * k * k * steps * dt)
    assert _rel_l2(sim.vorticity_field[2], decay * w0) < 5e-3
    assert np.abs(sim.vorticity_field[:2]).max() < 1e-10 * np.abs(w0).max()  # the flow stays two-dimensional
    assert 0 < sim.compute_stable_timestep() < 1


@pytest.mark.gpu
@pytest.mark.parametrize("step_mode", ["fused", "unfused"])
@pytest.mark.parametrize("grid", [(12, 20, 24), (5, 9, 160), (40, 24, 132)])
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_cuda_periodic_flow_step_matches_oracle(dtype, grid, step_mode):
    """The CUDA simulator - the fused step (marching kernels with wrap-around x / y neighbour loads on z-halo-padded
    arrays) and the pass-by-pass composition - against the numpy restatement on a random periodic field (identical
    step counts; ragged grids: several warps per row, partial warps, one-CTA and multi-chunk z ranges), and the
    Taylor-Green decay on the device."""
    import torch

    from oracle import flow as oflow
    from sopht_b200.simulator import PeriodicNavierStokesFlowSimulator3D

    real_t = np.float32 if dtype == "float32" else np.float64
    x_range, nu = 1.0, 1e-2
    sim = PeriodicNavierStokesFlowSimulator3D(grid, x_range, nu, real_t=real_t, step_mode=step_mode)
    assert sim.step_mode == step_mode
    ref = oflow.PeriodicNavierStokesFlowSimulator3D(grid, x_range, nu, real_t=np.float64)
    rng = np.random.default_rng(8)
    w0 = rng.standard_normal((3, *grid)).astype(real_t)
    sim.vorticity_field[...] = torch.from_numpy(w0).cuda()
    ref.vorticity_field[...] = w0
    sim.compute_velocity_from_vorticity()
    ref.compute_velocity_from_vorticity()
    tol = 2e-5 if dtype == "float32" else 1e-11
    assert _rel_l2(sim.velocity_field.cpu().numpy(), ref.velocity_field) < tol
    for _ in range(3):
        sim.time_step(1e-4)
        ref.time_step(1e-4)
    assert _rel_l2(sim.vorticity_field.cpu().numpy(), ref.vorticity_field) < tol
    assert _rel_l2(sim.velocity_field.cpu().numpy(), ref.velocity_field) < tol
    # (the float64 restatement adds 10 eps(float64) to the diffusion limit, the simulator 10 eps(real_t))
    assert sim.compute_stable_timestep() == pytest.approx(ref.compute_stable_timestep(), rel=3e-3)

    tg = PeriodicNavierStokesFlowSimulator3D((8, 32, 32), x_range, 5e-3, real_t=real_t, step_mode=step_mode)
    w, k = _taylor_green_2d_vorticity(tg.position_field.cpu().numpy().astype(np.float64), x_range)
    tg.vorticity_field[2] = torch.from_numpy(w.astype(real_t)).cuda()
    tg.compute_velocity_from_vorticity()
    for _ in range(40):
        tg.time_step(2e-3)
    decay = np.exp(-2 * 5e-3 * k * k * 40 * 2e-3)
    assert _rel_l2(tg.vorticity_field[2].cpu().numpy(), decay * w) < 5e-3
