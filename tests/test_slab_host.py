"""CPU (gloo, world_size 2 and 4): host logic of the z-slab decomposition — partition views, halo exchange
and the all-to-all choreography of the distributed Poisson solve (sopht_b200/parallel/slab*.py) with the
three compute phases played by numpy FFTs, against the oracle's single-process solve."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(rank, world, port, fn, args):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world, *args)
    finally:
        dist.destroy_process_group()


def _spawn(world, fn, *args):
    mp.spawn(_run, args=(world, _free_port(), fn, args), nprocs=world, join=True)


# ---------------------------------------------------------------------------------------------------------
def _halo_worker(rank, world, grid):
    from sopht_b200.parallel.slab import SlabPartition, exchange_halos

    nz, ny, nx = grid
    part = SlabPartition(grid, world, rank, halo=1)
    g = torch.arange(3 * nz * ny * nx, dtype=torch.float32).reshape(3, nz, ny, nx)
    local = torch.full((3, *part.local_shape), -1.0)
    part.owned(local)[...] = g[:, part.z_start : part.z_start + part.nz_local]
    exchange_halos(part, [local])
    lo = part.z_start - 1
    hi = part.z_start + part.nz_local
    if not part.is_first:
        assert torch.equal(local[:, 0], g[:, lo])
    else:
        assert (local[:, 0] == -1).all()  # outermost halo untouched
    if not part.is_last:
        assert torch.equal(local[:, -1], g[:, hi])
    else:
        assert (local[:, -1] == -1).all()
    # the view handed to the stencil kernels starts / ends on the global boundary plane on the edge ranks
    v = part.stencil_view(local)
    assert v.shape[1] == part.nz_local + (0 if part.is_first else 1) + (0 if part.is_last else 1)
    first_global = part.z_start - (0 if part.is_first else 1)
    assert torch.equal(v[:, 0], g[:, first_global])
    assert part.z_faces == (1 if rank == 0 else 0) | (2 if rank == world - 1 else 0)


@pytest.mark.parametrize("world", [2, 4])
def test_halo_exchange_and_views(world):
    _spawn(world, _halo_worker, (8, 3, 5))


def test_partition_errors():
    from sopht_b200.parallel.slab import SlabPartition

    with pytest.raises(ValueError):
        SlabPartition((6, 4, 4), 4, 0)
    with pytest.raises(ValueError):
        SlabPartition((8, 4, 4), 2, 2)
    p = SlabPartition((8, 4, 4), 1, 0)
    assert p.z_faces == 3 and p.nz_local == 8


# ---------------------------------------------------------------------------------------------------------
def _poisson_worker(rank, world, grid, tmp):
    """The three phases in numpy; the plan under test moves the data between them."""
    from oracle import poisson as opoisson
    from sopht_b200.parallel.slab import SlabPartition
    from sopht_b200.parallel.slab_poisson import SlabTransposePlan

    nz, ny, nx = grid
    C = 3
    rng = np.random.default_rng(7)
    rhs = rng.standard_normal((C, nz, ny, nx))
    ref_solver = opoisson.UnboundedPoissonSolver3D(nz, ny, nx, x_range=1.0, real_t=np.float64)
    g_hat = ref_solver.fourier_greens_function_times_dx_cubed  # (2nz, 2ny, nx+1), real part used by multiply
    part = SlabPartition(grid, world, rank)
    plan = SlabTransposePlan(part, C, torch.float64, "cpu")
    nzl, nxl = plan.nzl, plan.nxl
    z0 = part.z_start
    sol_local = np.zeros((C, nzl, ny, nx))

    def as_c(t):  # (..., 2) real tensor -> complex numpy view
        return t.numpy().view(np.complex128)[..., 0]

    def forward_x():
        spec = np.fft.rfft(rhs[:, z0 : z0 + nzl], n=2 * nx, axis=-1)  # (C, nzl, ny, nx+1)
        send = as_c(plan.send)  # (C, P, nzl, ny, nxl)
        for q in range(world):
            send[:, q] = spec[..., q * nxl : (q + 1) * nxl]
        as_c(plan.nyq_local)[...] = spec[..., nx]

    def middle():
        recv = as_c(plan.recv).reshape(C, nz, ny, nxl)
        k0 = rank * nxl
        f = np.fft.fft(np.fft.fft(recv, n=2 * ny, axis=2), n=2 * nz, axis=1)
        f *= g_hat[None, :, :, k0 : k0 + nxl]
        recv[...] = np.fft.ifft(np.fft.ifft(f, axis=1)[:, :nz], axis=2)[:, :, :ny]
        nq = as_c(plan.nyq_all)  # (C, nz, ny)
        f = np.fft.fft(np.fft.fft(nq, n=2 * ny, axis=2), n=2 * nz, axis=1) * g_hat[None, :, :, nx]
        nq[...] = np.fft.ifft(np.fft.ifft(f, axis=1)[:, :nz], axis=2)[:, :, :ny]

    def inverse_x():
        send = as_c(plan.send)
        spec = np.zeros((C, nzl, ny, nx + 1), dtype=np.complex128)
        for q in range(world):
            spec[..., q * nxl : (q + 1) * nxl] = send[:, q]
        spec[..., nx] = as_c(plan.nyq_local)
        sol_local[...] = np.fft.irfft(spec, n=2 * nx, axis=-1)[..., :nx]

    plan.solve(forward_x, middle, inverse_x)
    ref = np.zeros_like(rhs)
    for c in range(C):
        ref_solver.solve(ref[c], rhs[c])
    err = np.linalg.norm(sol_local - ref[:, z0 : z0 + nzl]) / np.linalg.norm(ref)
    assert err < 1e-12, err


@pytest.mark.parametrize("world", [1, 2, 4])
def test_slab_transpose_plan_matches_oracle(world, tmp_path):
    if world == 1:
        os.environ.pop("RANK", None)
        _poisson_worker(0, 1, (8, 4, 8), str(tmp_path))
    else:
        _spawn(world, _poisson_worker, (8, 4, 8), str(tmp_path))
