"""GPU: the z-slab decomposed path. With one rank it runs in-process on any box (same kernels, slab code path,
halo-padded local arrays); with >= 2 visible GPUs the multi-rank NCCL check (tests/mgpu_slab_check.py) is
launched through torchrun."""

import os
import subprocess
import sys

import numpy as np
import pytest
from conftest import REL_L2_TOL, rel_l2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("grid", [(16, 16, 32), (32, 8, 64)])
def test_slab_poisson_single_rank_matches_oracle(rng, grid):
    import torch

    from oracle import poisson as opoisson
    from sopht_b200.parallel import SlabUnboundedPoissonSolver3D

    nz, ny, nx = grid
    rhs = rng.standard_normal((3, nz, ny, nx)).astype(np.float32)
    solver = SlabUnboundedPoissonSolver3D(nz, ny, nx, x_range=1.0, real_t=np.float32)
    # strided views of a halo-padded array, as the simulator passes them
    padded_rhs = torch.zeros(3, nz + 2, ny, nx, device="cuda")
    padded_sol = torch.zeros_like(padded_rhs)
    padded_rhs[:, 1:-1] = torch.from_numpy(rhs).cuda()
    solver.vector_field_solve(padded_sol[:, 1:-1], padded_rhs[:, 1:-1])
    ref_solver = opoisson.UnboundedPoissonSolver3D(nz, ny, nx, x_range=1.0, real_t=np.float32)
    ref = np.zeros_like(rhs)
    for c in range(3):
        ref_solver.solve(ref[c], rhs[c])
    assert rel_l2(padded_sol[:, 1:-1].cpu().numpy(), ref) < REL_L2_TOL["single"]
    assert float(padded_sol[:, 0].abs().max()) == 0.0 and float(padded_sol[:, -1].abs().max()) == 0.0


@pytest.mark.parametrize("with_free_stream", [False, True])
@pytest.mark.parametrize("grid", [(16, 16, 32), (32, 16, 64)])
def test_slab_simulator_single_rank_matches_oracle(rng, grid, with_free_stream):
    import torch

    from oracle import flow as oflow
    from sopht_b200.parallel import SlabUnboundedNavierStokesFlowSimulator3D

    kw = dict(grid_size=grid, x_range=1.0, kinematic_viscosity=1e-2, real_t=np.float32,
              with_free_stream_flow=with_free_stream)
    sim = SlabUnboundedNavierStokesFlowSimulator3D(**kw)
    ref = oflow.UnboundedNavierStokesFlowSimulator3D(**kw)
    for name in ("vorticity_field", "velocity_field"):
        a = rng.standard_normal((3, *grid)).astype(np.float32)
        getattr(ref, name)[...] = a
        sim.set_owned(getattr(sim, name), a)
    dt = ref.compute_stable_timestep(dt_prefac=0.5)
    assert sim.compute_stable_timestep(dt_prefac=0.5) == pytest.approx(dt, rel=1e-6)
    fsv = [1.0, 2.0, 3.0] if with_free_stream else [0.0, 0.0, 0.0]
    for _ in range(2):
        sim.time_step(dt=dt, free_stream_velocity=fsv)
        ref.time_step(dt, free_stream_velocity=fsv)
        assert sim.compute_stable_timestep() == pytest.approx(ref.compute_stable_timestep(), rel=1e-5)
    for name in ("vorticity_field", "velocity_field", "stream_func_field"):
        err = rel_l2(sim.owned(getattr(sim, name)).cpu().numpy(), getattr(ref, name))
        assert err < REL_L2_TOL["single"], (name, err)
    assert torch.cuda.is_available()


def test_slab_single_rank_with_ib_matches_single_gpu_path():
    """world_size 1 through torchrun: forcing + SlabVirtualBoundaryForcing against the regular classes."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=1",
           "--master-addr", "127.0.0.1", "--master-port", "29516",
           os.path.join(ROOT, "tests", "mgpu_slab_check.py"), "16", "16", "32", "3"]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "SLAB CHECK OK" in out.stdout


@pytest.mark.parametrize("peer", ["1", "0"])  # transposes over peer memory (default) / NCCL all-to-all
@pytest.mark.parametrize("world", [2, 4])
def test_slab_multi_rank_nccl(world, peer):
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29517 + world),
           os.path.join(ROOT, "tests", "mgpu_slab_check.py"), "32", "16", "64", "3"]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, SOPHT_SLAB_PEER=peer))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "SLAB CHECK OK" in out.stdout


def test_peer_arena_single_rank():
    """Arena-backed tensors are ordinary zero-initialised CUDA tensors; with one rank the exchange and the barrier
    are no-ops (the multi-rank protocol is covered by test_slab_multi_rank_nccl on boxes with >= 2 GPUs)."""
    import torch

    from sopht_b200.parallel import PeerArena

    arena = PeerArena(2 * (3 * 6 * 8 * 16 * 4 + 256))
    a = arena.alloc((3, 6, 8, 16))
    b = arena.alloc((3, 6, 8, 16))
    assert a.is_cuda and a.dtype == torch.float32 and float(a.abs().max()) == 0.0
    assert a.data_ptr() % 256 == 0 and b.data_ptr() >= a.data_ptr() + a.numel() * 4
    a += 1.0
    b[:, 1:-1] = 2.0
    arena.halo_exchange((a, b), nz_local=4, halo=1)
    arena.barrier()
    torch.cuda.synchronize()
    assert float(a.sum()) == a.numel() and float(b[:, 0].abs().max()) == 0.0
    with pytest.raises(MemoryError):
        arena.alloc((3, 6, 8, 16))


# ---- periodic box (BASELINE config 4) on slabs -------------------------------------------------------------------------
def test_periodic_slab_single_rank_matches_single_gpu_path():
    """world_size 1 through torchrun: the slab classes of the periodic step (halo planes from the rank's own opposite
    planes, SlabPeriodicPoissonSolver3D with its in-place y / z passes) against PeriodicNavierStokesFlowSimulator3D."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=1",
           "--master-addr", "127.0.0.1", "--master-port", "29526",
           os.path.join(ROOT, "tests", "mgpu_periodic_check.py"), "32", "16", "64", "3"]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "PERIODIC SLAB CHECK OK" in out.stdout


@pytest.mark.parametrize("peer", ["1", "0"])  # transposes over peer memory (default) / NCCL all-to-all
@pytest.mark.parametrize("world", [2, 4, 8])
def test_periodic_slab_multi_rank_nccl(world, peer):
    """The ring halo exchange (rank 0 <-> rank P - 1) and the periodic slab Poisson solve on 2 / 4 / 8 GPUs."""
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29530 + world),
           os.path.join(ROOT, "tests", "mgpu_periodic_check.py"), "64", "16", "256", "3"]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, SOPHT_SLAB_PEER=peer))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "PERIODIC SLAB CHECK OK" in out.stdout
