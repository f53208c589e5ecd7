"""Multi-GPU parity check of the z-slab decomposed flow step (launched by tests/test_slab_gpu.py or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
        tests/mgpu_slab_check.py [nz ny nx] [steps]

Every rank also runs the single-GPU simulator (+ virtual boundary forcing of a small sphere of Lagrangian
nodes straddling the slab interfaces) on the whole grid from the same seeded state and compares its own slab
of vorticity, velocity, stream function and the Lagrangian forces after `steps` coupled steps."""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sphere_nodes(centre, radius, n=200):
    k = np.arange(n) + 0.5
    phi = np.arccos(1 - 2 * k / n)
    theta = np.pi * (1 + 5**0.5) * k
    return np.stack([centre[0] + radius * np.cos(theta) * np.sin(phi),
                     centre[1] + radius * np.sin(theta) * np.sin(phi),
                     centre[2] + radius * np.cos(phi)])


def main() -> int:
    from sopht_b200.numeric.immersed_boundary_ops import VirtualBoundaryForcing
    from sopht_b200.parallel import SlabUnboundedNavierStokesFlowSimulator3D, SlabVirtualBoundaryForcing
    from sopht_b200.simulator import UnboundedNavierStokesFlowSimulator3D

    grid = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (32, 16, 64)
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    nz, ny, nx = grid
    kw = dict(grid_size=grid, x_range=1.0, kinematic_viscosity=1e-2, real_t=np.float32,
              with_forcing=True, with_free_stream_flow=True, flow_density=0.5)
    slab = SlabUnboundedNavierStokesFlowSimulator3D(**kw)
    full = UnboundedNavierStokesFlowSimulator3D(**kw)
    rng = np.random.default_rng(11)
    w0 = rng.standard_normal((3, *grid)).astype(np.float32)
    u0 = rng.standard_normal((3, *grid)).astype(np.float32)
    full.vorticity_field[...] = torch.from_numpy(w0).cuda()
    full.velocity_field[...] = torch.from_numpy(u0).cuda()
    slab.set_owned(slab.vorticity_field, w0)
    slab.set_owned(slab.velocity_field, u0)
    # a sphere in the middle of the domain: its 4-point supports cross every slab interface near the centre
    dx = float(full.dx)
    pos = torch.from_numpy(sphere_nodes((0.5, 0.5 * ny / nx, 0.5 * nz / nx), 0.2 * min(nz, ny, nx) / nx)).cuda()
    vel = torch.zeros_like(pos)
    vb_kw = dict(virtual_boundary_stiffness_coeff=-5e2 * dx * dx, virtual_boundary_damping_coeff=-3.0 * dx * dx,
                 grid_dim=3, dx=full.dx, num_lag_nodes=pos.shape[1], real_t=np.float32)
    vb_full = VirtualBoundaryForcing(**vb_kw)
    vb_slab = SlabVirtualBoundaryForcing(**vb_kw, partition=slab.part)
    dt_full = full.compute_stable_timestep(dt_prefac=0.5)
    dt_slab = slab.compute_stable_timestep(dt_prefac=0.5)
    ok = abs(dt_full - dt_slab) <= 1e-6 * abs(dt_full)
    fsv = [1.0, 0.5, -0.25]
    for _ in range(steps):
        vb_full.time_step(dt_full)
        vb_slab.time_step(dt_full)
        vb_full.compute_interaction_force_on_eul_and_lag_grid(
            full.eul_grid_forcing_field, full.velocity_field, pos, vel)
        vb_slab.compute_interaction_force_on_eul_and_lag_grid(
            slab.owned(slab.eul_grid_forcing_field), slab.owned(slab.velocity_field), pos, vel)
        full.time_step(dt=dt_full, free_stream_velocity=fsv)
        slab.time_step(dt=dt_full, free_stream_velocity=fsv)
        ok = ok and abs(full.compute_stable_timestep() - slab.compute_stable_timestep()) <= 1e-5 * dt_full
    worst = 0.0
    for name in ("vorticity_field", "velocity_field", "stream_func_field"):
        a = slab.owned(getattr(slab, name)).double()
        b = getattr(full, name)[:, slab.z_slice].double()
        num = (a - b).pow(2).sum()
        dist.all_reduce(num)
        err = float((num / getattr(full, name).double().pow(2).sum()).sqrt())
        worst = max(worst, err)
        if rank == 0:
            print(f"slab check {grid} x{world} ranks: {name} rel-L2 vs single GPU = {err:.3e}")
    f_err = float((vb_slab.lag_grid_forcing_field - vb_full.lag_grid_forcing_field).norm()
                  / vb_full.lag_grid_forcing_field.norm())
    worst = max(worst, f_err)
    if rank == 0:
        print(f"slab check: lagrangian forces rel-L2 = {f_err:.3e}")
    ok = ok and worst < 1e-5
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SLAB CHECK", "OK" if int(flag.item()) else "FAILED")
    dist.destroy_process_group()
    return 0 if int(flag.item()) else 1


if __name__ == "__main__":
    sys.exit(main())
