"""Multi-GPU parity check of the z-slab decomposed flow step (launched by tests/test_slab_gpu.py or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
        tests/mgpu_slab_check.py [nz ny nx] [steps]

Every rank also runs the single-GPU simulator on the whole grid (same seeded state) and compares its own
slab of vorticity, velocity and stream function after `steps` steps; rank 0 prints the verdict."""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> int:
    from sopht_b200.parallel import SlabUnboundedNavierStokesFlowSimulator3D
    from sopht_b200.simulator import UnboundedNavierStokesFlowSimulator3D

    grid = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (32, 16, 64)
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    kw = dict(grid_size=grid, x_range=1.0, kinematic_viscosity=1e-2, real_t=np.float32, with_free_stream_flow=True)
    slab = SlabUnboundedNavierStokesFlowSimulator3D(**kw)
    full = UnboundedNavierStokesFlowSimulator3D(**kw)
    rng = np.random.default_rng(11)
    w0 = rng.standard_normal((3, *grid)).astype(np.float32)
    u0 = rng.standard_normal((3, *grid)).astype(np.float32)
    full.vorticity_field[...] = torch.from_numpy(w0).cuda()
    full.velocity_field[...] = torch.from_numpy(u0).cuda()
    slab.set_owned(slab.vorticity_field, w0)
    slab.set_owned(slab.velocity_field, u0)
    dt_full = full.compute_stable_timestep(dt_prefac=0.5)
    dt_slab = slab.compute_stable_timestep(dt_prefac=0.5)
    ok = abs(dt_full - dt_slab) <= 1e-6 * abs(dt_full)
    fsv = [1.0, 0.5, -0.25]
    worst = 0.0
    for _ in range(steps):
        full.time_step(dt=dt_full, free_stream_velocity=fsv)
        slab.time_step(dt=dt_full, free_stream_velocity=fsv)
        ok = ok and abs(full.compute_stable_timestep() - slab.compute_stable_timestep()) <= 1e-5 * dt_full
    for name in ("vorticity_field", "velocity_field", "stream_func_field"):
        a = slab.owned(getattr(slab, name)).double()
        b = getattr(full, name)[:, slab.z_slice].double()
        num = (a - b).pow(2).sum()
        den = getattr(full, name).double().pow(2).sum() / world
        dist.all_reduce(num)
        err = float((num / (den * world)).sqrt())
        worst = max(worst, err)
        if rank == 0:
            print(f"slab check {grid} x{world} ranks: {name} rel-L2 vs single GPU = {err:.3e}")
    ok = ok and worst < 1e-5
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SLAB CHECK", "OK" if int(flag.item()) else "FAILED")
    dist.destroy_process_group()
    return 0 if int(flag.item()) else 1


if __name__ == "__main__":
    sys.exit(main())
