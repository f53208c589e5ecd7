"""Multi-GPU parity check of the z-slab decomposed PERIODIC flow step (launched by tests/test_slab_gpu.py or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
        tests/mgpu_periodic_check.py [nz ny nx] [steps]

Every rank also runs the single-GPU periodic simulator on the whole grid from the same seeded state and compares its own
planes of vorticity, velocity and stream function after `steps` steps, plus the stable time step (an all-reduce)."""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> int:
    from sopht_b200.parallel import SlabPeriodicNavierStokesFlowSimulator3D
    from sopht_b200.simulator import PeriodicNavierStokesFlowSimulator3D

    grid = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (32, 16, 64)
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    for symbol in ("spectral", "three_point"):
        slab = SlabPeriodicNavierStokesFlowSimulator3D(grid, 1.0, 1e-2, poisson_symbol=symbol)
        full = PeriodicNavierStokesFlowSimulator3D(grid, 1.0, 1e-2, real_t=np.float32, poisson_symbol=symbol)
        rng = np.random.default_rng(13)
        w0 = rng.standard_normal((3, *grid)).astype(np.float32)
        full.vorticity_field[...] = torch.from_numpy(w0).cuda()
        slab.set_owned(slab.vorticity_field, w0)
        full.compute_velocity_from_vorticity()
        slab.compute_velocity_from_vorticity()
        dt = full.compute_stable_timestep(dt_prefac=0.5)
        ok = ok and abs(slab.compute_stable_timestep(dt_prefac=0.5) - dt) <= 1e-5 * abs(dt)
        for _ in range(steps):
            full.time_step(dt)
            slab.time_step(dt)
        ok = ok and abs(full.compute_stable_timestep() - slab.compute_stable_timestep()) <= 1e-5 * dt
        worst = 0.0
        for name in ("vorticity_field", "velocity_field", "stream_func_field"):
            a = slab.owned(getattr(slab, name)).double()
            b = getattr(full, name)[:, slab.z_slice].double()
            num = (a - b).pow(2).sum()
            dist.all_reduce(num)
            err = float((num / getattr(full, name).double().pow(2).sum()).sqrt())
            worst = max(worst, err)
            if rank == 0:
                print(f"periodic slab check {grid} x{world} ranks [{symbol}]: {name} rel-L2 vs single GPU = {err:.3e}")
        ok = ok and worst < 1e-5
        del slab, full
        torch.cuda.synchronize()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("PERIODIC SLAB CHECK", "OK" if int(flag.item()) else "FAILED")
    dist.destroy_process_group()
    return 0 if int(flag.item()) else 1


if __name__ == "__main__":
    sys.exit(main())
