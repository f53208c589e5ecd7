"""Periodic Taylor-Green step (BASELINE config 4) on one GPU - NOT part of bench.py's contract (round 1 has one
256^3 line from it, profiles/r01_bench_periodic_256.json; the simulator is covered by tests/test_periodic_poisson.py).

    python tools/bench_periodic.py [--n 512] [--steps 10] [--warmup 3]

Prints one JSON line: Gcell-updates/s of PeriodicNavierStokesFlowSimulator3D (pass-by-pass kernels on halo-padded
fields) with the per-kernel event timers of the library, against SURVEY 8d's 180 B/cell for the periodic step."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sopht_b200 import _lib
from sopht_b200.simulator import PeriodicNavierStokesFlowSimulator3D


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    grid = (a.n, a.n, a.n)
    x_range = 2 * np.pi
    sim = PeriodicNavierStokesFlowSimulator3D(grid, x_range, kinematic_viscosity=1e-3, real_t=np.float32)
    x, y, z = (sim.position_field[i] for i in range(3))
    # vorticity of u = (sin x cos y cos z, -cos x sin y cos z, 0)
    sim.vorticity_field[0] = -torch.cos(x) * torch.sin(y) * torch.sin(z)
    sim.vorticity_field[1] = -torch.sin(x) * torch.cos(y) * torch.sin(z)
    sim.vorticity_field[2] = 2 * torch.sin(x) * torch.sin(y) * torch.cos(z)
    sim.compute_velocity_from_vorticity()
    dt = float(0.1 * sim.dx)
    for _ in range(a.warmup):
        sim.time_step(dt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        sim.time_step(dt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    _lib.profile_enable(True)
    for _ in range(a.steps):
        sim.time_step(dt)
    torch.cuda.synchronize()
    report = _lib.profile_report()
    _lib.profile_enable(False)
    cells = float(np.prod(grid))
    print(json.dumps({
        "metric": "3D periodic flow step Gcell-updates/s", "value": cells / ms / 1e6, "unit": "Gcell/s",
        "ms_per_step": ms, "grid": list(grid), "dtype": "f32", "algorithmic_bytes_per_cell": 180,
        "achieved_gbs": 180 * cells / ms / 1e6, "poisson_path": sim._poisson.path,
        "kernels": {k: {"ms_per_step": v["ms"] / a.steps, "launches_per_step": v["launches"] / a.steps}
                    for k, v in report.items()}}))


if __name__ == "__main__":
    main()
