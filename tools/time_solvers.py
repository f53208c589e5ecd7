"""Times the cuFFT-based Poisson variants (periodic, Neumann) on one GPU with CUDA events:
python tools/time_solvers.py  ->  ms per vector solve and GB/s against the algorithmic bytes of SURVEY 8d."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sopht_b200.numeric.eulerian_grid_ops as spne


def time_solver(solver, grid, reps=10):
    rhs = torch.randn((3, *grid), device="cuda")
    sol = torch.zeros_like(rhs)
    for _ in range(3):
        solver.vector_field_solve(solution_vector_field=sol, rhs_vector_field=rhs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        solver.vector_field_solve(solution_vector_field=sol, rhs_vector_field=rhs)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, grid, make, bytes_per_cell in (
    ("periodic spectral 512^3 fp32", (512, 512, 512),
     lambda g: spne.PeriodicPoissonSolver3D(*g, real_t=np.float32), 120),
    ("periodic spectral 256^3 fp32", (256, 256, 256),
     lambda g: spne.PeriodicPoissonSolver3D(*g, real_t=np.float32), 120),
    ("neumann (fast-diag closed form) 256^3 fp32", (256, 256, 256),
     lambda g: spne.FastDiagPoissonSolver3D(*g, dx=1.0 / g[2], real_t=np.float32), 120),
    ("unbounded pow2 256^3 fp32", (256, 256, 256),
     lambda g: spne.UnboundedPoissonSolverPYFFTW3D(*g, real_t=np.float32), 324),
):
    solver = make(grid)
    ms = time_solver(solver, grid)
    cells = float(np.prod(grid))
    print(f"{name}: {ms:.3f} ms / vector solve, {cells / ms / 1e6:.2f} Gcell/s, "
          f"{bytes_per_cell * cells / ms / 1e6:.0f} GB/s at {bytes_per_cell} B/cell algorithmic ({solver.path})", flush=True)
    del solver
    torch.cuda.empty_cache()
