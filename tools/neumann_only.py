"""A few Neumann (FastDiag) vector solves on one grid (profiling helper, like tools/poisson_only.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sopht_b200.numeric.eulerian_grid_ops as spne
nz, ny, nx = (int(a) for a in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
s = spne.FastDiagPoissonSolver3D(nz, ny, nx, dx=1.0 / nx, real_t=np.float32)
rhs = torch.randn(3, nz, ny, nx, device="cuda")
sol = torch.zeros_like(rhs)
for _ in range(2):
    s.vector_field_solve(solution_vector_field=sol, rhs_vector_field=rhs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    s.vector_field_solve(solution_vector_field=sol, rhs_vector_field=rhs)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"neumann {s.path} ({nz},{ny},{nx}): {ms:.3f} ms per vector solve, {72 * 3 * nz * ny * nx / ms / 1e6:.0f} GB/s at 72 B/cell/component")
