"""Run a few unbounded Poisson vector solves on one grid (profiling helper: ncu -k regex:... python tools/poisson_only.py 512 512 512)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sopht_b200.numeric.eulerian_grid_ops import UnboundedPoissonSolverPYFFTW3D
if os.environ.get("SOPHT_L2_FETCH"):  # experiment: cudaLimitMaxL2FetchGranularity (32 / 64 / 128 bytes)
    import ctypes
    torch.cuda.init()
    torch.zeros(1, device="cuda")
    rt = ctypes.CDLL("libcudart.so.12")
    val = ctypes.c_size_t(0)
    rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(os.environ["SOPHT_L2_FETCH"])))
    rt.cudaDeviceGetLimit(ctypes.byref(val), 5)
    print(f"cudaLimitMaxL2FetchGranularity: rc={rc} now {val.value}")
nz, ny, nx = (int(a) for a in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
s = UnboundedPoissonSolverPYFFTW3D(nz, ny, nx, x_range=1.0, real_t=np.float32)
rhs = torch.randn(3, nz, ny, nx, device="cuda")
sol = torch.zeros_like(rhs)
for _ in range(2):
    s.vector_field_solve(sol, rhs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    s.vector_field_solve(sol, rhs)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
n = nz * ny * nx
from sopht_b200 import _lib
_lib.profile_enable(True)
for _ in range(reps):
    s.vector_field_solve(sol, rhs)
torch.cuda.synchronize()
rep = _lib.profile_report()
print("  " + "  ".join(f"{k.replace('poisson.', '')}={v['ms'] / reps:.3f}" for k, v in rep.items() if "nyq" not in k))
print(f"poisson {s.path} ({nz},{ny},{nx}): {ms:.3f} ms per vector solve, {324 * n / ms / 1e6:.0f} GB/s algorithmic (324 B/cell)")
