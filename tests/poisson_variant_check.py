"""Parity of one pow2 Poisson vector solve against the CPU oracle, run as a script so that the kernel-selection
environment variables (read once per process by the library) can be set by the caller:

    SOPHT_P2_ZQUAD=0 python tests/poisson_variant_check.py 512 16 32

Prints `POISSON VARIANT OK <rel-L2>` or exits non-zero. Launched by tests/test_cuda_parity.py."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main() -> int:
    from oracle import poisson as opoisson
    from sopht_b200.numeric.eulerian_grid_ops import UnboundedPoissonSolverPYFFTW3D

    grid = tuple(int(a) for a in sys.argv[1:4])
    rng = np.random.default_rng(11)
    solver = UnboundedPoissonSolverPYFFTW3D(*grid, x_range=1.0, real_t=np.float32)
    assert solver.path == "pow2", solver.path
    ref = opoisson.UnboundedPoissonSolver3D(*grid, x_range=1.0, real_t=np.float32, workers=8)
    rhs = rng.standard_normal((3, *grid)).astype(np.float32)
    want = np.zeros_like(rhs)
    ref.vector_field_solve(want, rhs)
    got = torch.zeros(3, *grid, device="cuda")
    for _ in range(2):  # twice: persistent state of the handle (workspaces, barriers) must survive a solve
        solver.vector_field_solve(solution_vector_field=got, rhs_vector_field=torch.from_numpy(rhs).cuda())
    err = float(np.linalg.norm(got.cpu().numpy() - want) / np.linalg.norm(want))
    if not err < 1e-5:
        print(f"POISSON VARIANT FAILED rel-L2 {err:.3e}")
        return 1
    print(f"POISSON VARIANT OK {err:.3e}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
