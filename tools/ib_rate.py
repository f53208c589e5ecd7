"""Immersed-boundary transfer rates (nodes/s) on one grid: gather, spread (atomic-free tile gather vs red.global.add),
and the run-to-run reproducibility of the spread.  python tools/ib_rate.py [n_nodes ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from sopht_b200.numeric.immersed_boundary_ops import EulerianLagrangianGridCommunicator3D


def sphere(n, radius, centre):
    k = np.arange(n) + 0.5
    phi = np.arccos(1 - 2 * k / n)
    theta = np.pi * (1 + 5**0.5) * k
    return np.stack([centre[0] + radius * np.cos(theta) * np.sin(phi), centre[1] + radius * np.sin(theta) * np.sin(phi),
                     centre[2] + radius * np.cos(phi)])


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


mode = os.environ.get("SOPHT_IB_SPREAD", "tiles")
for n in [int(a) for a in sys.argv[1:]] or [2914, 100000]:
    # 2914 nodes: the sphere of BASELINE configs[1] on 128x128x256; 1e5 nodes: the same sphere on 512^3 (node spacing
    # ~1.1 cells)
    grid = (128, 128, 256) if n < 20000 else (512, 512, 512)
    dx = 1.0 / grid[2]
    centre = (0.25, 0.25, 0.25) if n < 20000 else (0.5, 0.5, 0.5)
    radius = 0.1 if n < 20000 else 0.19
    comm = EulerianLagrangianGridCommunicator3D(dx=np.float32(dx), eul_grid_coord_shift=np.float32(dx / 2),
                                                num_lag_nodes=n, interp_kernel_width=2, real_t=np.float32,
                                                n_components=3)
    pos = torch.from_numpy(sphere(n, radius, centre)).cuda()
    support = torch.zeros((3, 4, 4, 4, n), device="cuda")
    weights = torch.zeros((4, 4, 4, n), device="cuda")
    idx = torch.zeros((3, n), dtype=torch.int64, device="cuda")
    comm.local_eulerian_grid_support_of_lagrangian_grid_kernel(support, idx, pos)
    comm.interpolation_weights_kernel(weights, support)
    eul = torch.randn((3, *grid), device="cuda")
    lag = torch.randn((3, n), device="cuda")
    out = torch.zeros_like(lag)
    t_g = timed(lambda: comm.eulerian_to_lagrangian_grid_interpolation_kernel(out, eul, weights, idx))
    frc = torch.zeros_like(eul)
    t_s = timed(lambda: comm.lagrangian_to_eulerian_grid_interpolation_kernel(frc, lag, weights, idx))
    runs = []
    for _ in range(3):
        frc.zero_()
        comm.lagrangian_to_eulerian_grid_interpolation_kernel(frc, lag, weights, idx)
        runs.append(frc.clone())
    same = all(torch.equal(runs[0], r) for r in runs[1:])
    print(f"ib[{mode}] n={n} grid={grid}: gather {t_g * 1e3:.1f} us = {n / t_g / 1e6:.1f} Mnodes/s; "
          f"spread {t_s * 1e3:.1f} us = {n / t_s / 1e6:.1f} Mnodes/s; 3 spreads bit-identical: {same}")
