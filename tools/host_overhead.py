"""Host-side cost of one coupled step through the public API: on a tiny grid the GPU work is negligible, so the
wall time per step is the Python / ctypes / launch overhead (python tools/host_overhead.py)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sopht_b200.simulator import UnboundedNavierStokesFlowSimulator3D
from sopht_b200.numeric.immersed_boundary_ops import VirtualBoundaryForcing
import cProfile, pstats

grid = (16, 16, 32)
sim = UnboundedNavierStokesFlowSimulator3D(grid_size=grid, x_range=1.0, kinematic_viscosity=1e-3, real_t=np.float32,
                                           with_forcing=True, with_free_stream_flow=True)
n = 64
pos = torch.rand(3, n, dtype=torch.float64, device="cuda") * 0.2 + 0.2
vel = torch.zeros_like(pos)
vb = VirtualBoundaryForcing(virtual_boundary_stiffness_coeff=-1.0, virtual_boundary_damping_coeff=-1.0, grid_dim=3,
                            dx=sim.dx, num_lag_nodes=n, real_t=np.float32)
dt = 1e-4
def step():
    vb.time_step(dt)
    vb.compute_interaction_force_on_eul_and_lag_grid(sim.eul_grid_forcing_field, sim.velocity_field, pos, vel)
    sim.time_step(dt=dt, free_stream_velocity=(1.0, 0.0, 0.0))
for _ in range(20):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
N = 300
for _ in range(N):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0) / N:.3f} ms/step, with final sync {1e3 * (t2 - t0) / N:.3f} ms/step")
pr = cProfile.Profile(); pr.enable()
for _ in range(100):
    step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
