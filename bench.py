#!/usr/bin/env python
"""bench.py — 3D unbounded flow step throughput (Gcell-updates/s) on N B200s, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Workloads (grid tuples are (nz, ny, nx) like the reference):
  u512    3D unbounded flow step 512^3 fp32 (no body): the north-star single-GPU size; weak-scaled with the rank
          count, so --gpus 8 runs 1024^3 (BASELINE configs[4])                                             [default]
  c1      2D flow past rigid cylinder, 512x256 fp32, 60 IB nodes (BASELINE configs[0]); latency-bound: the step is
          captured in a CUDA graph, reported as Gcell/s and steps/s
  c2      3D flow past rigid sphere, unbounded Poisson, 128x128x256 fp32, IB forcing (BASELINE configs[1])
  c3      3D Cosserat rod in cross-flow, 256x128x128 fp32, order-5 filter (BASELINE configs[2])
  u256    3D unbounded flow step 256^3 fp32 (no body)
  tg512 / tg256  3D periodic Taylor-Green vortex, 512^3 / 256^3 fp32 (BASELINE configs[3]; an extension: the reference
          has no periodic case, parity is against the numpy restatement + the analytic decay - "unpinned")
A "step" is one pass of the hot path: [IB gather + forcing + spread] -> vorticity update (rotational
advection + diffusion + boundary penalisation) -> unbounded FFT Poisson solve -> velocity = curl(psi) + U_inf.

`value`       device-resident fields, fixed dt, no host sync inside the timed region.
`e2e`         the loop a user of the reference writes (examples/3d_examples/FlowPastSphereCase/
              flow_past_sphere_case.py:191-197): every step copies that step's host inputs (Lagrangian
              body positions / velocities, pinned float64) to the device, runs the coupled step through
              the public simulator API, and reads the step's results (stable dt, Lagrangian forces) back.
`roofline`    dominant kernel, algorithmic bytes / CUDA-event time (a second K-step region with the
              library's per-phase event timers switched on), against MEASURED_PEAKS.json.
`cpu_baseline` the CPU oracle in its C / OpenMP form (one loop nest per reference kernel, unfused passes, + scipy.fft
              standing in for pyFFTW: the reference's dataflow) on the box's host cores, bounded sample, rank 0 at N=1.
`parity`      computed in this process: N=1 two coupled steps at 128x128x256 against the CPU oracle, N>1 the slab
              step against the single-GPU step on every rank's own planes (max rel-L2, tolerance 1e-5).
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    "c2": dict(grid=(128, 128, 256), body="sphere", desc="3D flow past rigid sphere, unbounded Poisson, 128x128x256 fp32"),
    # the rod is held straight and fixed: 40 elements x 16 surface points = 640 Lagrangian nodes,
    # order-5 convolution filter of the vorticity every step (flow_past_rod_case.py:94-121, :282-290)
    "c3": dict(grid=(256, 128, 128), body="rod", x_range=1.8, filter={"order": 5, "type": "convolution"},
               desc="3D Cosserat rod (straight, fixed) in cross-flow with IB forcing, 256x128x128 fp32, "
                    "order-5 convolution filter"),
    "c1": dict(grid=(256, 512), body="cylinder", two_d=True,
               desc="2D flow past rigid cylinder (Re=100), 512x256 grid, fp32"),
    "u256": dict(grid=(256, 256, 256), body=None, desc="3D unbounded flow step 256^3 fp32"),
    "u512": dict(grid=(512, 512, 512), body=None, desc="3D unbounded flow step 512^3 fp32"),
    "tg512": dict(grid=(512, 512, 512), body=None, periodic=True,
                  desc="3D periodic Taylor-Green vortex 512^3 fp32"),
    "tg256": dict(grid=(256, 256, 256), body=None, periodic=True,
                  desc="3D periodic Taylor-Green vortex 256^3 fp32"),
}
def global_grid(wl, world):
    """Weak scaling: the per-GPU cell count is fixed, the global grid grows with the number of ranks: x first, then y,
    then z, each up to 1024 (512^3 -> 512x512x1024 on 2, 512x1024x1024 on 4, 1024^3 on 8 GPUs = BASELINE configs[4]).
    The slab axis z is doubled LAST: its transform length 2 nz = 1024 is the two-pass radix-32 case the row-mode z
    kernel is built for, while the x transform (the cheapest pass) absorbs the growth first."""
    nz, ny, nx = wl["grid"]
    f = world
    while f > 1:
        if nx < 1024:
            nx *= 2
        elif ny < 1024:
            ny *= 2
        else:
            nz *= 2
        f //= 2
    return (nz, ny, nx)


NU = 1e-3


def fixed_dt(wl, dx):
    """The fixed time step of the timed loops: 0.1 dx for the body-free workloads; with an immersed body the
    virtual-boundary coupling (stiffness from the reference's examples) bounds it much lower - 0.02 dx for the sphere,
    0.005 dx for the rod (0.1 dx blows the coupled rod case up within six steps; the work per step does not depend on
    dt, a run on NaNs would still not be a measurement)."""
    return float({"sphere": 0.02, "rod": 0.005}.get(wl.get("body"), 0.1) * dx)


CPU_KIND_DESC = ("C/OpenMP loop nest per reference kernel (oracle/c/ref_kernels.c, unfused passes like pystencils) + "
                 "scipy.fft in place of pyFFTW")
X_RANGE = 1.0
U_INF = (1.0, 0.0, 0.0)


def algorithmic_bytes_per_cell(with_forcing: bool, with_filter: bool = False) -> int:
    """SURVEY.md §8(d) / BASELINE.md §2: 3D unbounded NS step, fp32 (+72 B for the per-direction fused filter)."""
    return (408 if with_forcing else 384) + (72 if with_filter else 0)


# per-launch algorithmic bytes per cell (fp32) of the instrumented kernels: one read of every distinct input
# and one write of every output (DESIGN.md "Kernels"); Poisson phases count all three components.
KERNEL_BYTES_PER_CELL = {
    "ns3d.advect": 36.0,             # read w(3) + u(3), write w'(3)
    "ns3d.diffuse": 24.0,            # read 3, write 3 (+12 when it also zeroes the forcing field)
    "ns3d.velocity": 24.0,           # read psi(3), write u(3)
    "poisson.x_fwd": 36.0,           # 3 x (4 real in + 8 half-spectrum out)
    "poisson.y_fwd": 72.0,           # 3 x (8 in + 16 out, y doubled)
    "poisson.z_conv": 100.0,         # 3 x (16 in + 16 out, in place) + 4 (folded real G_hat, read once)
    "poisson.y_inv": 72.0,
    "poisson.x_inv": 36.0,
    "update_vorticity_from_velocity_forcing": 36.0,  # read f(3) + w(3), write w(3)
    "laplacian_filter.x": 8.0,       # per component launch: one read + one write of a scalar field
    "laplacian_filter.y": 8.0,
    "laplacian_filter.z": 8.0,
}


def ncu_traffic(kernel, grid):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` on `grid`, from the committed
    `ncu --set full` capture summarised in profiles/ncu_traffic.json (None if that pair was never captured)."""
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return table["x".join(str(g) for g in grid)][kernel]["dram_bytes_per_launch"]
    except Exception:  # noqa: BLE001
        return None


def roofline_from_report(report, steps, cells, forcing, peak, peak_src, grid, side_stream):
    """Per-kernel achieved GB/s from the library's event timers; the roofline object describes the kernel
    with the largest share of the step."""
    kernels = {}
    # the Nyquist-plane kernels run on the solver's side stream underneath the main y / z kernels: their event
    # time is mostly waiting for SM slots and overlaps the main stream, so it is not part of the step's sum
    # (single-GPU solver, grids up to 2^25 cells; elsewhere they run on the main stream and count like any kernel)
    overlapped = {k for k in report if k.endswith(".nyquist")} if side_stream else set()
    total = sum(v["ms"] for k, v in report.items() if k not in overlapped) or 1.0
    for label, v in report.items():
        per_launch_ms = v["ms"] / max(v["launches"], 1)
        bpc = KERNEL_BYTES_PER_CELL.get(label)
        if label == "ns3d.diffuse" and forcing:
            bpc = 36.0
        entry = {"launches_per_step": v["launches"] / steps, "ms_per_step": v["ms"] / steps,
                 "share": None if label in overlapped else v["ms"] / total}
        if label in overlapped:
            entry["overlapped"] = "side stream, concurrent with the main y/z kernels"
        if bpc is not None:
            entry["algorithmic_bytes_per_cell"] = bpc
            entry["achieved_gbs"] = bpc * cells * (v["launches"] / steps) / (v["ms"] / steps * 1e-3) / 1e9 \
                if v["ms"] > 0 else None
            entry["frac"] = entry["achieved_gbs"] / peak if entry["achieved_gbs"] else None
        entry["avg_launch_ms"] = per_launch_ms
        kernels[label] = entry
    dom = max((k for k in kernels if "achieved_gbs" in kernels[k]), key=lambda k: kernels[k]["ms_per_step"])
    d = kernels[dom]
    roof = {"bound": "hbm", "achieved": d["achieved_gbs"], "peak": peak, "unit": "GB/s",
            "frac": d["frac"], "traffic": ncu_traffic(dom, grid), "kernel": dom,
            "algorithmic_bytes_per_launch": d["algorithmic_bytes_per_cell"] * cells,
            "avg_launch_ms": d["avg_launch_ms"], "share_of_step": d["share"], "peak_source": peak_src}
    return roof, kernels


def hill_vortex_vorticity(grid, x_range, real_t=np.float32, z_planes=None):
    """Smooth band-limited initial vorticity: Hill's spherical vortex (R = 0.25 x_range, U = 1) centred
    in the domain (analogue of examples/3d_examples/HillSphericalVortexCase)."""
    nz, ny, nx = grid
    dx = x_range / nx
    z = (np.arange(nz) + 0.5) * dx
    y = (np.arange(ny) + 0.5) * dx
    x = (np.arange(nx) + 0.5) * dx
    zc, yc, xc = z.mean(), y.mean(), x.mean()
    if z_planes is not None:  # only these global planes (a rank's slab)
        z = z[z_planes[0]:z_planes[1]]
    nz = len(z)
    Z, Y, X = np.meshgrid(z - zc, y - yc, x - xc, indexing="ij")
    R = 0.25 * min(grid) * dx
    r2 = X * X + Y * Y + Z * Z
    inside = r2 <= R * R
    # omega = (15 U / 2 R^2) * rho * e_phi about the z axis
    pref = 7.5 / (R * R)
    w = np.zeros((3, nz, ny, nx), dtype=real_t)
    w[0] = np.where(inside, -pref * Y, 0.0)
    w[1] = np.where(inside, pref * X, 0.0)
    return w


def sphere_lag_grid(n_eq=96, diameter=0.2, centre=(0.25, 0.25, 0.25)):
    """Forcing points on a sphere surface, equal-area latitude rings
    (rigid_body_forcing_grids.py:236-300 analogue; ~2914 points for n_eq = 96)."""
    r = diameter / 2
    n_lat = n_eq // 2
    pts = []
    for i in range(n_lat + 1):
        polar = np.pi * i / n_lat
        n_ring = max(1, int(round(n_eq * np.sin(polar))))
        az = 2 * np.pi * (np.arange(n_ring) + 0.5 * (i % 2)) / n_ring
        pts.append(np.stack([r * np.sin(polar) * np.cos(az) + centre[0],
                             r * np.sin(polar) * np.sin(az) + centre[1],
                             np.full(n_ring, r * np.cos(polar)) + centre[2]]))
    return np.concatenate(pts, axis=1)  # (3, N) float64


def straight_rod(wl):
    """The rod of flow_past_rod_case.py:44-48 at rest: 40 elements along -z from (0.2 X, 0.5 Y, 0.75 Z), length 1,
    diameter Y / 5 (pyelastica attribute names; the elastic solve itself is outside the hot path)."""
    from sopht_b200.simulator import CosseratRodState

    nz, ny, nx = wl["grid"]
    x_range = wl.get("x_range", X_RANGE)
    y_range, z_range = ny / nx * x_range, nz / nx * x_range
    return CosseratRodState.straight_rod(
        n_elements=5 * nx // 16, start=np.array([0.2 * x_range, 0.5 * y_range, 0.75 * z_range]),
        direction=np.array([0.0, 0.0, -1.0]), normal=np.array([0.0, 1.0, 0.0]), base_length=1.0,
        base_radius=y_range / 10.0)


class ClockSampler:
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.1) -> None:
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self._nv = None
        self._period = period

    def _run(self) -> None:
        nv = self._nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(self._period)

    def __enter__(self):
        if self._nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self) -> dict:
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle's restatement of the reference step (numpy stencils + scipy.fft + IB loops)
# ---------------------------------------------------------------------------------------------------------
def build_cpu_case(wl, cores):
    from oracle import cstencils
    from oracle import flow as oflow

    cstencils.load()
    cstencils.set_num_threads(cores)
    grid = wl["grid"]
    forcing = wl["body"] is not None
    x_range = wl.get("x_range", X_RANGE)
    filt = wl.get("filter")
    sim = oflow.UnboundedNavierStokesFlowSimulator3D(
        grid_size=grid, x_range=x_range, kinematic_viscosity=NU, real_t=np.float32, kernels=cstencils,
        with_forcing=forcing, with_free_stream_flow=True, workers=cores, filter_vorticity=filt is not None,
        **({"filter_setting_dict": filt} if filt else {}))
    sim.vorticity_field[...] = hill_vortex_vorticity(grid, x_range)
    sim._poisson.vector_field_solve(sim.stream_func_field, sim.vorticity_field)
    from oracle import stencils as ost

    ost.curl_3d(sim.velocity_field, sim.stream_func_field, np.float32(0.5 / sim.dx))
    vb = None
    if forcing:
        from oracle import ib as oib

        if wl["body"] == "rod":
            from oracle import forcing_grids as ofg

            rod = straight_rod(wl)
            points, ratio, angles = ofg.rod_surface_layout(rod, wl["grid"][2] // 8)
            pos, _, _ = ofg.rod_surface_kinematics(rod, points, ratio, ofg.rod_surface_tables(points, angles)[2])
            ds = max(np.amax(rod.lengths), np.amax(rod.radius) * 2 * np.pi / (wl["grid"][2] // 8))
            k, c = -2e4, -1e2  # flow_past_rod_case.py:19-20
        else:
            pos = sphere_lag_grid()
            ds = np.pi * 0.2 / 96  # ~ lagrangian spacing
            k, c = -1.5e5, -87.5
        vb = (oib.VirtualBoundaryForcing(k * ds * ds, c * ds * ds, 3, sim.dx, pos.shape[1], np.float32),
              pos, np.zeros_like(pos))
    return sim, vb


def cpu_step(sim, vb, dt):
    if vb is not None:
        f, pos, vel = vb
        f.time_step(dt)
        f.compute_interaction_force_on_eul_and_lag_grid(
            sim.eul_grid_forcing_field, sim.velocity_field, pos, vel)
    sim.time_step(dt, free_stream_velocity=U_INF)


def cpu_sample_grid(grid, max_cells=2**24):
    """Bounded CPU sample of a workload: the same flow on a grid halved (largest axis first) until it has at most
    `max_cells` cells. Per-cell work is identical, so the throughput unit (Gcell/s) carries over."""
    g = list(grid)
    while int(np.prod(g)) > max_cells:
        g[int(np.argmax(g))] //= 2
    return tuple(g)


def time_cpu(wl, steps, warmup):
    cores = len(os.sched_getaffinity(0))
    sim, vb = build_cpu_case(wl, cores)
    dt = fixed_dt(wl, sim.dx)
    for _ in range(warmup):
        cpu_step(sim, vb, dt)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(sim, vb, dt)
    el = time.perf_counter() - t0
    cells = int(np.prod(wl["grid"]))
    return cells * steps / el / 1e9, el / steps * 1e3, cores


def workload_config(wl, grid, world, n_lag):
    """The `config` object of the JSON line: names the workload only, identical for both arms."""
    cells_local = int(np.prod(grid)) // world
    return {"workload": wl["desc"] + (f", weak-scaled to {world} z-slabs" if world > 1 else ""),
            "grid": list(grid), "cells_per_gpu": cells_local,
            "parallelism": f"z-slab x{world}" if world > 1 else "single GPU",
            "lagrangian_nodes": n_lag,
            "l2": "inputs larger than L2: per-GPU working set per step (fields + FFT workspace, "
                  f"{cells_local * 4 * 33 / 1e6:.0f} MB) exceeds the 126 MB L2"}


def workload_lag_nodes(wl):
    if wl["body"] == "sphere":
        return int(sphere_lag_grid().shape[1])
    if wl["body"] == "rod":
        n_elem = 5 * wl["grid"][2] // 16
        return n_elem * (wl["grid"][2] // 8)
    return 0


def run_reference_arm(args, wl):
    """CPU arm: the oracle's restatement of the reference step on the host cores, same metric / config / steps /
    warm-up as the GPU arm; every step is one pass over a bounded sample of the workload (cpu_sample_grid)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(args.gpus, 1)
    grid = global_grid(wl, world)
    sample_grid = cpu_sample_grid(grid)
    steps, warm = args.steps, args.warmup
    val, ms, cores = time_cpu(dict(wl, grid=sample_grid), steps, warm)
    sample = (f"{steps} steps (+{warm} warm-up) of the same flow step on a {sample_grid[0]}x{sample_grid[1]}x"
              f"{sample_grid[2]} grid" + ("" if sample_grid == tuple(grid) else " (bounded sample of the workload grid)")
              + f"; {CPU_KIND_DESC}, {cores} threads, {ms:.0f} ms/step")
    line = {
        "impl": "reference", "metric": "3D flow step Gcell-updates/s", "value": val, "unit": "Gcell/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(wl, grid, world, workload_lag_nodes(wl)),
        "timing": "host perf_counter (CPU arm)",
        "cpu_baseline": {"value": val, "unit": "Gcell/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Gcell/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def parity_single_gpu():
    """N = 1: two coupled steps at BASELINE configs[1] size (128x128x256, forcing + free stream) from a seeded random
    state through the same fused CUDA step the timed loop uses, against the CPU oracle; worst relative L2 error."""
    import torch

    from oracle import flow as oflow
    from sopht_b200.simulator import UnboundedNavierStokesFlowSimulator3D

    grid = (128, 128, 256)
    kw = dict(grid_size=grid, x_range=1.0, kinematic_viscosity=NU, real_t=np.float32, with_forcing=True,
              with_free_stream_flow=True)
    sim = UnboundedNavierStokesFlowSimulator3D(**kw)
    ref = oflow.UnboundedNavierStokesFlowSimulator3D(workers=len(os.sched_getaffinity(0)), **kw)
    rng = np.random.default_rng(2024)
    for name in ("vorticity_field", "velocity_field", "eul_grid_forcing_field"):
        a = rng.standard_normal((3, *grid)).astype(np.float32)
        getattr(ref, name)[...] = a
        getattr(sim, name)[...] = torch.from_numpy(a).cuda()
    dt = float(ref.compute_stable_timestep(dt_prefac=0.5))
    worst = abs(float(sim.compute_stable_timestep(dt_prefac=0.5)) - dt) / dt
    for _ in range(2):
        sim.time_step(dt=dt, free_stream_velocity=U_INF)
        ref.time_step(dt, free_stream_velocity=U_INF)
    errs = {}
    for name in ("vorticity_field", "velocity_field", "stream_func_field"):
        a = getattr(sim, name).cpu().numpy().astype(np.float64)
        b = getattr(ref, name).astype(np.float64)
        errs[name] = float(np.linalg.norm(a - b) / np.linalg.norm(b))
    worst = max(worst, *errs.values())
    del sim
    torch.cuda.empty_cache()
    return {"value": worst, "metric": "max rel-L2 (vorticity, velocity, stream function, dt)", "tolerance": 1e-5,
            "ok": bool(worst < 1e-5), "against": "CPU oracle (oracle/flow.py), 2 coupled steps, 128x128x256 fp32, "
            "seeded random state, forcing + free stream", "fields": errs}


def parity_slab(world):
    """N > 1: the z-slab decomposed step (peer-memory halos, NVLink transposes) against the single-GPU step of the
    same library on the whole grid, every rank on its own planes; worst relative L2 error over the ranks."""
    import torch
    import torch.distributed as dist

    from sopht_b200.parallel import SlabUnboundedNavierStokesFlowSimulator3D
    from sopht_b200.simulator import UnboundedNavierStokesFlowSimulator3D

    grid = (128, 64, 128)
    kw = dict(grid_size=grid, x_range=1.0, kinematic_viscosity=NU, real_t=np.float32, with_forcing=True,
              with_free_stream_flow=True)
    slab = SlabUnboundedNavierStokesFlowSimulator3D(**kw)
    full = UnboundedNavierStokesFlowSimulator3D(**kw)
    rng = np.random.default_rng(2025)
    for name in ("vorticity_field", "velocity_field", "eul_grid_forcing_field"):
        a = rng.standard_normal((3, *grid)).astype(np.float32)
        getattr(full, name)[...] = torch.from_numpy(a).cuda()
        slab.set_owned(getattr(slab, name), a)
    dt = float(full.compute_stable_timestep(dt_prefac=0.5))
    worst = abs(float(slab.compute_stable_timestep(dt_prefac=0.5)) - dt) / dt
    for _ in range(2):
        full.time_step(dt=dt, free_stream_velocity=U_INF)
        slab.time_step(dt=dt, free_stream_velocity=U_INF)
    errs, norms = {}, {}
    for name in ("vorticity_field", "velocity_field", "stream_func_field"):
        a = slab.owned(getattr(slab, name)).double()
        b = getattr(full, name)[:, slab.z_slice].double()
        num = (a - b).pow(2).sum()
        dist.all_reduce(num)
        den = getattr(full, name).double().pow(2).sum()
        if not (float(den) > 0.0 and np.isfinite(float(den)) and np.isfinite(float(num))):
            raise RuntimeError(f"parity check: degenerate {name} (norm^2 = {float(den)})")
        errs[name] = float((num / den).sqrt())
        norms[name] = float(den.sqrt())
    worst = max(worst, *errs.values())
    torch.cuda.synchronize()
    del slab, full
    torch.cuda.empty_cache()
    return {"value": worst, "metric": "max rel-L2 (vorticity, velocity, stream function, dt)", "tolerance": 1e-5,
            "ok": bool(worst < 1e-5), "against": f"single-GPU step of the same library on the whole 128x64x128 grid "
            f"(itself pinned to the CPU oracle at N = 1), {world} z-slabs, 2 steps, seeded random state; 0.0 = bit-identical "
            "(same kernels, same transform lengths, no atomics on this path)", "fields": errs, "reference_l2_norms": norms}


def guarded(fn, *a, limit_s=240.0):
    """Run a parity check; a hang (a peer that never answers) must not take the bench line with it."""
    box = {}
    import torch

    device = torch.cuda.current_device()  # the current device is thread-local: the worker must select it itself

    def work():
        try:
            torch.cuda.set_device(device)
            box["r"] = fn(*a)
        except Exception as e:  # noqa: BLE001
            box["r"] = {"value": None, "ok": False, "error": f"{type(e).__name__}: {e}"[:300]}

    th = threading.Thread(target=work, daemon=True)
    th.start()
    th.join(limit_s)
    return box.get("r", {"value": None, "ok": False, "error": f"parity check did not finish in {limit_s:.0f} s"})


def run_ours(args, wl):
    import torch
    import torch.distributed as dist

    from sopht_b200 import _lib
    from sopht_b200.simulator import UnboundedNavierStokesFlowSimulator3D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.load()

    grid = global_grid(wl, world)
    forcing = wl["body"] is not None
    x_range = wl.get("x_range", X_RANGE)
    filt = wl.get("filter")
    parity = None
    if not args.no_parity:
        parity = guarded(parity_single_gpu) if world == 1 else guarded(parity_slab, world)
    if world > 1 and (filt or wl["body"] == "rod"):
        raise SystemExit("workload c3 (rod + filter) is a single-GPU bench line; use c2 / u256 / u512 with --gpus N")
    dt_value = None
    if world == 1:
        sim = UnboundedNavierStokesFlowSimulator3D(
            grid_size=grid, x_range=x_range, kinematic_viscosity=NU, real_t=np.float32,
            with_forcing=forcing, with_free_stream_flow=True, step_mode=args.step_mode,
            filter_vorticity=filt is not None, **({"filter_setting_dict": filt} if filt else {}))
        sim.vorticity_field[...] = torch.from_numpy(hill_vortex_vorticity(grid, x_range)).cuda()
        sim._unbounded_poisson_solver.vector_field_solve(
            solution_vector_field=sim.stream_func_field, rhs_vector_field=sim.vorticity_field)
        sim._curl(curl=sim.velocity_field, field=sim.stream_func_field, prefactor=np.float32(0.5 / sim.dx))
        own = lambda f: f  # noqa: E731
        cells_local = int(np.prod(grid))
    else:
        from sopht_b200.parallel import SlabUnboundedNavierStokesFlowSimulator3D

        sim = SlabUnboundedNavierStokesFlowSimulator3D(
            grid_size=grid, x_range=X_RANGE, kinematic_viscosity=NU, real_t=np.float32,
            with_forcing=forcing, with_free_stream_flow=True)
        # this rank's planes of the same initial vorticity; one zero-dt-free way to get u0: a first step
        z0, nzl = sim.part.z_start, sim.part.nz_local
        w0 = hill_vortex_vorticity(grid, X_RANGE, z_planes=(max(z0 - 1, 0), min(z0 + nzl + 1, grid[0])))
        lo = 1 - (z0 - max(z0 - 1, 0))
        sim.vorticity_field[:, lo:lo + w0.shape[1]] = torch.from_numpy(w0).cuda()
        sim._unbounded_poisson_solver.vector_field_solve(
            solution_vector_field=sim.owned(sim.stream_func_field), rhs_vector_field=sim.owned(sim.vorticity_field))
        sim._halos(sim.stream_func_field)
        lib = _lib.load()
        fu = _lib.field_desc(sim.part.stencil_view(sim.velocity_field), 0)
        fpsi = _lib.field_desc(sim.part.stencil_view(sim.stream_func_field), 0)
        import ctypes

        _lib.check(lib.sopht_ns3d_velocity_from_stream_function(
            0, ctypes.byref(fu), ctypes.byref(fpsi), float(np.float32(0.5 / sim.dx)), None, None,
            _lib.current_stream()))
        own = sim.owned
        cells_local = int(np.prod(grid)) // world
    dt = fixed_dt(wl, sim.dx)
    cells = int(np.prod(grid))

    interactor = None
    rod_interactor = None
    if wl["body"] == "rod":
        # the reference's own object graph: CosseratRodFlowInteraction owning a surface forcing grid on the device
        from sopht_b200.simulator import CosseratRodFlowInteraction, CosseratRodSurfaceForcingGrid

        rod = straight_rod(wl)
        rod_interactor = CosseratRodFlowInteraction(
            cosserat_rod=rod, eul_grid_forcing_field=sim.eul_grid_forcing_field,
            eul_grid_velocity_field=sim.velocity_field, virtual_boundary_stiffness_coeff=-2e4,
            virtual_boundary_damping_coeff=-1e2, dx=sim.dx, grid_dim=3, real_t=np.float32,
            forcing_grid_cls=CosseratRodSurfaceForcingGrid,
            surface_grid_density_for_largest_element=grid[2] // 8)
        n_lag = rod_interactor.forcing_grid.num_lag_nodes
        rod_state_bytes = rod_interactor.forcing_grid._state_host.numel() * 8
    elif forcing:
        pos_h = torch.from_numpy(sphere_lag_grid()).pin_memory()
        vel_h = torch.zeros_like(pos_h).pin_memory()
        ds = np.pi * 0.2 / 96
        vb_kw = dict(virtual_boundary_stiffness_coeff=-1.5e5 * ds * ds, virtual_boundary_damping_coeff=-87.5 * ds * ds,
                     grid_dim=3, dx=sim.dx, num_lag_nodes=pos_h.shape[1], real_t=np.float32)
        if world == 1:
            from sopht_b200.numeric.immersed_boundary_ops import VirtualBoundaryForcing

            interactor = VirtualBoundaryForcing(**vb_kw)
        else:
            from sopht_b200.parallel import SlabVirtualBoundaryForcing

            interactor = SlabVirtualBoundaryForcing(**vb_kw, partition=sim.part)
        pos_d, vel_d = pos_h.cuda(), vel_h.cuda()
        force_h = torch.zeros(3, pos_h.shape[1], dtype=torch.float32).pin_memory()

    def device_step():
        if rod_interactor is not None:
            rod_interactor.time_step(dt)
            rod_interactor()  # rod state H2D (7 KB), grid kinematics, interpolate / force / spread
        if interactor is not None:
            interactor.time_step(dt)
            interactor.compute_interaction_force_on_eul_and_lag_grid(
                own(sim.eul_grid_forcing_field), own(sim.velocity_field), pos_d, vel_d)
        sim.time_step(dt=dt, free_stream_velocity=U_INF)

    h2d = d2h = 0

    def e2e_step():
        nonlocal h2d, d2h
        step_dt = sim.compute_stable_timestep(dt_prefac=0.5)  # device reduction + D2H scalar
        d2h_n = 4
        h2d_n = 0
        if rod_interactor is not None:
            # the coupled loop of flow_past_rod_case.py:236-251 with one rod sub-step per flow step: the rod's
            # forcing hook reads the flow forces / torques back, then the flow interaction spreads the forcing
            rod_interactor.compute_flow_forces_and_torques()
            rod_interactor.time_step(step_dt)
            rod_interactor()
            h2d_n += 5 * rod_state_bytes  # every grid update / transfer uploads the packed rod state
            d2h_n += rod_interactor.forcing_grid._out_host.numel() * 8
        if interactor is not None:
            p = pos_h.to("cuda", non_blocking=True)
            v = vel_h.to("cuda", non_blocking=True)
            h2d_n += pos_h.numel() * 8 + vel_h.numel() * 8
            interactor.time_step(step_dt)
            interactor.compute_interaction_force_on_eul_and_lag_grid(
                own(sim.eul_grid_forcing_field), own(sim.velocity_field), p, v)
            force_h.copy_(interactor.lag_grid_forcing_field, non_blocking=True)
            d2h_n += force_h.numel() * 4
        sim.time_step(dt=step_dt, free_stream_velocity=U_INF)
        h2d, d2h = h2d_n, d2h_n

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(max(args.warmup, 3)):
        device_step()
    # small grids: the fixed-dt step (interaction + flow step) replayed from a CUDA graph - the launches of a step
    # are issued more slowly by the host than the device executes them
    use_graph = world == 1 and (args.graph == "on" or (args.graph == "auto" and cells <= 2**24))
    timed_step = device_step
    if use_graph:
        def ib_part():
            if rod_interactor is not None:
                rod_interactor.time_step(dt)
                rod_interactor()
            if interactor is not None:
                interactor.time_step(dt)
                interactor.compute_interaction_force_on_eul_and_lag_grid(
                    own(sim.eul_grid_forcing_field), own(sim.velocity_field), pos_d, vel_d)

        ib_state = []
        for obj in (interactor, getattr(rod_interactor, "forcing_grid", None), rod_interactor):
            ib_state += [v for v in vars(obj).values() if isinstance(v, torch.Tensor)] if obj is not None else []
        try:
            timed_step = sim.graph_time_step(dt, before=ib_part if forcing else None, extra_state=ib_state,
                                             free_stream_velocity=U_INF)
            for _ in range(3):
                timed_step()
        except Exception as e:  # noqa: BLE001 - a step that cannot be captured is measured eagerly
            print(f"bench: CUDA graph capture failed ({type(e).__name__}: {e}); eager launches", file=sys.stderr)
            timed_step, use_graph = device_step, False
    n0 = _lib.launch_count()
    with ClockSampler(local) as clk:
        ms = timed(timed_step, args.steps)
    launches = _lib.launch_count() - n0
    value = cells * args.steps / (ms * 1e-3) / 1e9

    # end-to-end through the public API with host-side per-step inputs/outputs
    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    e2e_val = cells * args.steps / (ms_e2e * 1e-3) / 1e9

    # roofline of the dominant kernel: the library's per-kernel CUDA-event timers (recorded on the launching
    # stream) switched on for a second K-step region of the same loop
    peak, peak_src = measured_peak_hbm()
    _lib.profile_enable(True)
    barrier()
    for _ in range(args.steps):
        device_step()
    barrier()
    report = _lib.profile_report()
    _lib.profile_enable(False)
    roof, kernels = roofline_from_report(report, args.steps, cells_local, forcing, peak, peak_src, grid,
                                         side_stream=world == 1 and cells <= 2**25)
    step_bpc = algorithmic_bytes_per_cell(forcing, filt is not None)
    whole = step_bpc * cells_local * args.steps / (ms * 1e-3) / 1e9  # per GPU

    if rank != 0:
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sample_grid = cpu_sample_grid(grid)
        csteps = 8 if int(np.prod(sample_grid)) <= 2**23 else 4
        cval, cms, cores = time_cpu(dict(wl, grid=sample_grid), csteps, 1)
        cpu = {"value": cval, "unit": "Gcell/s", "cores": cores, "kind": "port",
               "sample": f"{csteps} steps (+1 warm-up) of the same flow step on a {sample_grid[0]}x{sample_grid[1]}x"
                         f"{sample_grid[2]} grid" + ("" if sample_grid == tuple(grid) else
                                                     " (bounded sample of the workload grid)")
                         + f"; {CPU_KIND_DESC}, {cores} threads, {cms:.0f} ms/step"}
    line = {
        "metric": "3D flow step Gcell-updates/s", "value": value, "unit": "Gcell/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(
            wl, grid, world, (n_lag if rod_interactor is not None else int(pos_h.shape[1])) if forcing else 0),
        "path": {"step_mode": sim.step_mode, "poisson_path": sim._unbounded_poisson_solver.path,
                 "cuda_graph": bool(use_graph)},
        "parity": parity,
        "e2e": {"value": e2e_val, "unit": "Gcell/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "note": "public-API loop of the reference's examples: stable-dt read-back every step (device reduction "
                        "+ D2H scalar), then the coupled step; body workloads also upload the Lagrangian positions / "
                        "velocities from pinned host memory and read the Lagrangian forces back"},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
        "roofline": roof,
        "step_roofline": {"algorithmic_bytes_per_cell": step_bpc, "achieved": whole, "unit": "GB/s",
                          "frac": whole / peak},
        "kernels": kernels,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------
# periodic Taylor-Green workload (BASELINE configs[3]); single GPU
# ---------------------------------------------------------------------------------------------------------
PERIODIC_BYTES_PER_CELL = 180  # stencils 60 + periodic Poisson 3 x 40
PERIODIC_KERNEL_BYTES = {"ns3d.advect": 36.0, "ns3d.diffuse": 24.0, "ns3d.velocity": 24.0}


def taylor_green_vorticity(grid, real_t=np.float32, z_offset=0, nx_global=None):
    """Vorticity of u = (sin x cos y cos z, -cos x sin y cos z, 0) on [0, 2 pi)^3 mapped onto the unit box
    (grid: the planes to produce, starting at global plane z_offset)."""
    nz, ny, nx = grid
    nx = nx_global or nx
    k = 2 * np.pi
    z = (np.arange(nz) + z_offset + 0.5) / nx * k
    y = (np.arange(ny) + 0.5) / nx * k
    x = (np.arange(nx) + 0.5) / nx * k
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
    w = np.empty((3, nz, ny, nx), dtype=real_t)
    w[0] = -k * np.cos(X) * np.sin(Y) * np.sin(Z)
    w[1] = -k * np.sin(X) * np.cos(Y) * np.sin(Z)
    w[2] = 2 * k * np.sin(X) * np.sin(Y) * np.cos(Z)
    return w


def parity_periodic():
    """Fused periodic step against the numpy restatement (oracle/flow.py, itself checked against the analytic
    Taylor-Green decay; no reference code exists for this case: unpinned), 3 steps at 32x64x128."""
    import torch

    from oracle import flow as oflow
    from sopht_b200.simulator import PeriodicNavierStokesFlowSimulator3D

    grid = (32, 64, 128)
    sim = PeriodicNavierStokesFlowSimulator3D(grid, 1.0, NU, real_t=np.float32)
    ref = oflow.PeriodicNavierStokesFlowSimulator3D(grid, 1.0, NU, real_t=np.float32)
    w0 = np.random.default_rng(7).standard_normal((3, *grid)).astype(np.float32)
    sim.vorticity_field[...] = torch.from_numpy(w0).cuda()
    ref.vorticity_field[...] = w0
    sim.compute_velocity_from_vorticity()
    ref.compute_velocity_from_vorticity()
    dt = float(ref.compute_stable_timestep(0.5))
    for _ in range(3):
        sim.time_step(dt)
        ref.time_step(dt)
    errs = {}
    for name in ("vorticity_field", "velocity_field"):
        a = getattr(sim, name).cpu().numpy().astype(np.float64)
        b = np.asarray(getattr(ref, name), dtype=np.float64)
        errs[name] = float(np.linalg.norm(a - b) / np.linalg.norm(b))
    worst = max(errs.values())
    return {"value": worst, "metric": "max rel-L2 (vorticity, velocity)", "tolerance": 1e-5, "ok": bool(worst < 1e-5),
            "against": "numpy restatement of the periodic step (oracle/flow.py), 3 steps, 32x64x128 fp32, seeded "
                       "random state; UNPINNED: the reference has no periodic code", "fields": errs}


def parity_periodic_slab(world):
    """N > 1: the slab-decomposed periodic step (ring halo exchange, NVLink transposes) against the single-GPU periodic
    step of the same library on the whole grid, every rank on its own planes."""
    import torch
    import torch.distributed as dist

    from sopht_b200.parallel import SlabPeriodicNavierStokesFlowSimulator3D
    from sopht_b200.simulator import PeriodicNavierStokesFlowSimulator3D

    grid = (128, 64, 256)
    slab = SlabPeriodicNavierStokesFlowSimulator3D(grid, 1.0, NU)
    full = PeriodicNavierStokesFlowSimulator3D(grid, 1.0, NU, real_t=np.float32)
    w0 = np.random.default_rng(7).standard_normal((3, *grid)).astype(np.float32)
    full.vorticity_field[...] = torch.from_numpy(w0).cuda()
    slab.set_owned(slab.vorticity_field, w0)
    full.compute_velocity_from_vorticity()
    slab.compute_velocity_from_vorticity()
    dt = float(full.compute_stable_timestep(0.5))
    for _ in range(3):
        full.time_step(dt)
        slab.time_step(dt)
    errs, norms = {}, {}
    for name in ("vorticity_field", "velocity_field", "stream_func_field"):
        a = slab.owned(getattr(slab, name)).double()
        b = getattr(full, name)[:, slab.z_slice].double()
        num = (a - b).pow(2).sum()
        dist.all_reduce(num)
        den = getattr(full, name).double().pow(2).sum()
        if not (float(den) > 0.0 and np.isfinite(float(den)) and np.isfinite(float(num))):
            raise RuntimeError(f"parity check: degenerate {name} (norm^2 = {float(den)})")
        errs[name] = float((num / den).sqrt())
        norms[name] = float(den.sqrt())
    worst = max(errs.values())
    torch.cuda.synchronize()
    del slab, full
    torch.cuda.empty_cache()
    return {"value": worst, "metric": "max rel-L2 (vorticity, velocity, stream function)", "tolerance": 1e-5,
            "ok": bool(worst < 1e-5), "against": f"single-GPU periodic step of the same library on the whole 128x64x256 "
            f"grid, {world} z-slabs, 3 steps (UNPINNED by the reference: it has no periodic code); 0.0 = bit-identical",
            "fields": errs, "reference_l2_norms": norms}


def run_periodic(args, wl):
    import torch

    from sopht_b200 import _lib
    from sopht_b200.simulator import PeriodicNavierStokesFlowSimulator3D

    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.load()
    grid = global_grid(wl, world)
    cells = int(np.prod(grid))
    cells_local = cells // world
    if world == 1:
        parity = None if args.no_parity else guarded(parity_periodic)
        sim = PeriodicNavierStokesFlowSimulator3D(grid, X_RANGE, NU, real_t=np.float32, step_mode=(
            "auto" if args.step_mode == "auto" else args.step_mode))
        sim.vorticity_field[...] = torch.from_numpy(taylor_green_vorticity(grid)).cuda()
    else:
        from sopht_b200.parallel import SlabPeriodicNavierStokesFlowSimulator3D

        parity = None if args.no_parity else guarded(parity_periodic_slab, world)
        sim = SlabPeriodicNavierStokesFlowSimulator3D(grid, X_RANGE, NU)
        z0, nzl = sim.part.z_start, sim.part.nz_local
        w0 = taylor_green_vorticity((nzl, grid[1], grid[2]), z_offset=z0, nx_global=grid[2])
        sim.owned(sim.vorticity_field)[...] = torch.from_numpy(w0).cuda()
    sim.compute_velocity_from_vorticity()
    dt = float(0.1 * sim.dx)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def device_step():
        sim.time_step(dt)

    def e2e_step():
        sim.time_step(sim.compute_stable_timestep(dt_prefac=0.5))  # device reduction + D2H scalar every step

    for _ in range(max(args.warmup, 3)):
        device_step()
    n0 = _lib.launch_count()
    with ClockSampler(local) as clk:
        ms = timed(device_step, args.steps)
    launches = _lib.launch_count() - n0
    value = cells * args.steps / (ms * 1e-3) / 1e9
    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    peak, peak_src = measured_peak_hbm()
    _lib.profile_enable(True)
    barrier()
    for _ in range(args.steps):
        device_step()
    barrier()
    report = _lib.profile_report()
    _lib.profile_enable(False)
    if rank != 0:
        return
    cells_all, cells = cells, cells_local  # per-GPU figures below
    kernels = {}
    total = sum(v["ms"] for v in report.values()) or 1.0
    for label, v in report.items():
        e = {"launches_per_step": v["launches"] / args.steps, "ms_per_step": v["ms"] / args.steps,
             "share": v["ms"] / total}
        bpc = PERIODIC_KERNEL_BYTES.get(label)
        if bpc:
            e["algorithmic_bytes_per_cell"] = bpc
            e["achieved_gbs"] = bpc * cells * v["launches"] / (v["ms"] * 1e-3) / 1e9
            e["frac"] = e["achieved_gbs"] / peak
        kernels[label] = e
    # the Poisson solve as one entry (cuFFT passes + the symbol kernel): 120 B / cell algorithmic
    pois_ms = sum(v["ms"] for k, v in report.items() if k.startswith("poisson"))
    pois = {"ms_per_step": pois_ms / args.steps, "algorithmic_bytes_per_cell": 120.0,
            "achieved_gbs": 120.0 * cells * args.steps / (pois_ms * 1e-3) / 1e9 if pois_ms else None}
    if pois["achieved_gbs"]:
        pois["frac"] = pois["achieved_gbs"] / peak
    whole = PERIODIC_BYTES_PER_CELL * cells * args.steps / (ms * 1e-3) / 1e9  # per GPU
    roof = {"bound": "hbm", "achieved": pois["achieved_gbs"], "peak": peak, "unit": "GB/s",
            "frac": pois.get("frac"), "traffic": None, "kernel": "periodic Poisson solve (all passes)",
            "algorithmic_bytes_per_launch": 120.0 * cells, "avg_launch_ms": pois["ms_per_step"],
            "share_of_step": pois_ms / total, "peak_source": peak_src}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        from oracle import flow as oflow

        g = cpu_sample_grid(grid, 2**22)
        ref = oflow.PeriodicNavierStokesFlowSimulator3D(g, X_RANGE, NU, real_t=np.float32)
        ref.vorticity_field[...] = taylor_green_vorticity(g)
        ref.compute_velocity_from_vorticity()
        ref.time_step(dt)
        t0 = time.perf_counter()
        for _ in range(3):
            ref.time_step(dt)
        el = (time.perf_counter() - t0) / 3
        cpu = {"value": int(np.prod(g)) / el / 1e9, "unit": "Gcell/s", "cores": 1, "kind": "port",
               "sample": f"3 steps (+1 warm-up) of the numpy restatement of the periodic step on a {g[0]}x{g[1]}x{g[2]} "
                         f"grid (bounded sample), {el * 1e3:.0f} ms/step"}
    line = {
        "metric": "3D flow step Gcell-updates/s", "value": value, "unit": "Gcell/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(wl, grid, world, 0),
        "path": {"step_mode": sim.step_mode, "poisson_path": sim._poisson.path},
        "parity": parity,
        "e2e": {"value": cells_all * args.steps / (ms_e2e * 1e-3) / 1e9, "unit": "Gcell/s",
                "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4,
                "note": "stable-dt read-back every step, then the step"},
        "gpu_launches": int(launches), "clocks": clk.summary(), "roofline": roof,
        "step_roofline": {"algorithmic_bytes_per_cell": PERIODIC_BYTES_PER_CELL, "achieved": whole, "unit": "GB/s",
                          "frac": whole / peak},
        "kernels": kernels, "poisson": pois, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------
# 2-D flow past a cylinder (BASELINE configs[0]); single GPU, latency-bound
# ---------------------------------------------------------------------------------------------------------
def cylinder_case():
    radius = 0.03  # flow_past_cylinder.py:31-66
    n = 60
    a = 2 * np.pi * (np.arange(n) + 0.5) / n
    pos = np.stack([2.5 * radius + radius * np.cos(a), 0.25 + radius * np.sin(a)])
    ds = 2 * np.pi * radius / n
    return pos, radius * 1.0 / 100.0, -5e4 * ds, -20.0 * ds


def run_2d(args, wl):
    import torch

    from sopht_b200 import _lib

    grid = wl["grid"]
    cells = int(np.prod(grid))
    pos, nu, stiff, damp = cylinder_case()
    kw = dict(grid_size=grid, x_range=1.0, kinematic_viscosity=nu, real_t=np.float32, with_forcing=True,
              with_free_stream_flow=True)
    cores = len(os.sched_getaffinity(0))
    if args.impl == "reference":
        from oracle import cstencils
        from oracle import flow as oflow
        from oracle import ib as oib

        cstencils.load()
        cstencils.set_num_threads(cores)
        ref = oflow.UnboundedNavierStokesFlowSimulator2D(workers=cores, kernels=cstencils, **kw)
        vb = oib.VirtualBoundaryForcing(stiff, damp, 2, ref.dx, pos.shape[1], np.float32)
        ref.velocity_field[0] = 1.0
        vel = np.zeros_like(pos)
        dt = float(0.05 * ref.dx)

        def cpu_step():
            vb.time_step(dt)
            vb.compute_interaction_force_on_eul_and_lag_grid(ref.eul_grid_forcing_field, ref.velocity_field, pos, vel)
            ref.time_step(dt, free_stream_velocity=[1.0, 0.0])

        for _ in range(args.warmup):
            cpu_step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_step()
        el = time.perf_counter() - t0
        val = cells * args.steps / el / 1e9
        print(json.dumps({
            "impl": "reference", "metric": "2D flow step Gcell-updates/s", "value": val, "unit": "Gcell/s",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
            "steps_per_s": args.steps / el, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(wl, grid, 1, int(pos.shape[1])),
            "cpu_baseline": {"value": val, "unit": "Gcell/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} full coupled steps; {CPU_KIND_DESC}"},
            "e2e": {"value": val, "unit": "Gcell/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}), flush=True)
        return
    from sopht_b200.numeric.immersed_boundary_ops import VirtualBoundaryForcing
    from sopht_b200.simulator import UnboundedNavierStokesFlowSimulator2D

    torch.cuda.set_device(0)
    _lib.load()
    sim = UnboundedNavierStokesFlowSimulator2D(**kw)
    sim.velocity_field[0] = 1.0
    vb = VirtualBoundaryForcing(virtual_boundary_stiffness_coeff=stiff, virtual_boundary_damping_coeff=damp,
                                grid_dim=2, dx=sim.dx, num_lag_nodes=pos.shape[1], real_t=np.float32)
    pos_h = torch.from_numpy(pos).pin_memory()
    vel_h = torch.zeros_like(pos_h).pin_memory()
    pos_d, vel_d = pos_h.cuda(), vel_h.cuda()
    force_h = torch.zeros(2, pos.shape[1], dtype=torch.float32).pin_memory()
    dt = float(0.05 * sim.dx)
    u_inf = [1.0, 0.0]

    def ib_part():
        vb.time_step(dt)
        vb.compute_interaction_force_on_eul_and_lag_grid(sim.eul_grid_forcing_field, sim.velocity_field, pos_d, vel_d)

    def device_step():
        ib_part()
        sim.time_step(dt=dt, free_stream_velocity=u_inf)

    def e2e_step():
        step_dt = sim.compute_stable_timestep(dt_prefac=0.5)
        p, v = pos_h.to("cuda", non_blocking=True), vel_h.to("cuda", non_blocking=True)
        vb.time_step(step_dt)
        vb.compute_interaction_force_on_eul_and_lag_grid(sim.eul_grid_forcing_field, sim.velocity_field, p, v)
        force_h.copy_(vb.lag_grid_forcing_field, non_blocking=True)
        sim.time_step(dt=step_dt, free_stream_velocity=u_inf)

    def timed(fn, steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    steps = max(args.steps, 200)  # a step is tens of microseconds: time at least a few milliseconds
    for _ in range(max(args.warmup, 3)):
        device_step()
    ms_eager = timed(device_step, steps)
    graphed = None
    if args.graph != "off":
        state = [v for v in vars(vb).values() if isinstance(v, torch.Tensor)]
        graphed = sim.graph_time_step(dt, before=ib_part, extra_state=state, free_stream_velocity=u_inf)
        for _ in range(3):
            graphed()
    n0 = _lib.launch_count()
    with ClockSampler(0) as clk:
        ms = timed(graphed or device_step, steps)
    launches = _lib.launch_count() - n0
    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, steps)
    peak, peak_src = measured_peak_hbm()
    bpc = 88.0  # 2-D unbounded step, fp32 (L2-resident at this size: reported, not a roofline claim)
    line = {
        "metric": "2D flow step Gcell-updates/s", "value": cells * steps / (ms * 1e-3) / 1e9, "unit": "Gcell/s",
        "n_gpus": 1, "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / steps,
        "steps_per_s": steps / (ms * 1e-3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(wl, grid, 1, int(pos.shape[1])),
        "path": {"step_mode": "unfused kernels", "cuda_graph": graphed is not None,
                 "eager_ms_per_step": ms_eager / steps, "launches_per_step": launches / steps},
        "e2e": {"value": cells * steps / (ms_e2e * 1e-3) / 1e9, "unit": "Gcell/s", "ms_per_step": ms_e2e / steps,
                "h2d_bytes_per_step": int(pos_h.numel() * 16), "d2h_bytes_per_step": int(4 + force_h.numel() * 4)},
        "gpu_launches": int(launches), "clocks": clk.summary(),
        "roofline": {"bound": "hbm", "achieved": bpc * cells * steps / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": bpc * cells * steps / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                     "kernel": "whole 2-D step (0.5 MB fields are L2-resident: launch-latency-bound)",
                     "peak_source": peak_src},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="u512", choices=sorted(WORKLOADS))
    ap.add_argument("--step-mode", default="auto", choices=["auto", "fused", "unfused"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the device-resident step from a CUDA graph (auto: grids up to 2^24 cells, one GPU)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if wl.get("two_d"):
        run_2d(args, wl)
    elif wl.get("periodic"):
        if args.impl == "reference":
            raise SystemExit("the periodic workloads have no reference arm (the reference has no periodic case)")
        run_periodic(args, wl)
    elif args.impl == "reference":
        run_reference_arm(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
