"""3-D unbounded Navier-Stokes flow simulator on a z-slab decomposed grid (one process per GPU).

Same step as UnboundedNavierStokesFlowSimulator3D (sopht/simulator/flow/navier_stokes_flow_simulators.py:
449-498, fused form in sopht_b200/simulator/flow/navier_stokes_flow_simulators.py), per rank on its slab:

    halo(w, u) -> advect -> halo(buf) -> diffuse -> penalise (z faces only on the boundary ranks)
    -> slab Poisson (two all-to-all transposes) -> halo(psi) -> velocity (+ local max for dt) -> [all-reduce max]

Local arrays are (3, nz/P + 2, ny, nx) with one halo plane per z side; kernels see views chosen so that
their ghost-ring rule lands on the true global boundary (SlabPartition.stencil_view).
"""

from __future__ import annotations

import ctypes
import os
from typing import Any

import numpy as np
import torch
import torch.distributed as dist

from sopht_b200 import _lib
from sopht_b200.numeric.eulerian_grid_ops.stencil_ops_3d import _sine_ramps
from sopht_b200.simulator.flow.navier_stokes_flow_simulators import stable_timestep_from_max

from .peer import PeerArena
from .slab import HaloExchanger, SlabPartition
from .slab_poisson import SlabPeriodicPoissonSolver3D, SlabUnboundedPoissonSolver3D


class SlabUnboundedNavierStokesFlowSimulator3D:
    """Constructor mirrors UnboundedNavierStokesFlowSimulator3D; `grid_size` is the GLOBAL (nz, ny, nx).

    Public field attributes hold this rank's slab INCLUDING the two halo planes; use `owned(field)` for the
    (3, nz/P, ny, nx) planes this rank owns and `z_slice` for their global z range."""

    def __init__(
        self,
        grid_size: tuple[int, int, int],
        x_range: float,
        kinematic_viscosity: float,
        cfl: float = 0.1,
        real_t: type = np.float32,
        num_threads: int = 1,
        time: float = 0.0,
        with_forcing: bool = False,
        with_free_stream_flow: bool = False,
        flow_density: float = 1.0,
        group: Any = None,
        **kwargs: Any,
    ) -> None:
        if _lib.dtype_code(real_t) != _lib.SOPHT_F32:
            msg = "the slab-decomposed simulator is implemented for fp32"
            raise ValueError(msg)
        self.grid_dim = 3
        self.grid_size = tuple(grid_size)
        self.x_range, self.real_t, self.num_threads, self.time = x_range, real_t, num_threads, time
        self.kinematic_viscosity, self.cfl = kinematic_viscosity, cfl
        self.with_free_stream_flow = with_free_stream_flow
        self.with_forcing, self.flow_density = with_forcing, flow_density
        self.penalty_zone_width = kwargs.get("penalty_zone_width", 2)
        self.group = group
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.part = SlabPartition(self.grid_size, world, rank, halo=1)
        if not torch.cuda.is_available():
            msg = "sopht_b200 flow simulators need a CUDA device (no CPU fallback)"
            raise _lib.SophtLibraryError(msg)
        self.device = torch.device("cuda", torch.cuda.current_device())
        nz, ny, nx = self.grid_size
        self.dx = real_t(x_range / nx)
        self.z_slice = slice(self.part.z_start, self.part.z_start + self.part.nz_local)
        # cell-centred coordinates (flow_simulators.py:50-81), only what the ramps need
        shift = self.dx / 2.0
        coords = [np.linspace(shift, x_range * n / nx - shift, n).astype(real_t) for n in (nz, ny, nx)]
        w = self.penalty_zone_width
        if w:
            self._ramps = [_sine_ramps(coords[2], w, self.dx, real_t), _sine_ramps(coords[1], w, self.dx, real_t),
                           _sine_ramps(coords[0], w, self.dx, real_t)]
            if world > 1 and self.part.nz_local < w:
                msg = "fewer planes per rank than the penalisation width"
                raise ValueError(msg)
        # width 2: the penalisation is a per-axis factor and rides in the diffusion pass (like the single-GPU
        # simulator, navier_stokes_flow_simulators.py: _penalty_ramps); the z factors are this rank's planes of the
        # stencil view (halo planes of real neighbours: 1, global boundary planes: 0 / sin(pi / 4))
        self._fused_ramps = None
        if (w == 2 and (nx * 4) % 16 == 0 and min(nz, ny, nx) >= 4
                and os.environ.get("SOPHT_FUSE_PENALISE", "1") != "0"):
            def factors(axis_coords, n):
                r = _sine_ramps(axis_coords, 2, self.dx, real_t)
                f = np.ones(n, dtype=real_t)
                f[:2], f[-2:] = r[:2], r[2:]
                return f
            fz = factors(coords[0], nz)
            lo = self.part.z_start - (0 if self.part.is_first else 1)
            hi = self.part.z_start + self.part.nz_local + (0 if self.part.is_last else 1)
            self._fused_ramps = [torch.from_numpy(factors(coords[2], nx)).to(self.device),
                                 torch.from_numpy(factors(coords[1], ny)).to(self.device),
                                 torch.from_numpy(np.ascontiguousarray(fz[lo:hi])).to(self.device)]
        shape = (3, *self.part.local_shape)
        # Fields live in a peer-memory arena when there are neighbours: the halo exchange is then one kernel of
        # direct NVLink stores into the neighbours' halo planes (SOPHT_SLAB_PEER=0: torch tensors + NCCL send/recv)
        self._arena = None
        if world > 1 and os.environ.get("SOPHT_SLAB_PEER", "1") != "0" and (ny * nx * 4) % 16 == 0:
            nfields = 5 if with_forcing else 4
            self._arena = PeerArena(nfields * (int(np.prod(shape)) * 4 + 256), group)
            zeros = lambda: self._arena.alloc(shape, np.float32)  # noqa: E731
        else:
            zeros = lambda: torch.zeros(shape, dtype=torch.float32, device=self.device)  # noqa: E731
        self.vorticity_field, self.velocity_field = zeros(), zeros()
        self.buffer_vector_field, self.stream_func_field = zeros(), zeros()
        if with_forcing:
            self.eul_grid_forcing_field = zeros()
        self._unbounded_poisson_solver = SlabUnboundedPoissonSolver3D(
            nz, ny, nx, x_range=x_range, real_t=real_t, num_threads=num_threads, group=group,
            peer_arena=self._arena)
        self._exchangers: dict = {}
        self._vel_absmax = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._have_absmax = False
        self._vel_version = None
        self.step_mode = "fused-slab"

    # -- helpers ------------------------------------------------------------------------------------------
    def owned(self, field: torch.Tensor) -> torch.Tensor:
        return self.part.owned(field)

    def set_owned(self, field: torch.Tensor, global_values: np.ndarray | torch.Tensor) -> None:
        """Fill this rank's planes (and the halo planes that have a neighbour) from a GLOBAL array."""
        g = torch.as_tensor(global_values)
        h, n, z0 = self.part.halo, self.part.nz_local, self.part.z_start
        lo, hi = max(z0 - h, 0), min(z0 + n + h, self.grid_size[0])
        field[..., h - (z0 - lo) : h + n + (hi - z0 - n), :, :] = g[..., lo:hi, :, :].to(field.device, field.dtype)
        if field is self.velocity_field:
            self._have_absmax = False  # the maximum the last step left on the device no longer describes this field

    def _halos(self, *fields: torch.Tensor) -> None:
        if self._arena is not None:
            self._arena.halo_exchange(fields, self.part.nz_local, self.part.halo)
            return
        key = tuple(f.data_ptr() for f in fields)  # one persistent exchanger (buffers) per call site
        ex = self._exchangers.get(key)
        if ex is None:
            ex = self._exchangers[key] = HaloExchanger(self.part, self.group)
        with _lib.profile_range("comm.halo_exchange"):
            ex(fields)

    # -- the step -----------------------------------------------------------------------------------------
    def time_step(self, dt: float, free_stream_velocity=(0.0, 0.0, 0.0)) -> None:
        rt, dc, part = self.real_t, _lib.SOPHT_F32, self.part
        lib, fd, st = _lib.load(), _lib.field_desc, _lib.current_stream()
        sv = part.stencil_view
        if self.with_forcing:  # w += dt/(2 dx rho) curl(f)   (navier_stokes_flow_simulators.py:487-492)
            self._halos(self.eul_grid_forcing_field)
            _lib.call("sopht_update_vorticity_from_velocity_forcing_3d", dc, sv(self.vorticity_field),
                      sv(self.eul_grid_forcing_field), float(rt(dt / (2 * self.dx * self.flow_density))))
        self._halos(self.vorticity_field, self.velocity_field)
        fw, fu, fb = fd(sv(self.vorticity_field), dc), fd(sv(self.velocity_field), dc), fd(sv(self.buffer_vector_field), dc)
        _lib.check(lib.sopht_ns3d_advect_rotational(
            dc, ctypes.byref(fb), ctypes.byref(fw), ctypes.byref(fu), float(rt(dt / (2 * self.dx))), st))
        self._halos(self.buffer_vector_field)
        ff = fd(sv(self.eul_grid_forcing_field), dc) if self.with_forcing else None  # f <- 0 in the same pass
        nu_dt_by_dx2 = float(rt(self.kinematic_viscosity * dt / self.dx / self.dx))
        if self._fused_ramps is not None:
            rx, ry, rz = self._fused_ramps
            _lib.check(lib.sopht_ns3d_diffuse_penalise(
                dc, ctypes.byref(fw), ctypes.byref(fb), nu_dt_by_dx2, ctypes.byref(ff) if ff is not None else None,
                ctypes.c_void_p(rx.data_ptr()), ctypes.c_void_p(ry.data_ptr()), ctypes.c_void_p(rz.data_ptr()), st))
        else:
            _lib.check(lib.sopht_ns3d_diffuse(
                dc, ctypes.byref(fw), ctypes.byref(fb), nu_dt_by_dx2, ctypes.byref(ff) if ff is not None else None, st))
            if self.penalty_zone_width:
                _lib.call("sopht_penalise_field_boundary_3d_slab", dc, self.owned(self.vorticity_field),
                          self.penalty_zone_width, self._ramps[0], self._ramps[1], self._ramps[2], part.z_faces)
        self._unbounded_poisson_solver.vector_field_solve(
            solution_vector_field=self.owned(self.stream_func_field),
            rhs_vector_field=self.owned(self.vorticity_field))
        self._halos(self.stream_func_field)
        fpsi = fd(sv(self.stream_func_field), dc)
        fsv = _lib.double_array(free_stream_velocity, 3) if self.with_free_stream_flow else None
        _lib.check(lib.sopht_ns3d_velocity_from_stream_function(
            dc, ctypes.byref(fu), ctypes.byref(fpsi), float(rt(0.5 / self.dx)), fsv,
            ctypes.c_void_p(self._vel_absmax.data_ptr()), st))
        self._have_absmax = True
        self._vel_version = self.velocity_field._version
        self.time += dt

    def compute_stable_timestep(self, dt_prefac: float = 1.0) -> float:
        """min(cfl dx / max sum|u|, 0.9 dx^2 / (6 nu)) with the maximum taken over all ranks
        (passive_transport_flow_simulators.py:139-155)."""
        if not self._have_absmax or self._vel_version != self.velocity_field._version:  # before the first step / after a write: the library's sum|u| + max kernel (writes buffer[0])
            sv = self.part.stencil_view
            _lib.call("sopht_abs_sum_max", _lib.SOPHT_F32, sv(self.buffer_vector_field)[0],
                      sv(self.velocity_field), self._vel_absmax.data_ptr())
        m = self._vel_absmax.clone()
        if self._arena is not None:
            self._arena.check()  # the host synchronises here anyway: surface a timed-out peer exchange as an error
        if self.part.world_size > 1:
            dist.all_reduce(m, op=dist.ReduceOp.MAX, group=self.group)
        dt = stable_timestep_from_max(self.real_t(m.item()), 3, self.dx, self.cfl, self.kinematic_viscosity, self.real_t)
        return dt * dt_prefac


class SlabPeriodicNavierStokesFlowSimulator3D:
    """Periodic 3-D Navier-Stokes step (BASELINE config 4: Taylor-Green vortex) on a z-slab decomposed grid - the
    distributed twin of PeriodicNavierStokesFlowSimulator3D (an extension: the reference has no periodic case, nearest
    reference step navier_stokes_flow_simulators.py:449-485 without the penalisation; parity is against the single-GPU
    class, which is checked against its numpy restatement and the analytic Taylor-Green decay).

    Per rank: (3, nz/P + 2, ny, nx) arrays in a peer-memory arena; x and y wrap inside the marching kernels, the z halo
    planes come from the ring neighbours (rank 0 <-> rank P - 1 close the ring) in one peer-store kernel per exchange;
    the Poisson solve is SlabPeriodicPoissonSolver3D. With one rank the halo planes are the rank's own opposite
    planes."""

    def __init__(self, grid_size: tuple[int, int, int], x_range: float, kinematic_viscosity: float, cfl: float = 0.1,
                 real_t: type = np.float32, num_threads: int = 1, time: float = 0.0, group: Any = None,
                 poisson_symbol: str = "spectral") -> None:
        if _lib.dtype_code(real_t) != _lib.SOPHT_F32:
            msg = "the slab-decomposed simulator is implemented for fp32"
            raise ValueError(msg)
        if not torch.cuda.is_available():
            msg = "sopht_b200 flow simulators need a CUDA device (no CPU fallback)"
            raise _lib.SophtLibraryError(msg)
        self.grid_dim = 3
        self.grid_size = tuple(grid_size)
        self.x_range, self.real_t, self.num_threads, self.time = x_range, real_t, num_threads, time
        self.kinematic_viscosity, self.cfl = kinematic_viscosity, cfl
        self.group = group
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.part = SlabPartition(self.grid_size, world, rank, halo=1)
        self.device = torch.device("cuda", torch.cuda.current_device())
        nz, ny, nx = self.grid_size
        if (nx * 4) % 16:
            msg = "the periodic marching kernels need rows that are multiples of 16 bytes"
            raise ValueError(msg)
        self.dx = real_t(x_range / nx)
        self.z_slice = slice(self.part.z_start, self.part.z_start + self.part.nz_local)
        shape = (3, *self.part.local_shape)
        self._arena = None
        if world > 1:
            self._arena = PeerArena(4 * (int(np.prod(shape)) * 4 + 256), group, periodic_z=True)
            zeros = lambda: self._arena.alloc(shape, np.float32)  # noqa: E731
        else:
            zeros = lambda: torch.zeros(shape, dtype=torch.float32, device=self.device)  # noqa: E731
        self.vorticity_field, self.velocity_field = zeros(), zeros()
        self.buffer_vector_field, self.stream_func_field = zeros(), zeros()
        self._poisson = SlabPeriodicPoissonSolver3D(
            nz, ny, nx, x_range=x_range, real_t=real_t, num_threads=num_threads, group=group,
            peer_arena=self._arena, symbol=poisson_symbol)
        self._vel_absmax = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._vel_version = None
        self.step_mode = "fused-slab-periodic"

    def owned(self, field: torch.Tensor) -> torch.Tensor:
        return self.part.owned(field)

    def set_owned(self, field: torch.Tensor, global_values: np.ndarray | torch.Tensor) -> None:
        """Fill this rank's owned planes from a GLOBAL array (the halo planes are refreshed by the next exchange)."""
        g = torch.as_tensor(global_values)
        self.owned(field)[...] = g[..., self.z_slice, :, :].to(field.device, field.dtype)
        if field is self.velocity_field:
            self._vel_version = None

    def _halos(self, *fields: torch.Tensor) -> None:
        if self._arena is not None:
            self._arena.halo_exchange(fields, self.part.nz_local, self.part.halo)
            return
        for f in fields:
            _lib.call("sopht_wrap_z_halos", _lib.SOPHT_F32, f)

    def compute_velocity_from_vorticity(self) -> None:
        lib, fd, st, dc = _lib.load(), _lib.field_desc, _lib.current_stream(), _lib.SOPHT_F32
        self._poisson.vector_field_solve(solution_vector_field=self.owned(self.stream_func_field),
                                         rhs_vector_field=self.owned(self.vorticity_field))
        self._halos(self.stream_func_field)
        fu, fpsi = fd(self.velocity_field, dc), fd(self.stream_func_field, dc)
        _lib.check(lib.sopht_ns3d_velocity_from_stream_function_periodic_xy(
            dc, ctypes.byref(fu), ctypes.byref(fpsi), float(self.real_t(0.5 / self.dx)), None,
            ctypes.c_void_p(self._vel_absmax.data_ptr()), st))
        self._vel_version = self.velocity_field._version

    def time_step(self, dt: float) -> None:
        rt, dc = self.real_t, _lib.SOPHT_F32
        lib, fd, st = _lib.load(), _lib.field_desc, _lib.current_stream()
        self._halos(self.vorticity_field, self.velocity_field)
        fw, fu, fb = fd(self.vorticity_field, dc), fd(self.velocity_field, dc), fd(self.buffer_vector_field, dc)
        _lib.check(lib.sopht_ns3d_advect_rotational_periodic_xy(
            dc, ctypes.byref(fb), ctypes.byref(fw), ctypes.byref(fu), float(rt(dt / (2 * self.dx))), st))
        self._halos(self.buffer_vector_field)
        _lib.check(lib.sopht_ns3d_diffuse_periodic_xy(
            dc, ctypes.byref(fw), ctypes.byref(fb), float(rt(self.kinematic_viscosity * dt / self.dx / self.dx)),
            None, st))
        self.compute_velocity_from_vorticity()
        self.time += dt

    def compute_stable_timestep(self, dt_prefac: float = 1.0) -> float:
        if self._vel_version is None or self._vel_version != self.velocity_field._version:
            _lib.call("sopht_abs_sum_max", _lib.SOPHT_F32, self.owned(self.buffer_vector_field)[0],
                      self.owned(self.velocity_field), self._vel_absmax.data_ptr())
        m = self._vel_absmax.clone()
        if self._arena is not None:
            self._arena.check()
        if self.part.world_size > 1:
            dist.all_reduce(m, op=dist.ReduceOp.MAX, group=self.group)
        dt = stable_timestep_from_max(self.real_t(m.item()), 3, self.dx, self.cfl, self.kinematic_viscosity, self.real_t)
        return dt * dt_prefac
