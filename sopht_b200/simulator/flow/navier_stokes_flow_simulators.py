"""Unbounded Navier-Stokes flow simulators, 2D and 3D.

Same constructor kwargs, attribute names and step order as
sopht/simulator/flow/navier_stokes_flow_simulators.py:24-522; fields are torch CUDA tensors and every
sub-step is a CUDA kernel behind the C ABI. Two step implementations are kept:

* ``step_mode="fused"`` (default when the grid qualifies): the fused phase kernels of
  csrc/fused_step3d.cu + the pruned FFT pipeline of csrc/poisson_pow2.cu (few launches, minimum HBM traffic);
* ``step_mode="unfused"``: the reference's own composition, one public factory kernel per sub-step
  (what the reference's integration tests rebuild by hand, tests/.../test_navier_stokes_flow_simulators.py:163-308).
"""

from __future__ import annotations

import ctypes
import logging
import os
from typing import Any, Literal

import numpy as np
import torch

import sopht_b200.numeric.eulerian_grid_ops as spne
from sopht_b200 import _lib
from sopht_b200.utils.field import VectorField

from .flow_simulators import FlowSimulator

logger = logging.getLogger(__name__)


def compute_advection_diffusion_stable_timestep(
    velocity_field: torch.Tensor,
    velocity_magnitude_field: torch.Tensor,
    grid_dim: int,
    dx: float,
    cfl: float,
    kinematic_viscosity: float,
    real_t: type = np.float32,
) -> float:
    """Stable dt from advection and diffusion limits (passive_transport_flow_simulators.py:139-155).

    One fused kernel writes ``velocity_magnitude_field = sum_c |u_c|`` (the reference's side effect)
    and reduces its maximum on the device; the only host sync is the read of that scalar.
    """
    dt_code = _lib.dtype_code(real_t)
    vmax = torch.zeros(1, dtype=velocity_field.dtype, device=velocity_field.device)
    _lib.call("sopht_abs_sum_max", dt_code, velocity_magnitude_field, velocity_field, vmax.data_ptr())
    return stable_timestep_from_max(real_t(vmax.item()), grid_dim, dx, cfl, kinematic_viscosity, real_t)


def stable_timestep_from_max(vel_max, grid_dim, dx, cfl, kinematic_viscosity, real_t):
    tol = 10 * np.finfo(real_t).eps
    with np.errstate(divide="ignore"):
        diffusion_limit = 0.9 * dx**2 / (2 * grid_dim) / kinematic_viscosity + tol if kinematic_viscosity \
            else np.inf
    return min(cfl * dx / (vel_max + tol), diffusion_limit)


class UnboundedNavierStokesFlowSimulator3D(FlowSimulator):
    """3D unbounded Navier-Stokes flow simulator (vorticity-velocity, rotational form)."""

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
        filter_vorticity: bool = False,
        flow_density: float = 1.0,
        poisson_solver_type: Literal[
            "greens_function_convolution", "fast_diagonalisation"
        ] = "greens_function_convolution",
        **kwargs: Any,
    ) -> None:
        self.kinematic_viscosity = kinematic_viscosity
        self.cfl = cfl
        self.with_forcing = with_forcing
        self.with_free_stream_flow = with_free_stream_flow
        self.penalty_zone_width = kwargs.get("penalty_zone_width", 2)
        self.filter_vorticity = filter_vorticity
        self.filter_setting_dict = None
        if self.filter_vorticity:
            self.filter_setting_dict = kwargs.get(
                "filter_setting_dict", {"order": 2, "type": "multiplicative"}
            )
            logger.info(
                "Vorticity filtering is turned on: order = %s, type = %s",
                self.filter_setting_dict["order"], self.filter_setting_dict["type"],
            )
        self.flow_density = flow_density
        self.poisson_solver_type = poisson_solver_type
        if poisson_solver_type not in ["greens_function_convolution", "fast_diagonalisation"]:
            msg = "Invalid Poisson solver type given"
            raise ValueError(msg)
        self.step_mode = kwargs.get("step_mode", "auto")
        if self.step_mode not in ("auto", "fused", "unfused"):
            msg = "step_mode must be 'auto', 'fused' or 'unfused'"
            raise ValueError(msg)
        self._poisson_flags = kwargs.get("poisson_flags", spne.POISSON_AUTO)
        super().__init__(grid_dim=3, grid_size=grid_size, x_range=x_range, real_t=real_t,
                         num_threads=num_threads, time=time)

    def _init_fields(self) -> None:
        """navier_stokes_flow_simulators.py:314-325 (buffer_scalar_field aliases buffer_vector_field[0])."""
        shape = (self.grid_dim, *self.grid_size)
        self.vorticity_field = self._zeros(shape)
        self.velocity_field = self._zeros(shape)
        self.buffer_vector_field = self._zeros(shape)
        self.buffer_scalar_field = self.buffer_vector_field[0]
        self.stream_func_field = self._zeros(shape)
        if self.with_forcing:
            self.eul_grid_forcing_field = self._zeros(shape)

    def _compile_kernels(self) -> None:
        """Kernel factories of the step (navier_stokes_flow_simulators.py:328-440)."""
        rt, nt, gs = self.real_t, self.num_threads, self.grid_size
        self._diffusion_timestep = spne.gen_diffusion_timestep_euler_forward_pyst_kernel_3d(
            real_t=rt, fixed_grid_size=gs, num_threads=nt, field_type="vector")
        nz, ny, nx = gs
        if self.poisson_solver_type == "greens_function_convolution":
            self._unbounded_poisson_solver = spne.UnboundedPoissonSolverPYFFTW3D(
                grid_size_z=nz, grid_size_y=ny, grid_size_x=nx, x_range=self.x_range, real_t=rt,
                num_threads=nt, flags=self._poisson_flags)
        else:  # "fast_diagonalisation": Neumann walls (navier_stokes_flow_simulators.py:349-358)
            self._unbounded_poisson_solver = spne.FastDiagPoissonSolver3D(
                grid_size_z=nz, grid_size_y=ny, grid_size_x=nx, dx=self.dx, real_t=rt,
                bc_type="homogenous_neumann_along_xyz")
        self._curl = spne.gen_curl_pyst_kernel_3d(real_t=rt, num_threads=nt, fixed_grid_size=gs)
        self._penalise_field_towards_boundary = spne.gen_penalise_field_boundary_pyst_kernel_3d(
            width=self.penalty_zone_width, dx=self.dx,
            x_grid_field=self.position_field[VectorField.x_axis_idx()],
            y_grid_field=self.position_field[VectorField.y_axis_idx()],
            z_grid_field=self.position_field[VectorField.z_axis_idx()],
            real_t=rt, num_threads=nt, fixed_grid_size=gs, field_type="vector")
        self._elementwise_cross_product = spne.gen_elementwise_cross_product_pyst_kernel_3d(
            real_t=rt, num_threads=nt, fixed_grid_size=gs)
        self._update_vorticity_from_velocity_forcing = (
            spne.gen_update_vorticity_from_velocity_forcing_pyst_kernel_3d(
                real_t=rt, fixed_grid_size=gs, num_threads=nt))
        self._compute_divergence = spne.gen_divergence_pyst_kernel_3d(
            real_t=rt, fixed_grid_size=gs, num_threads=nt)

        def filter_vector_field(vector_field: torch.Tensor) -> None: ...

        self._filter_vector_field = filter_vector_field
        if self.filter_vorticity and self.filter_setting_dict is not None:
            self._filter_vector_field = spne.gen_laplacian_filter_kernel_3d(
                filter_order=self.filter_setting_dict["order"],
                filter_flux_buffer=self.buffer_vector_field[0],
                field_buffer=self.buffer_vector_field[1],
                real_t=rt, num_threads=nt, fixed_grid_size=gs, field_type="vector",
                filter_type=self.filter_setting_dict["type"])
        if self.with_forcing:
            self._set_field = spne.gen_set_fixed_val_pyst_kernel_3d(
                real_t=rt, num_threads=nt, field_type="vector")
        if self.with_free_stream_flow:
            add_fixed_val = spne.gen_add_fixed_val_pyst_kernel_3d(
                real_t=rt, fixed_grid_size=gs, num_threads=nt, field_type="vector")

            def update_velocity_with_free_stream(free_stream_velocity) -> None:
                add_fixed_val(sum_field=self.velocity_field, vector_field=self.velocity_field,
                              fixed_vals=free_stream_velocity)
        else:

            def update_velocity_with_free_stream(free_stream_velocity) -> None: ...

        self._update_velocity_with_free_stream = update_velocity_with_free_stream

    def _finalise_flow_time_step(self) -> None:
        self._dt_code = _lib.dtype_code(self.real_t)
        self._vel_absmax = self._zeros(1)  # max_cells sum_c |u_c|, refreshed by the fused velocity pass
        self._vel_absmax_version = None
        if self.step_mode == "auto":
            self.step_mode = "fused"
        if self.step_mode == "fused":
            self._flow_time_step = self._navier_stokes_fused_time_step
            return
        self._flow_time_step = self._navier_stokes_time_step
        if self.with_forcing:
            self._flow_time_step = self._navier_stokes_with_forcing_time_step

    def _navier_stokes_fused_time_step(self, dt: float, free_stream_velocity=(0.0, 0.0, 0.0)) -> None:
        """The step of navier_stokes_flow_simulators.py:449-498 with its stencil sub-steps fused:

        [forcing] w += dt/(2 dx rho) curl(f)                      (in place, f read at neighbours)
        advect    buf = w + dt/(2 dx) curl(u x w)                 (cross product never stored)
        diffuse   w = buf + nu dt/dx^2 Lap(buf)  [and f <- 0]
        [filter], penalise, Poisson (pruned FFT pipeline)
        velocity  u = curl(psi)/(2 dx) + U_inf, and max sum|u| for the next stable-dt estimate
        """
        rt, dc = self.real_t, self._dt_code
        if self.with_forcing:
            self._update_vorticity_from_velocity_forcing(
                vorticity_field=self.vorticity_field, velocity_forcing_field=self.eul_grid_forcing_field,
                prefactor=rt(dt / (2 * self.dx * self.flow_density)))
        lib = _lib.load()
        fd = _lib.field_desc
        st = _lib.current_stream()
        fw, fu, fb = fd(self.vorticity_field, dc), fd(self.velocity_field, dc), fd(self.buffer_vector_field, dc)
        _lib.check(lib.sopht_ns3d_advect_rotational(
            dc, ctypes.byref(fb), ctypes.byref(fw), ctypes.byref(fu), float(rt(dt / (2 * self.dx))), st))
        ff = fd(self.eul_grid_forcing_field, dc) if self.with_forcing else None
        nu_dt_by_dx2 = float(rt(self.kinematic_viscosity * dt / self.dx / self.dx))
        ramps = self._penalty_ramps()
        if ramps is not None:  # width-2 penalisation folded into the diffusion pass (no filter in between)
            _lib.check(lib.sopht_ns3d_diffuse_penalise(
                dc, ctypes.byref(fw), ctypes.byref(fb), nu_dt_by_dx2, ctypes.byref(ff) if ff is not None else None,
                ctypes.c_void_p(ramps[0].data_ptr()), ctypes.c_void_p(ramps[1].data_ptr()),
                ctypes.c_void_p(ramps[2].data_ptr()), st))
        else:
            _lib.check(lib.sopht_ns3d_diffuse(
                dc, ctypes.byref(fw), ctypes.byref(fb), nu_dt_by_dx2, ctypes.byref(ff) if ff is not None else None, st))
            self._filter_vector_field(vector_field=self.vorticity_field)
            self._penalise_field_towards_boundary(vector_field=self.vorticity_field)
        self._unbounded_poisson_solver.vector_field_solve(
            solution_vector_field=self.stream_func_field, rhs_vector_field=self.vorticity_field)
        fpsi = fd(self.stream_func_field, dc)
        fsv = _lib.double_array(free_stream_velocity, 3) if self.with_free_stream_flow else None
        _lib.check(lib.sopht_ns3d_velocity_from_stream_function(
            dc, ctypes.byref(fu), ctypes.byref(fpsi), float(rt(0.5 / self.dx)), fsv,
            ctypes.c_void_p(self._vel_absmax.data_ptr()), st))
        self._vel_absmax_version = (self.velocity_field.data_ptr(), self.velocity_field._version)

    def _penalty_ramps(self):
        """Per-axis factors (x, y, z device arrays) that ARE the sine penalisation for the default zone width 2
        (penalise_field_boundary_3d.py:182-208: the boundary cell takes its neighbour's value times sin(0) = 0, the
        neighbour is scaled by sin(pi/4)); None when the fused form does not apply (other widths, vorticity filter,
        rows that are not 16-byte multiples)."""
        if getattr(self, "_ramps_cache", False) is not False:
            return self._ramps_cache
        self._ramps_cache = None
        nz, ny, nx = self.grid_size
        elem = np.dtype(self.real_t).itemsize
        if (self.penalty_zone_width == 2 and not self.filter_vorticity and (nx * elem) % 16 == 0
                and min(nz, ny, nx) >= 4 and os.environ.get("SOPHT_FUSE_PENALISE", "1") != "0"):
            from sopht_b200.numeric.eulerian_grid_ops.stencil_ops_3d import _sine_ramps

            out = []
            for axis, n in ((2, nx), (1, ny), (0, nz)):  # x, y, z
                idx = [0, 0, 0]
                idx[axis] = slice(None)
                coords = self.position_field[2 - axis][tuple(idx)].cpu().numpy()
                r = _sine_ramps(coords, 2, self.dx, self.real_t)
                f = np.ones(n, dtype=self.real_t)
                f[:2] = r[:2]
                f[-2:] = r[2:]
                out.append(torch.from_numpy(f).to(self.vorticity_field.device))
            self._ramps_cache = out
        return self._ramps_cache

    def _navier_stokes_time_step(self, dt: float, free_stream_velocity=(0.0, 0.0, 0.0)) -> None:
        """navier_stokes_flow_simulators.py:449-485."""
        velocity_cross_vorticity = self.buffer_vector_field
        self._elementwise_cross_product(
            result_field=velocity_cross_vorticity, field_1=self.velocity_field,
            field_2=self.vorticity_field)
        self._update_vorticity_from_velocity_forcing(
            vorticity_field=self.vorticity_field, velocity_forcing_field=velocity_cross_vorticity,
            prefactor=self.real_t(dt / (2 * self.dx)))
        self._diffusion_timestep(
            vector_field=self.vorticity_field, diffusion_flux=self.buffer_scalar_field,
            nu_dt_by_dx2=self.real_t(self.kinematic_viscosity * dt / self.dx / self.dx))
        self._filter_vector_field(vector_field=self.vorticity_field)
        self._penalise_field_towards_boundary(vector_field=self.vorticity_field)
        self._unbounded_poisson_solver.vector_field_solve(
            solution_vector_field=self.stream_func_field, rhs_vector_field=self.vorticity_field)
        self._curl(curl=self.velocity_field, field=self.stream_func_field,
                   prefactor=self.real_t(0.5 / self.dx))
        self._update_velocity_with_free_stream(free_stream_velocity=free_stream_velocity)

    def _navier_stokes_with_forcing_time_step(self, dt: float, free_stream_velocity=(0.0, 0.0, 0.0)) -> None:
        """navier_stokes_flow_simulators.py:487-498."""
        self._update_vorticity_from_velocity_forcing(
            vorticity_field=self.vorticity_field, velocity_forcing_field=self.eul_grid_forcing_field,
            prefactor=self.real_t(dt / (2 * self.dx * self.flow_density)))
        self._navier_stokes_time_step(dt=dt, free_stream_velocity=free_stream_velocity)
        self._set_field(vector_field=self.eul_grid_forcing_field, fixed_vals=[0.0] * self.grid_dim)

    def compute_stable_timestep(self, dt_prefac: float = 1.0) -> float:
        """Stable dt (navier_stokes_flow_simulators.py:501-512). After a fused step the velocity maximum is
        already on the device (reduced inside the velocity pass), so only one scalar is read back; the
        reference's side effect of leaving sum|u| in buffer_scalar_field is kept on the uncached path only."""
        if self._vel_absmax_version == (self.velocity_field.data_ptr(), self.velocity_field._version):
            dt = stable_timestep_from_max(
                self.real_t(self._vel_absmax.item()), self.grid_dim, self.dx, self.cfl,
                self.kinematic_viscosity, self.real_t)
            return dt * dt_prefac
        dt = compute_advection_diffusion_stable_timestep(
            velocity_field=self.velocity_field, velocity_magnitude_field=self.buffer_scalar_field,
            grid_dim=self.grid_dim, dx=self.dx, cfl=self.cfl,
            kinematic_viscosity=self.kinematic_viscosity, real_t=self.real_t)
        return dt * dt_prefac

    def invalidate_stable_dt_cache(self) -> None:
        """Forget the velocity maximum the last fused step left on the device. torch's version counter sees in-place
        torch writes to ``velocity_field`` (and a rebound tensor has another data pointer), but not writes made by
        library kernels through raw pointers (``gen_add_fixed_val...`` etc. called on ``velocity_field`` by user code):
        call this after such a write, or ``compute_stable_timestep`` keeps answering for the field of the last step."""
        self._vel_absmax_version = None

    def get_vorticity_divergence_l2_norm(self) -> float:
        """L2 norm of div(vorticity) (navier_stokes_flow_simulators.py:514-522)."""
        divergence_field = self.buffer_scalar_field
        self._compute_divergence(divergence=divergence_field, field=self.vorticity_field,
                                 inv_dx=(1.0 / self.dx))
        return float(torch.linalg.vector_norm(divergence_field).item()) * self.dx ** (self.grid_dim / 2.0)


class UnboundedNavierStokesFlowSimulator2D(FlowSimulator):
    """2D unbounded Navier-Stokes flow simulator (navier_stokes_flow_simulators.py:24-209)."""

    def __init__(
        self,
        grid_size: tuple[int, int],
        x_range: float,
        kinematic_viscosity: float,
        cfl: float = 0.1,
        real_t: type = np.float32,
        num_threads: int = 1,
        time: float = 0.0,
        with_forcing: bool = False,
        with_free_stream_flow: bool = False,
        flow_density: float = 1.0,
        **kwargs: Any,
    ) -> None:
        self.kinematic_viscosity = kinematic_viscosity
        self.cfl = cfl
        self.with_forcing = with_forcing
        self.with_free_stream_flow = with_free_stream_flow
        self.flow_density = flow_density
        self.penalty_zone_width = kwargs.get("penalty_zone_width", 2)
        self._poisson_flags = kwargs.get("poisson_flags", spne.POISSON_AUTO)
        super().__init__(grid_dim=2, grid_size=grid_size, x_range=x_range, real_t=real_t,
                         num_threads=num_threads, time=time)

    def _init_fields(self) -> None:
        self.vorticity_field = self._zeros(self.grid_size)
        self.velocity_field = self._zeros((self.grid_dim, *self.grid_size))
        self.buffer_scalar_field = self._zeros(self.grid_size)
        self.stream_func_field = self._zeros(self.grid_size)
        if self.with_forcing:
            self.eul_grid_forcing_field = self._zeros((self.grid_dim, *self.grid_size))

    def _compile_kernels(self) -> None:
        rt, nt, gs = self.real_t, self.num_threads, self.grid_size
        self._diffusion_timestep = spne.gen_diffusion_timestep_euler_forward_pyst_kernel_2d(
            real_t=rt, fixed_grid_size=gs, num_threads=nt)
        self._advection_timestep = (
            spne.gen_advection_timestep_euler_forward_conservative_eno3_pyst_kernel_2d(
                real_t=rt, fixed_grid_size=gs, num_threads=nt))
        ny, nx = gs
        self._unbounded_poisson_solver = spne.UnboundedPoissonSolverPYFFTW2D(
            grid_size_y=ny, grid_size_x=nx, x_range=self.x_range, real_t=rt, num_threads=nt,
            flags=self._poisson_flags)
        self._curl = spne.gen_outplane_field_curl_pyst_kernel_2d(
            real_t=rt, num_threads=nt, fixed_grid_size=gs)
        self._penalise_field_towards_boundary = spne.gen_penalise_field_boundary_pyst_kernel_2d(
            width=self.penalty_zone_width, dx=self.dx,
            x_grid_field=self.position_field[VectorField.x_axis_idx()],
            y_grid_field=self.position_field[VectorField.y_axis_idx()],
            real_t=rt, num_threads=nt, fixed_grid_size=gs)
        if self.with_forcing:
            self._update_vorticity_from_velocity_forcing = (
                spne.gen_update_vorticity_from_velocity_forcing_pyst_kernel_2d(
                    real_t=rt, fixed_grid_size=gs, num_threads=nt))
            self._set_field = spne.gen_set_fixed_val_pyst_kernel_2d(
                real_t=rt, num_threads=nt, field_type="vector")
        if self.with_free_stream_flow:
            add_fixed_val = spne.gen_add_fixed_val_pyst_kernel_2d(
                real_t=rt, fixed_grid_size=gs, num_threads=nt, field_type="vector")

            def update_velocity_with_free_stream(free_stream_velocity) -> None:
                add_fixed_val(sum_field=self.velocity_field, vector_field=self.velocity_field,
                              fixed_vals=free_stream_velocity)
        else:

            def update_velocity_with_free_stream(free_stream_velocity) -> None: ...

        self._update_velocity_with_free_stream = update_velocity_with_free_stream

    def _finalise_flow_time_step(self) -> None:
        self._flow_time_step = self._navier_stokes_time_step
        if self.with_forcing:
            self._flow_time_step = self._navier_stokes_with_forcing_time_step

    def _navier_stokes_time_step(self, dt: float, free_stream_velocity=(0.0, 0.0)) -> None:
        """navier_stokes_flow_simulators.py:171-195."""
        self._advection_timestep(
            field=self.vorticity_field, advection_flux=self.buffer_scalar_field,
            velocity=self.velocity_field, dt_by_dx=self.real_t(dt / self.dx))
        self._diffusion_timestep(
            field=self.vorticity_field, diffusion_flux=self.buffer_scalar_field,
            nu_dt_by_dx2=self.real_t(self.kinematic_viscosity * dt / self.dx / self.dx))
        self._penalise_field_towards_boundary(field=self.vorticity_field)
        self._unbounded_poisson_solver.solve(
            solution_field=self.stream_func_field, rhs_field=self.vorticity_field)
        self._curl(curl=self.velocity_field, field=self.stream_func_field,
                   prefactor=self.real_t(0.5 / self.dx))
        self._update_velocity_with_free_stream(free_stream_velocity=free_stream_velocity)

    def _navier_stokes_with_forcing_time_step(self, dt: float, free_stream_velocity=(0.0, 0.0)) -> None:
        """navier_stokes_flow_simulators.py:197-209."""
        self._update_vorticity_from_velocity_forcing(
            vorticity_field=self.vorticity_field, velocity_forcing_field=self.eul_grid_forcing_field,
            prefactor=self.real_t(dt / (2 * self.dx * self.flow_density)))
        self._navier_stokes_time_step(dt=dt, free_stream_velocity=free_stream_velocity)
        self._set_field(vector_field=self.eul_grid_forcing_field, fixed_vals=[0.0] * self.grid_dim)

    def compute_stable_timestep(self, dt_prefac: float = 1.0) -> float:
        dt = compute_advection_diffusion_stable_timestep(
            velocity_field=self.velocity_field, velocity_magnitude_field=self.buffer_scalar_field,
            grid_dim=self.grid_dim, dx=self.dx, cfl=self.cfl,
            kinematic_viscosity=self.kinematic_viscosity, real_t=self.real_t)
        return dt * dt_prefac
