"""CPU ORACLE (test infrastructure): flow simulator time steps, restated from
sopht/simulator/flow/navier_stokes_flow_simulators.py and passive_transport_flow_simulators.py on top
of the kernel oracles in stencils.py / poisson.py (same sub-kernel order, same buffer reuse).
"""

from __future__ import annotations

import numpy as np

from . import poisson as opoisson
from . import stencils as ost


def compute_advection_diffusion_stable_timestep(
    velocity_field, velocity_magnitude_field, grid_dim, dx, cfl, kinematic_viscosity, real_t=np.float32,
    kernels=None,
):
    """passive_transport_flow_simulators.py:139-155 (writes velocity_magnitude_field)."""
    tol = 10 * np.finfo(real_t).eps
    if kernels is not None and hasattr(kernels, "abs_sum_max"):
        vmax = kernels.abs_sum_max(velocity_magnitude_field, velocity_field)
    else:
        velocity_magnitude_field[...] = np.sum(np.fabs(velocity_field), axis=0)
        vmax = np.amax(velocity_magnitude_field)
    return min(
        cfl * dx / (vmax + tol),
        0.9 * dx**2 / (2 * grid_dim) / kinematic_viscosity + tol,
    )


def _grid_coords(grid_size, x_range, real_t):
    """flow_simulators.py:50-81: cell centres dx/2 ... range - dx/2 per array axis (z, y, x order)."""
    nx = grid_size[-1]
    dx = real_t(x_range / nx)
    shift = dx / 2.0
    coords = []
    for n in grid_size:
        rng = x_range * n / nx
        coords.append(np.linspace(shift, rng - shift, n).astype(real_t))
    return dx, coords


class UnboundedNavierStokesFlowSimulator3D:
    """navier_stokes_flow_simulators.py:212-522 (greens_function_convolution Poisson solver only)."""

    def __init__(self, grid_size, x_range, kinematic_viscosity, cfl=0.1, real_t=np.float32, time=0.0,
                 with_forcing=False, with_free_stream_flow=False, filter_vorticity=False,
                 flow_density=1.0, workers=1, kernels=None, **kwargs):
        self._k = kernels if kernels is not None else ost  # kernel vocabulary: numpy (default) or oracle.cstencils
        self.grid_dim = 3
        self.grid_size = tuple(grid_size)
        self.x_range = x_range
        self.real_t = real_t
        self.kinematic_viscosity = kinematic_viscosity
        self.cfl = cfl
        self.time = time
        self.with_forcing = with_forcing
        self.with_free_stream_flow = with_free_stream_flow
        self.filter_vorticity = filter_vorticity
        self.flow_density = flow_density
        self.penalty_zone_width = kwargs.get("penalty_zone_width", 2)
        self.filter_setting_dict = kwargs.get("filter_setting_dict", {"order": 2, "type": "multiplicative"})
        self.dx, self.coords = _grid_coords(self.grid_size, x_range, real_t)
        z, y, x = self.coords
        self.position_field = np.flipud(np.array(np.meshgrid(z, y, x, indexing="ij")))
        shape = (3, *self.grid_size)
        self.vorticity_field = np.zeros(shape, dtype=real_t)
        self.velocity_field = np.zeros_like(self.vorticity_field)
        self.buffer_vector_field = np.zeros_like(self.vorticity_field)
        self.buffer_scalar_field = self.buffer_vector_field[0].view()
        self.stream_func_field = np.zeros_like(self.vorticity_field)
        if with_forcing:
            self.eul_grid_forcing_field = np.zeros_like(self.velocity_field)
        nz, ny, nx = self.grid_size
        self._poisson = opoisson.UnboundedPoissonSolver3D(
            nz, ny, nx, x_range=x_range, real_t=real_t, workers=workers, kernels=kernels)

    def _navier_stokes_time_step(self, dt, free_stream_velocity=(0.0, 0.0, 0.0)):
        """navier_stokes_flow_simulators.py:449-485."""
        t = self.real_t
        self._k.elementwise_cross_product(self.buffer_vector_field, self.velocity_field, self.vorticity_field)
        self._k.update_vorticity_from_velocity_forcing_3d(
            self.vorticity_field, self.buffer_vector_field, t(dt / (2 * self.dx)))
        self._k.diffusion_timestep_euler_forward_vector(
            self.vorticity_field, self.buffer_scalar_field, t(self.kinematic_viscosity * dt / self.dx / self.dx))
        if self.filter_vorticity:
            self._k.laplacian_filter_3d_vector(
                self.vorticity_field, self.buffer_vector_field[0], self.buffer_vector_field[1],
                self.filter_setting_dict["order"], self.filter_setting_dict["type"])
        self._k.penalise_field_boundary_vector(self.vorticity_field, self.penalty_zone_width, self.dx, self.coords)
        self._poisson.vector_field_solve(self.stream_func_field, self.vorticity_field)
        self._k.curl_3d(self.velocity_field, self.stream_func_field, t(0.5 / self.dx))
        if self.with_free_stream_flow:
            self._k.add_fixed_val_vector(self.velocity_field, self.velocity_field, free_stream_velocity)

    def _navier_stokes_with_forcing_time_step(self, dt, free_stream_velocity=(0.0, 0.0, 0.0)):
        """navier_stokes_flow_simulators.py:487-498."""
        self._k.update_vorticity_from_velocity_forcing_3d(
            self.vorticity_field, self.eul_grid_forcing_field,
            self.real_t(dt / (2 * self.dx * self.flow_density)))
        self._navier_stokes_time_step(dt, free_stream_velocity)
        self._k.set_fixed_val_vector(self.eul_grid_forcing_field, [0.0] * 3)

    def time_step(self, dt, free_stream_velocity=(0.0, 0.0, 0.0)):
        if self.with_forcing:
            self._navier_stokes_with_forcing_time_step(dt, free_stream_velocity)
        else:
            self._navier_stokes_time_step(dt, free_stream_velocity)
        self.time += dt

    def compute_stable_timestep(self, dt_prefac=1.0):
        return dt_prefac * compute_advection_diffusion_stable_timestep(
            self.velocity_field, self.buffer_scalar_field, 3, self.dx, self.cfl,
            self.kinematic_viscosity, self.real_t, kernels=self._k)


class UnboundedNavierStokesFlowSimulator2D:
    """navier_stokes_flow_simulators.py:24-209."""

    def __init__(self, grid_size, x_range, kinematic_viscosity, cfl=0.1, real_t=np.float32, time=0.0,
                 with_forcing=False, with_free_stream_flow=False, flow_density=1.0, workers=1, kernels=None, **kwargs):
        self._k = kernels if kernels is not None else ost
        self.grid_dim = 2
        self.grid_size = tuple(grid_size)
        self.x_range = x_range
        self.real_t = real_t
        self.kinematic_viscosity = kinematic_viscosity
        self.cfl = cfl
        self.time = time
        self.with_forcing = with_forcing
        self.with_free_stream_flow = with_free_stream_flow
        self.flow_density = flow_density
        self.penalty_zone_width = kwargs.get("penalty_zone_width", 2)
        self.dx, self.coords = _grid_coords(self.grid_size, x_range, real_t)
        y, x = self.coords
        self.position_field = np.flipud(np.array(np.meshgrid(y, x, indexing="ij")))
        self.vorticity_field = np.zeros(self.grid_size, dtype=real_t)
        self.velocity_field = np.zeros((2, *self.grid_size), dtype=real_t)
        self.buffer_scalar_field = np.zeros_like(self.vorticity_field)
        self.stream_func_field = np.zeros_like(self.vorticity_field)
        if with_forcing:
            self.eul_grid_forcing_field = np.zeros_like(self.velocity_field)
        ny, nx = self.grid_size
        self._poisson = opoisson.UnboundedPoissonSolver2D(
            ny, nx, x_range=x_range, real_t=real_t, workers=workers, kernels=kernels)

    def _navier_stokes_time_step(self, dt, free_stream_velocity=(0.0, 0.0)):
        """navier_stokes_flow_simulators.py:171-195."""
        t = self.real_t
        self._k.advection_timestep_euler_forward_conservative_eno3(
            self.vorticity_field, self.buffer_scalar_field, self.velocity_field, t(dt / self.dx))
        self._k.diffusion_timestep_euler_forward(
            self.vorticity_field, self.buffer_scalar_field, t(self.kinematic_viscosity * dt / self.dx / self.dx))
        self._k.penalise_field_boundary(self.vorticity_field, self.penalty_zone_width, self.dx, self.coords)
        self._poisson.solve(self.stream_func_field, self.vorticity_field)
        self._k.outplane_field_curl_2d(self.velocity_field, self.stream_func_field, t(0.5 / self.dx))
        if self.with_free_stream_flow:
            self._k.add_fixed_val_vector(self.velocity_field, self.velocity_field, free_stream_velocity)

    def _navier_stokes_with_forcing_time_step(self, dt, free_stream_velocity=(0.0, 0.0)):
        """navier_stokes_flow_simulators.py:197-209."""
        self._k.update_vorticity_from_velocity_forcing_2d(
            self.vorticity_field, self.eul_grid_forcing_field,
            self.real_t(dt / (2 * self.dx * self.flow_density)))
        self._navier_stokes_time_step(dt, free_stream_velocity)
        self._k.set_fixed_val_vector(self.eul_grid_forcing_field, [0.0] * 2)

    def time_step(self, dt, free_stream_velocity=(0.0, 0.0)):
        if self.with_forcing:
            self._navier_stokes_with_forcing_time_step(dt, free_stream_velocity)
        else:
            self._navier_stokes_time_step(dt, free_stream_velocity)
        self.time += dt

    def compute_stable_timestep(self, dt_prefac=1.0):
        return dt_prefac * compute_advection_diffusion_stable_timestep(
            self.velocity_field, self.buffer_scalar_field, 2, self.dx, self.cfl,
            self.kinematic_viscosity, self.real_t, kernels=self._k)


def wrap_halos(field):
    """Periodic ghost cells of a halo-1 padded (..., nz+2, ny+2, nx+2) array, axis by axis so edges and corners follow."""
    for axis in (-3, -2, -1):
        f = np.moveaxis(field, axis, 0)
        f[0] = f[-2]
        f[-1] = f[1]


class PeriodicNavierStokesFlowSimulator3D:
    """Periodic 3-D vorticity-velocity step - an EXTENSION (BASELINE config 4). **Parity unpinned**: the reference has
    no periodic case. The sub-steps and their order are the reference's unbounded step
    (navier_stokes_flow_simulators.py:449-485) without the boundary penalisation; periodicity comes from a one-cell
    halo that is refilled by wrap-around before every stencil, so the reference's own ghost-ring stencils update
    exactly the true cells. Checked against the analytic decay of the 2-D Taylor-Green vortex."""

    def __init__(self, grid_size, x_range, kinematic_viscosity, cfl=0.1, real_t=np.float32, time=0.0,
                 poisson_symbol="spectral"):
        self.grid_dim, self.grid_size, self.x_range, self.real_t = 3, tuple(grid_size), x_range, real_t
        self.kinematic_viscosity, self.cfl, self.time = kinematic_viscosity, cfl, time
        self.dx, self.coords = _grid_coords(self.grid_size, x_range, real_t)
        z, y, x = self.coords
        self.position_field = np.flipud(np.array(np.meshgrid(z, y, x, indexing="ij")))
        padded = (3, *(n + 2 for n in self.grid_size))
        self._w, self._u = np.zeros(padded, dtype=real_t), np.zeros(padded, dtype=real_t)
        self._buf, self._psi = np.zeros(padded, dtype=real_t), np.zeros(padded, dtype=real_t)
        inner = (slice(None), slice(1, -1), slice(1, -1), slice(1, -1))
        self.vorticity_field, self.velocity_field = self._w[inner], self._u[inner]
        self.stream_func_field = self._psi[inner]
        self._poisson = opoisson.PeriodicPoissonSolver(self.grid_size, float(self.dx), poisson_symbol)

    def compute_velocity_from_vorticity(self):
        for c in range(3):
            self._poisson.solve(self.stream_func_field[c], self.vorticity_field[c])
        wrap_halos(self._psi)
        ost.curl_3d(self._u, self._psi, self.real_t(0.5 / self.dx))
        wrap_halos(self._u)

    def time_step(self, dt):
        t = self.real_t
        wrap_halos(self._w)
        wrap_halos(self._u)
        ost.elementwise_cross_product(self._buf, self._u, self._w)
        ost.update_vorticity_from_velocity_forcing_3d(self._w, self._buf, t(dt / (2 * self.dx)))
        wrap_halos(self._w)
        ost.diffusion_timestep_euler_forward_vector(
            self._w, self._buf[0], t(self.kinematic_viscosity * dt / self.dx / self.dx))
        self.compute_velocity_from_vorticity()
        self.time += dt

    def compute_stable_timestep(self, dt_prefac=1.0):
        return dt_prefac * compute_advection_diffusion_stable_timestep(
            self._u, self._buf[0], 3, self.dx, self.cfl, self.kinematic_viscosity, self.real_t)
