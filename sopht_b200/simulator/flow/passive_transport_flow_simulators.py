"""Passive transport (advection-diffusion of a scalar or vector field by a given velocity field) and the
backward-compatibility factory functions (sopht/simulator/flow/passive_transport_flow_simulators.py:14-136,
flow_simulators_2d.py:8-55, flow_simulators_3d.py:8-65). Fields are torch CUDA tensors; the ENO3 advection and the
diffusion step are the library's kernels behind the reference's factory names."""

from __future__ import annotations

from typing import Any

import numpy as np

import sopht_b200.numeric.eulerian_grid_ops as spne

from .flow_simulators import FlowSimulator
from .navier_stokes_flow_simulators import (
    UnboundedNavierStokesFlowSimulator2D,
    UnboundedNavierStokesFlowSimulator3D,
    compute_advection_diffusion_stable_timestep,
)


class PassiveTransportFlowSimulator(FlowSimulator):
    """passive_transport_flow_simulators.py:14-136."""

    def __init__(
        self,
        kinematic_viscosity: float,
        grid_dim: int,
        grid_size,
        x_range: float,
        cfl: float = 0.1,
        real_t: type = np.float32,
        num_threads: int = 1,
        time: float = 0.0,
        field_type: str = "scalar",
    ) -> None:
        if field_type not in ["scalar", "vector"]:
            msg = "Invalid field type. Supported values include 'scalar' and 'vector'"
            raise ValueError(msg)
        if grid_dim == 2 and field_type == "vector":
            msg = "Passive transport of vector 2D fields not supported yet."
            raise ValueError(msg)
        self.kinematic_viscosity = kinematic_viscosity
        self.cfl = cfl
        self.field_type = field_type
        super().__init__(grid_dim, grid_size, x_range, real_t, num_threads, time)

    def _init_fields(self) -> None:
        if self.field_type == "scalar":
            self.primary_field = self._zeros(self.grid_size)
        else:
            self.primary_field = self._zeros((self.grid_dim, *self.grid_size))
        self.velocity_field = self._zeros((self.grid_dim, *self.grid_size))
        self.buffer_scalar_field = self._zeros(self.grid_size)  # shared by the advection and diffusion fluxes

    def _compile_kernels(self) -> None:
        kw = dict(real_t=self.real_t, fixed_grid_size=self.grid_size, num_threads=self.num_threads)
        if self.grid_dim == 2:
            self._diffusion_timestep = spne.gen_diffusion_timestep_euler_forward_pyst_kernel_2d(**kw)
            self._advection_timestep = (
                spne.gen_advection_timestep_euler_forward_conservative_eno3_pyst_kernel_2d(**kw))
        else:
            self._diffusion_timestep = spne.gen_diffusion_timestep_euler_forward_pyst_kernel_3d(
                field_type=self.field_type, **kw)
            self._advection_timestep = (
                spne.gen_advection_timestep_euler_forward_conservative_eno3_pyst_kernel_3d(
                    field_type=self.field_type, **kw))

    def _advection_and_diffusion_time_step(self, dt: float, **kwargs: Any) -> None:
        del kwargs  # unused, as in the reference
        self._advection_timestep(
            self.primary_field, advection_flux=self.buffer_scalar_field, velocity=self.velocity_field,
            dt_by_dx=self.real_t(dt / self.dx))
        self._diffusion_timestep(
            self.primary_field, diffusion_flux=self.buffer_scalar_field,
            nu_dt_by_dx2=self.real_t(self.kinematic_viscosity * dt / self.dx / self.dx))

    def _finalise_flow_time_step(self) -> None:
        self._flow_time_step = self._advection_and_diffusion_time_step

    def compute_stable_timestep(self, dt_prefac: float = 1.0) -> float:
        dt = compute_advection_diffusion_stable_timestep(
            velocity_field=self.velocity_field, velocity_magnitude_field=self.buffer_scalar_field,
            grid_dim=self.grid_dim, dx=self.dx, cfl=self.cfl, kinematic_viscosity=self.kinematic_viscosity,
            real_t=self.real_t)
        return dt * dt_prefac


def _with_forcing(flow_type: str) -> bool:
    if flow_type == "navier_stokes":
        return False
    if flow_type == "navier_stokes_with_forcing":
        return True
    msg = "Invalid flow type given"
    raise ValueError(msg)


def create_unbounded_flow_simulator_2d(grid_size, x_range: float, kinematic_viscosity: float, cfl: float = 0.1,
                                       flow_type: str = "navier_stokes", real_t: type = np.float32,
                                       num_threads: int = 1, time: float = 0.0,
                                       **kwargs: Any) -> UnboundedNavierStokesFlowSimulator2D:
    """flow_simulators_2d.py:8-55 (kept for backward compatibility in the reference)."""
    return UnboundedNavierStokesFlowSimulator2D(
        grid_size=grid_size, x_range=x_range, kinematic_viscosity=kinematic_viscosity, cfl=cfl, real_t=real_t,
        num_threads=num_threads, time=time, with_forcing=_with_forcing(flow_type), **kwargs)


def create_unbounded_flow_simulator_3d(grid_size, x_range: float, kinematic_viscosity: float, cfl: float = 0.1,
                                       flow_type: str = "navier_stokes", real_t: type = np.float32,
                                       num_threads: int = 1, filter_vorticity: bool = False,
                                       poisson_solver_type: str = "greens_function_convolution", time: float = 0.0,
                                       **kwargs: Any) -> UnboundedNavierStokesFlowSimulator3D:
    """flow_simulators_3d.py:8-65."""
    return UnboundedNavierStokesFlowSimulator3D(
        grid_size=grid_size, x_range=x_range, kinematic_viscosity=kinematic_viscosity, cfl=cfl, real_t=real_t,
        num_threads=num_threads, time=time, with_forcing=_with_forcing(flow_type),
        filter_vorticity=filter_vorticity, poisson_solver_type=poisson_solver_type, **kwargs)
