"""Flow simulators (step sequencers of the hot path)."""

from .flow_simulators import FlowSimulator
from .navier_stokes_flow_simulators import (
    UnboundedNavierStokesFlowSimulator2D,
    UnboundedNavierStokesFlowSimulator3D,
    compute_advection_diffusion_stable_timestep,
)

__all__ = [
    "FlowSimulator",
    "UnboundedNavierStokesFlowSimulator2D",
    "UnboundedNavierStokesFlowSimulator3D",
    "compute_advection_diffusion_stable_timestep",
]
