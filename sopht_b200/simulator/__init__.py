"""Simulator-level call sites of the hot path (step sequencers), same names as sopht.simulator."""

from .flow import (
    FlowSimulator,
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
