"""Flow simulators (step sequencers of the hot path), same names as sopht.simulator.flow."""

from .flow_simulators import FlowSimulator
from .navier_stokes_flow_simulators import (
    UnboundedNavierStokesFlowSimulator2D,
    UnboundedNavierStokesFlowSimulator3D,
    compute_advection_diffusion_stable_timestep,
)
from .periodic_flow_simulators import PeriodicNavierStokesFlowSimulator3D
from .passive_transport_flow_simulators import (
    PassiveTransportFlowSimulator,
    create_unbounded_flow_simulator_2d,
    create_unbounded_flow_simulator_3d,
)

__all__ = [
    "FlowSimulator",
    "PassiveTransportFlowSimulator",
    "PeriodicNavierStokesFlowSimulator3D",
    "UnboundedNavierStokesFlowSimulator2D",
    "UnboundedNavierStokesFlowSimulator3D",
    "compute_advection_diffusion_stable_timestep",
    "create_unbounded_flow_simulator_2d",
    "create_unbounded_flow_simulator_3d",
]
