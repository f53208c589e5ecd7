"""Simulator-level call sites of the hot path (step sequencers, rigid-body forcing grids), same names as
sopht.simulator."""

from .flow import (
    FlowSimulator,
    PassiveTransportFlowSimulator,
    UnboundedNavierStokesFlowSimulator2D,
    UnboundedNavierStokesFlowSimulator3D,
    compute_advection_diffusion_stable_timestep,
    create_unbounded_flow_simulator_2d,
    create_unbounded_flow_simulator_3d,
)
from .immersed_body import (
    CircularCylinderForcingGrid,
    ImmersedBodyForcingGrid,
    OpenEndCircularCylinderForcingGrid,
    RigidBodyState,
    SphereForcingGrid,
    ThreeDimensionalRigidBodyForcingGrid,
    TwoDimensionalCylinderForcingGrid,
)

__all__ = [
    "CircularCylinderForcingGrid",
    "FlowSimulator",
    "ImmersedBodyForcingGrid",
    "OpenEndCircularCylinderForcingGrid",
    "PassiveTransportFlowSimulator",
    "RigidBodyState",
    "SphereForcingGrid",
    "ThreeDimensionalRigidBodyForcingGrid",
    "TwoDimensionalCylinderForcingGrid",
    "UnboundedNavierStokesFlowSimulator2D",
    "UnboundedNavierStokesFlowSimulator3D",
    "compute_advection_diffusion_stable_timestep",
    "create_unbounded_flow_simulator_2d",
    "create_unbounded_flow_simulator_3d",
]
