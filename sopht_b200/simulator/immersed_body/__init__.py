"""Immersed bodies on the device (sopht/simulator/immersed_body/__init__.py:3-43): forcing grids of
rigid bodies and Cosserat rods, the body <-> flow interaction objects and the pyelastica forcing wrapper."""

from .cosserat_rod_forcing_grids import (
    CosseratRodEdgeForcingGrid,
    CosseratRodElementCentricForcingGrid,
    CosseratRodNodalForcingGrid,
    CosseratRodState,
    CosseratRodSurfaceForcingGrid,
)
from .immersed_body_flow_interaction import (
    CosseratRodFlowInteraction,
    FlowForces,
    ImmersedBodyFlowInteraction,
    RigidBodyFlowInteraction,
)
from .rigid_body_forcing_grids import (
    CircularCylinderForcingGrid,
    ImmersedBodyForcingGrid,
    OpenEndCircularCylinderForcingGrid,
    RectangularPlane,
    RectangularPlaneForcingGrid,
    RigidBodyState,
    SphereForcingGrid,
    ThreeDimensionalRigidBodyForcingGrid,
    TwoDimensionalCylinderForcingGrid,
)

__all__ = [
    "CircularCylinderForcingGrid",
    "CosseratRodEdgeForcingGrid",
    "CosseratRodElementCentricForcingGrid",
    "CosseratRodFlowInteraction",
    "CosseratRodNodalForcingGrid",
    "CosseratRodState",
    "CosseratRodSurfaceForcingGrid",
    "FlowForces",
    "ImmersedBodyFlowInteraction",
    "ImmersedBodyForcingGrid",
    "OpenEndCircularCylinderForcingGrid",
    "RectangularPlane",
    "RectangularPlaneForcingGrid",
    "RigidBodyFlowInteraction",
    "RigidBodyState",
    "SphereForcingGrid",
    "ThreeDimensionalRigidBodyForcingGrid",
    "TwoDimensionalCylinderForcingGrid",
]
