"""Immersed-body forcing grids on the device (sopht/simulator/immersed_body/)."""

from .rigid_body_forcing_grids import (
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
    "ImmersedBodyForcingGrid",
    "OpenEndCircularCylinderForcingGrid",
    "RigidBodyState",
    "SphereForcingGrid",
    "ThreeDimensionalRigidBodyForcingGrid",
    "TwoDimensionalCylinderForcingGrid",
]
