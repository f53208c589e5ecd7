"""Immersed boundary operations — same public names as sopht.numeric.immersed_boundary_ops
(sopht/numeric/immersed_boundary_ops/__init__.py:3-13), backed by sm_100a CUDA kernels (csrc/ib.cu)."""

from .eulerian_lagrangian_grid_communicator import (
    EulerianLagrangianGridCommunicator2D,
    EulerianLagrangianGridCommunicator3D,
)
from .virtual_boundary_forcing import VirtualBoundaryForcing

__all__ = [
    "EulerianLagrangianGridCommunicator2D",
    "EulerianLagrangianGridCommunicator3D",
    "VirtualBoundaryForcing",
]
