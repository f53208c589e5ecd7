"""Body <-> flow interaction objects: a forcing grid on the device plus VirtualBoundaryForcing.

Same class names, constructor arguments (positional order included), attributes and methods as
sopht/simulator/immersed_body/immersed_body_flow_interaction.py:15-141,
rigid_body/rigid_body_flow_interaction.py:9-60, cosserat_rod/cosserat_rod_flow_interaction.py:9-58 and
flow_forces.py:11-25. The Eulerian fields are torch CUDA tensors (the flow simulator's), the forcing grid's fields are
device tensors too, and `body_flow_forces` / `body_flow_torques` are the host numpy arrays pyelastica adds to its
external forces: they are the only per-step device -> host traffic.
"""

from __future__ import annotations

import logging
from typing import Any

import numpy as np
import torch

from sopht_b200.numeric.immersed_boundary_ops import VirtualBoundaryForcing

from .rigid_body_forcing_grids import ImmersedBodyForcingGrid

logger = logging.getLogger(__name__)

try:  # pyelastica is optional here: FlowForces only needs its NoForces base to be registered on a simulator
    from elastica import NoForces as _NoForces
except ImportError:  # pragma: no cover - the image has no pyelastica

    class _NoForces:  # type: ignore[no-redef]
        """Stand-in with the interface of elastica.NoForces."""

        def __init__(self) -> None:
            pass

        def apply_forces(self, system: Any, time: float = 0.0) -> None:
            pass

        def apply_torques(self, system: Any, time: float = 0.0) -> None:
            pass


class ImmersedBodyFlowInteraction(VirtualBoundaryForcing):
    """Base class for immersed body flow interaction (immersed_body_flow_interaction.py:15-141)."""

    def __init__(
        self, eul_grid_forcing_field: torch.Tensor, eul_grid_velocity_field: torch.Tensor,
        body_flow_forces: np.ndarray, body_flow_torques: np.ndarray,
        forcing_grid_cls: type[ImmersedBodyForcingGrid], virtual_boundary_stiffness_coeff: float,
        virtual_boundary_damping_coeff: float, dx: float, grid_dim: int, real_t: type = np.float64,
        eul_grid_coord_shift: float | None = None, interp_kernel_width: float | None = None,
        enable_eul_grid_forcing_reset: bool = False, num_threads: int | bool = False, start_time: float = 0.0,
        **forcing_grid_kwargs: Any,
    ) -> None:
        self.body_flow_forces = body_flow_forces
        self.body_flow_torques = body_flow_torques
        self.forcing_grid = forcing_grid_cls(grid_dim=grid_dim, **forcing_grid_kwargs)
        # references to the simulator's fields (:46-50); the velocity is only ever read here
        self.eul_grid_forcing_field = eul_grid_forcing_field
        self.eul_grid_velocity_field = eul_grid_velocity_field

        # relative resolution of the two grids (:52-82): the delta function's support is 2 cells
        max_lag_grid_dx = float(self.forcing_grid.get_maximum_lagrangian_grid_spacing())
        grid_type = type(self.forcing_grid).__name__
        if max_lag_grid_dx > 2 * dx:
            verdict = (
                f"\nMax Lagrangian grid spacing: {max_lag_grid_dx} > 2 * dx"
                "\nThe Lagrangian grid of the body is too coarse relative to"
                "\nthe Eulerian grid of the flow, which can lead to unexpected"
                "\nconvergence. Please make the Lagrangian grid finer."
            )
            log = logger.warning
        elif max_lag_grid_dx < 0.5 * dx:
            verdict = (
                f"\nMax Lagrangian grid spacing: {max_lag_grid_dx} < 0.5 * dx"
                "\nThe Lagrangian grid of the body is too fine relative to"
                "\nthe Eulerian grid of the flow, which corresponds to redundant"
                "\nforcing points. Please make the Lagrangian grid coarser."
            )
            log = logger.warning
        else:
            verdict = "\nLagrangian grid is resolved almost the same\nas the Eulerian grid of the flow."
            log = logger.info
        bar = "\n" + "=" * 50
        # the message text is pinned by the reference's test_immersed_body_interactor_warnings
        log(f"{bar}\nFor {grid_type}:\nEulerian grid spacing (dx): {dx}{verdict}{bar}")

        # penalty coefficients scale with the Lagrangian cell size (:84-87)
        virtual_boundary_stiffness_coeff *= max_lag_grid_dx ** (grid_dim - 1)
        virtual_boundary_damping_coeff *= max_lag_grid_dx ** (grid_dim - 1)

        super().__init__(
            virtual_boundary_stiffness_coeff, virtual_boundary_damping_coeff, grid_dim, dx,
            self.forcing_grid.num_lag_nodes, real_t, eul_grid_coord_shift, interp_kernel_width,
            enable_eul_grid_forcing_reset, num_threads, start_time)

    def compute_interaction_on_lag_grid(self) -> None:
        """Compute interaction forces on the Lagrangian forcing grid (:105-113)."""
        self.forcing_grid.compute_lag_grid_position_field()
        self.forcing_grid.compute_lag_grid_velocity_field()
        self.compute_interaction_force_on_lag_grid(
            eul_grid_velocity_field=self.eul_grid_velocity_field,
            lag_grid_position_field=self.forcing_grid.position_field,
            lag_grid_velocity_field=self.forcing_grid.velocity_field,
        )

    def __call__(self) -> None:
        """The full interaction: Lagrangian forcing and its spreading onto the Eulerian grid (:115-125)."""
        self.forcing_grid.compute_lag_grid_position_field()
        self.forcing_grid.compute_lag_grid_velocity_field()
        self.compute_interaction_forcing(
            eul_grid_forcing_field=self.eul_grid_forcing_field,
            eul_grid_velocity_field=self.eul_grid_velocity_field,
            lag_grid_position_field=self.forcing_grid.position_field,
            lag_grid_velocity_field=self.forcing_grid.velocity_field,
        )

    def compute_flow_forces_and_torques(self) -> None:
        """Flow forces and torques on the body from the forces on the Lagrangian grid (:127-134)."""
        self.compute_interaction_on_lag_grid()
        self.forcing_grid.transfer_forcing_from_grid_to_body(
            body_flow_forces=self.body_flow_forces,
            body_flow_torques=self.body_flow_torques,
            lag_grid_forcing_field=self.lag_grid_forcing_field,
        )

    def get_grid_deviation_error_l2_norm(self) -> float:
        """L2 norm of the deviation between the flow's and the body's grids (:136-141)."""
        norm = torch.linalg.vector_norm(self.lag_grid_position_mismatch_field.double())
        return float(norm) / np.sqrt(self.forcing_grid.num_lag_nodes)


class RigidBodyFlowInteraction(ImmersedBodyFlowInteraction):
    """Rigid body (pyelastica attribute names) <-> flow (rigid_body_flow_interaction.py:9-60)."""

    def __init__(
        self,
        rigid_body: Any,
        eul_grid_forcing_field: torch.Tensor, eul_grid_velocity_field: torch.Tensor,
        virtual_boundary_stiffness_coeff: float, virtual_boundary_damping_coeff: float, dx: float, grid_dim: int,
        forcing_grid_cls: type[ImmersedBodyForcingGrid], real_t: type = np.float64,
        eul_grid_coord_shift: float | None = None, interp_kernel_width: float | None = None,
        enable_eul_grid_forcing_reset: bool = False, num_threads: int | bool = False, start_time: float = 0.0,
        **forcing_grid_kwargs: Any,
    ) -> None:
        forcing_grid_kwargs["rigid_body"] = rigid_body
        super().__init__(
            eul_grid_forcing_field, eul_grid_velocity_field, np.zeros((3, 1)), np.zeros((3, 1)), forcing_grid_cls,
            virtual_boundary_stiffness_coeff, virtual_boundary_damping_coeff, dx, grid_dim, real_t,
            eul_grid_coord_shift, interp_kernel_width, enable_eul_grid_forcing_reset, num_threads, start_time,
            **forcing_grid_kwargs)


class CosseratRodFlowInteraction(ImmersedBodyFlowInteraction):
    """Cosserat rod <-> flow (cosserat_rod_flow_interaction.py:9-58)."""

    def __init__(
        self,
        cosserat_rod: Any,
        eul_grid_forcing_field: torch.Tensor, eul_grid_velocity_field: torch.Tensor,
        virtual_boundary_stiffness_coeff: float, virtual_boundary_damping_coeff: float, dx: float, grid_dim: int,
        forcing_grid_cls: type[ImmersedBodyForcingGrid], real_t: type = np.float64,
        eul_grid_coord_shift: float | None = None, interp_kernel_width: float | None = None,
        enable_eul_grid_forcing_reset: bool = False, num_threads: int | bool = False, start_time: float = 0.0,
        **forcing_grid_kwargs: Any,
    ) -> None:
        forcing_grid_kwargs["cosserat_rod"] = cosserat_rod
        super().__init__(
            eul_grid_forcing_field, eul_grid_velocity_field, np.zeros((3, cosserat_rod.n_elems + 1)),
            np.zeros((3, cosserat_rod.n_elems)), forcing_grid_cls, virtual_boundary_stiffness_coeff,
            virtual_boundary_damping_coeff, dx, grid_dim, real_t, eul_grid_coord_shift, interp_kernel_width,
            enable_eul_grid_forcing_reset, num_threads, start_time, **forcing_grid_kwargs)


class FlowForces(_NoForces):
    """pyelastica forcing that adds the flow's forces and torques to a body (flow_forces.py:11-25)."""

    def __init__(self, body_flow_interactor: CosseratRodFlowInteraction | RigidBodyFlowInteraction) -> None:
        super().__init__()
        self.body_flow_interactor = body_flow_interactor

    def apply_forces(self, system: Any, time: float = 0.0) -> None:
        self.body_flow_interactor.compute_flow_forces_and_torques()
        system.external_forces += self.body_flow_interactor.body_flow_forces
        system.external_torques += self.body_flow_interactor.body_flow_torques
