"""Virtual boundary forcing on a z-slab decomposed grid (IB row).

The Lagrangian arrays ((dim, N) positions, velocities, mismatches, forces) are replicated on every rank.
Each rank gathers only the taps that fall on planes it owns (the kernels skip taps outside the view they
are given), the partial flow velocities are summed with one all-reduce of 3 N values, the O(N) force
update is recomputed identically everywhere and every rank spreads only onto its own planes - the
scatter needs no exchange. Step order as VirtualBoundaryForcing.py:187-253.
"""

from __future__ import annotations

from typing import Any

import torch
import torch.distributed as dist

from sopht_b200 import _lib
from sopht_b200.numeric.immersed_boundary_ops import VirtualBoundaryForcing

from .slab import SlabPartition


class SlabVirtualBoundaryForcing(VirtualBoundaryForcing):
    """Same constructor as VirtualBoundaryForcing plus the slab partition; Eulerian arguments of the
    interaction methods are this rank's OWNED planes (3, nz/P, ny, nx) (views of halo-padded arrays)."""

    def __init__(self, *args: Any, partition: SlabPartition, group: Any = None, **kwargs: Any) -> None:
        kwargs["fused"] = False  # the gather must be reduced across ranks before the force is formed
        kwargs.setdefault("enable_eul_grid_forcing_reset", False)
        super().__init__(*args, **kwargs)
        if self.grid_dim != 3:
            msg = "slab decomposition is 3-D only"
            raise ValueError(msg)
        self.part, self.group = partition, group
        self._local_index = torch.empty_like(self.nearest_eul_grid_index_to_lag_grid)

    def compute_interaction_force_on_lag_grid(
        self, eul_grid_velocity_field: Any, lag_grid_position_field: Any, lag_grid_velocity_field: Any
    ) -> None:
        comm = self.eul_lag_grid_communicator
        comm.local_eulerian_grid_support_of_lagrangian_grid_kernel(
            local_eul_grid_support_of_lag_grid=self.local_eul_grid_support_of_lag_grid,
            nearest_eul_grid_index_to_lag_grid=self.nearest_eul_grid_index_to_lag_grid,
            lag_positions=lag_grid_position_field)
        comm.interpolation_weights_kernel(
            interp_weights=self.interp_weights,
            local_eul_grid_support_of_lag_grid=self.local_eul_grid_support_of_lag_grid)
        # global -> slab-local plane index (row 2 of the index array is z)
        self._local_index.copy_(self.nearest_eul_grid_index_to_lag_grid)
        self._local_index[2] -= self.part.z_start
        comm.eulerian_to_lagrangian_grid_interpolation_kernel(
            lag_grid_field=self.lag_grid_flow_velocity_field, eul_grid_field=eul_grid_velocity_field,
            interp_weights=self.interp_weights, nearest_eul_grid_index_to_lag_grid=self._local_index)
        if self.part.world_size > 1:
            with _lib.profile_range("comm.ib_allreduce"):
                dist.all_reduce(self.lag_grid_flow_velocity_field, op=dist.ReduceOp.SUM, group=self.group)
        self.compute_lag_grid_velocity_mismatch_field(
            self.lag_grid_velocity_mismatch_field, self.lag_grid_flow_velocity_field, lag_grid_velocity_field)
        self.compute_lag_grid_forcing_field(
            self.lag_grid_forcing_field, self.lag_grid_position_mismatch_field,
            self.lag_grid_velocity_mismatch_field, self.virtual_boundary_stiffness_coeff,
            self.virtual_boundary_damping_coeff)

    def compute_interaction_force_on_eul_and_lag_grid(
        self, eul_grid_forcing_field: Any, eul_grid_velocity_field: Any, lag_grid_position_field: Any,
        lag_grid_velocity_field: Any,
    ) -> None:
        self.compute_interaction_force_on_lag_grid(
            eul_grid_velocity_field, lag_grid_position_field, lag_grid_velocity_field)
        self.eul_lag_grid_communicator.lagrangian_to_eulerian_grid_interpolation_kernel(
            eul_grid_field=eul_grid_forcing_field, lag_grid_field=self.lag_grid_forcing_field,
            interp_weights=self.interp_weights, nearest_eul_grid_index_to_lag_grid=self._local_index)
