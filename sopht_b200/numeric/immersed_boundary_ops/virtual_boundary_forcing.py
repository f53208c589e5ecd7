"""Virtual boundary forcing (Goldstein 1993 penalty force) for flow-body feedback.

Drop-in counterpart of sopht/numeric/immersed_boundary_ops/VirtualBoundaryForcing.py:20-283: same
constructor, buffers (as torch CUDA tensors), methods and call order. With the cosine kernel the six
sub-steps of ``compute_interaction_force_on_eul_and_lag_grid`` (:187-253) run as ONE fused CUDA launch
(sopht_ib_virtual_boundary_forcing); ``fused=False`` keeps the reference's kernel-by-kernel sequence.
"""

from __future__ import annotations

import ctypes
from typing import Any

import numpy as np
import torch

from sopht_b200 import _lib
from sopht_b200.numeric.eulerian_grid_ops import (
    gen_set_fixed_val_pyst_kernel_2d,
    gen_set_fixed_val_pyst_kernel_3d,
)

from .eulerian_lagrangian_grid_communicator import (
    EulerianLagrangianGridCommunicator2D,
    EulerianLagrangianGridCommunicator3D,
    _pos_code,
)


class VirtualBoundaryForcing:
    """Feedback between a Lagrangian body and the Eulerian grid flow via the virtual boundary method."""

    def __init__(
        self,
        virtual_boundary_stiffness_coeff: float,
        virtual_boundary_damping_coeff: float,
        grid_dim: int,
        dx: float,
        num_lag_nodes: int,
        real_t: type,
        eul_grid_coord_shift: float | None = None,
        interp_kernel_width: int | None = None,
        enable_eul_grid_forcing_reset: bool = True,
        num_threads: int | bool = False,
        start_time: float = 0.0,
        fused: bool = True,
    ) -> None:
        if grid_dim not in (2, 3):
            msg = "Invalid grid dimensions, must be either 2 or 3"
            raise ValueError(msg)
        self.grid_dim = grid_dim
        self.virtual_boundary_stiffness_coeff = virtual_boundary_stiffness_coeff
        self.virtual_boundary_damping_coeff = virtual_boundary_damping_coeff
        self.time = start_time
        self.real_t = real_t
        if eul_grid_coord_shift is None:
            eul_grid_coord_shift = real_t(dx / 2)
        if interp_kernel_width is None:
            interp_kernel_width = 2
        self._dt = _lib.dtype_code(real_t)
        tt = _lib.torch_dtype(real_t)
        if not torch.cuda.is_available():
            msg = "sopht_b200 immersed-boundary kernels need a CUDA device (no CPU fallback)"
            raise _lib.SophtLibraryError(msg)
        dev = torch.device("cuda", torch.cuda.current_device())
        taps = (2 * interp_kernel_width,) * grid_dim
        # buffers (VirtualBoundaryForcing.py:79-91)
        self.nearest_eul_grid_index_to_lag_grid = torch.empty(
            (grid_dim, num_lag_nodes), dtype=torch.int64, device=dev)
        self.local_eul_grid_support_of_lag_grid = torch.empty(
            (grid_dim, *taps, num_lag_nodes), dtype=tt, device=dev)
        self.interp_weights = torch.empty((*taps, num_lag_nodes), dtype=tt, device=dev)
        self.lag_grid_flow_velocity_field = torch.zeros((grid_dim, num_lag_nodes), dtype=tt, device=dev)
        self.lag_grid_position_mismatch_field = torch.zeros_like(self.lag_grid_flow_velocity_field)
        self.lag_grid_velocity_mismatch_field = torch.zeros_like(self.lag_grid_flow_velocity_field)
        self.lag_grid_forcing_field = torch.zeros_like(self.lag_grid_flow_velocity_field)

        comm_cls = EulerianLagrangianGridCommunicator2D if grid_dim == 2 else EulerianLagrangianGridCommunicator3D
        self.eul_lag_grid_communicator = comm_cls(
            dx=dx, eul_grid_coord_shift=eul_grid_coord_shift, num_lag_nodes=num_lag_nodes,
            interp_kernel_width=interp_kernel_width, real_t=real_t, n_components=grid_dim)
        self._fused = bool(fused)
        if enable_eul_grid_forcing_reset:
            gen = gen_set_fixed_val_pyst_kernel_2d if grid_dim == 2 else gen_set_fixed_val_pyst_kernel_3d
            self.set_eul_grid_vector_field = gen(real_t=real_t, num_threads=num_threads, field_type="vector")
            self.compute_interaction_forcing = (
                self.compute_interaction_force_on_eul_and_lag_grid_with_eul_grid_forcing_reset
            )
        else:
            self.compute_interaction_forcing = self.compute_interaction_force_on_eul_and_lag_grid

    # ---- O(N_lag) Lagrangian updates (VirtualBoundaryForcing.py:132-185), elementwise CUDA kernels ----
    def compute_lag_grid_velocity_mismatch_field(
        self, lag_grid_velocity_mismatch_field, lag_grid_flow_velocity_field, lag_grid_body_velocity_field
    ) -> None:
        body = lag_grid_body_velocity_field.to(lag_grid_flow_velocity_field.dtype)
        _lib.call("sopht_elementwise_saxpby", self._dt, lag_grid_velocity_mismatch_field,
                  lag_grid_flow_velocity_field, body, 1.0, -1.0)

    def update_lag_grid_position_mismatch_field_via_euler_forward(
        self, lag_grid_position_mismatch_field, lag_grid_velocity_mismatch_field, dt
    ) -> None:
        _lib.call("sopht_elementwise_saxpby", self._dt, lag_grid_position_mismatch_field,
                  lag_grid_position_mismatch_field, lag_grid_velocity_mismatch_field, 1.0, dt)

    def compute_lag_grid_forcing_field(
        self, lag_grid_forcing_field, lag_grid_position_mismatch_field, lag_grid_velocity_mismatch_field,
        virtual_boundary_stiffness_coeff, virtual_boundary_damping_coeff,
    ) -> None:
        _lib.call("sopht_elementwise_saxpby", self._dt, lag_grid_forcing_field,
                  lag_grid_position_mismatch_field, lag_grid_velocity_mismatch_field,
                  virtual_boundary_stiffness_coeff, virtual_boundary_damping_coeff)

    # ---- interaction -------------------------------------------------------------------------------------
    def _fused_interaction(self, eul_grid_forcing_field, eul_grid_velocity_field, lag_pos, lag_vel) -> None:
        c = self.eul_lag_grid_communicator._consts
        dt = self._dt
        with _lib.Staging() as s:
            vel = s.inp(eul_grid_velocity_field)
            pos = s.inp(lag_pos)
            bvel = s.inp(lag_vel)
            if bvel.dtype != pos.dtype:
                bvel = bvel.to(pos.dtype)
            ff = None
            if eul_grid_forcing_field is not None:
                frc = s.out(eul_grid_forcing_field)
                ff = _lib.field_desc(frc, dt)
            fv = _lib.field_desc(vel, dt)
            fp, fb = _lib.raw_desc(pos), _lib.raw_desc(bvel)
            d = [
                _lib.field_desc(self.local_eul_grid_support_of_lag_grid, dt),
                _lib.field_desc(self.interp_weights, dt),
                _lib.raw_desc(self.nearest_eul_grid_index_to_lag_grid),
                _lib.field_desc(self.lag_grid_flow_velocity_field, dt),
                _lib.field_desc(self.lag_grid_velocity_mismatch_field, dt),
                _lib.field_desc(self.lag_grid_position_mismatch_field, dt),
                _lib.field_desc(self.lag_grid_forcing_field, dt),
            ]
            _lib.check(_lib.load().sopht_ib_virtual_boundary_forcing(
                dt, c["dim"], ctypes.byref(ff) if ff is not None else None, ctypes.byref(fv),
                ctypes.byref(fp), ctypes.byref(fb), _pos_code(pos), *[ctypes.byref(x) for x in d],
                c["dx"], c["shift"], c["weight_prefactor"], c["dx_pow_dim"],
                float(self.virtual_boundary_stiffness_coeff), float(self.virtual_boundary_damping_coeff),
                _lib.current_stream()))

    def compute_interaction_force_on_lag_grid(
        self, eul_grid_velocity_field: Any, lag_grid_position_field: Any, lag_grid_velocity_field: Any
    ) -> None:
        """Virtual boundary: compute interaction force on Lagrangian grid (:187-230)."""
        if self._fused:
            self._fused_interaction(None, eul_grid_velocity_field, lag_grid_position_field,
                                    lag_grid_velocity_field)
            return
        comm = self.eul_lag_grid_communicator
        with _lib.Staging() as s:
            vel, pos, bvel = s.inp(eul_grid_velocity_field), s.inp(lag_grid_position_field), s.inp(lag_grid_velocity_field)
            comm.local_eulerian_grid_support_of_lagrangian_grid_kernel(
                local_eul_grid_support_of_lag_grid=self.local_eul_grid_support_of_lag_grid,
                nearest_eul_grid_index_to_lag_grid=self.nearest_eul_grid_index_to_lag_grid,
                lag_positions=pos)
            comm.interpolation_weights_kernel(
                interp_weights=self.interp_weights,
                local_eul_grid_support_of_lag_grid=self.local_eul_grid_support_of_lag_grid)
            comm.eulerian_to_lagrangian_grid_interpolation_kernel(
                lag_grid_field=self.lag_grid_flow_velocity_field, eul_grid_field=vel,
                interp_weights=self.interp_weights,
                nearest_eul_grid_index_to_lag_grid=self.nearest_eul_grid_index_to_lag_grid)
            self.compute_lag_grid_velocity_mismatch_field(
                self.lag_grid_velocity_mismatch_field, self.lag_grid_flow_velocity_field, bvel)
            self.compute_lag_grid_forcing_field(
                self.lag_grid_forcing_field, self.lag_grid_position_mismatch_field,
                self.lag_grid_velocity_mismatch_field, self.virtual_boundary_stiffness_coeff,
                self.virtual_boundary_damping_coeff)

    def compute_interaction_force_on_eul_and_lag_grid(
        self, eul_grid_forcing_field: Any, eul_grid_velocity_field: Any, lag_grid_position_field: Any,
        lag_grid_velocity_field: Any,
    ) -> None:
        """Virtual boundary: compute interaction on Eulerian grid (:232-253); accumulates into the forcing field."""
        if self._fused:
            self._fused_interaction(eul_grid_forcing_field, eul_grid_velocity_field,
                                    lag_grid_position_field, lag_grid_velocity_field)
            return
        self.compute_interaction_force_on_lag_grid(
            eul_grid_velocity_field, lag_grid_position_field, lag_grid_velocity_field)
        self.eul_lag_grid_communicator.lagrangian_to_eulerian_grid_interpolation_kernel(
            eul_grid_field=eul_grid_forcing_field, lag_grid_field=self.lag_grid_forcing_field,
            interp_weights=self.interp_weights,
            nearest_eul_grid_index_to_lag_grid=self.nearest_eul_grid_index_to_lag_grid)

    def compute_interaction_force_on_eul_and_lag_grid_with_eul_grid_forcing_reset(
        self, eul_grid_forcing_field: Any, eul_grid_velocity_field: Any, lag_grid_position_field: Any,
        lag_grid_velocity_field: Any,
    ) -> None:
        """Same, after resetting eul_grid_forcing_field to zero (:255-274)."""
        self.set_eul_grid_vector_field(vector_field=eul_grid_forcing_field, fixed_vals=([0] * self.grid_dim))
        self.compute_interaction_force_on_eul_and_lag_grid(
            eul_grid_forcing_field, eul_grid_velocity_field, lag_grid_position_field, lag_grid_velocity_field)

    def time_step(self, dt: float) -> None:
        """Virtual boundary forcing time step, updates grid deviation (:276-283)."""
        self.update_lag_grid_position_mismatch_field_via_euler_forward(
            lag_grid_position_mismatch_field=self.lag_grid_position_mismatch_field,
            lag_grid_velocity_mismatch_field=self.lag_grid_velocity_mismatch_field, dt=dt)
        self.time += dt
