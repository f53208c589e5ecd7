"""Eulerian <-> Lagrangian grid communicators, 2D and 3D.

Drop-in counterparts of EulerianLagrangianGridCommunicator{2,3}D
(sopht/numeric/immersed_boundary_ops/EulerianLagrangianGridCommunicator3D.py:7-65, ...2D.py:7-65):
same constructor arguments, the same four callables with the same parameter names (positional or
keyword), same ValueErrors. Arrays are torch CUDA tensors (numpy arrays are staged through the device).
"""

from __future__ import annotations

import ctypes
from typing import Any

import numpy as np
import torch

from sopht_b200 import _lib


def _pos_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float64:
        return _lib.SOPHT_F64
    if t.dtype == torch.float32:
        return _lib.SOPHT_F32
    msg = f"Lagrangian positions must be float32 or float64, got {t.dtype}"
    raise ValueError(msg)


class _EulerianLagrangianGridCommunicator:
    _dim: int

    def __init__(
        self,
        dx: float,
        eul_grid_coord_shift: float,
        num_lag_nodes: int,
        interp_kernel_width: int,
        real_t: type,
        n_components: int = 1,
        interp_kernel_type: str = "cosine",
    ) -> None:
        dim = self._dim
        dt = _lib.dtype_code(real_t)
        # ...3D.py:194-196 / :331-333
        if n_components not in (1, dim):
            msg = f"Invalid number of components for interpolation, must be either 1 or {dim}"
            raise ValueError(msg)
        if interp_kernel_type not in ("cosine", "peskin"):
            msg = (
                "Invalid interpolation kernel type. Current supported types are"
                "'cosine' and 'peskin'."
            )
            raise ValueError(msg)
        # ...3D.py:397-399 / :425-427
        if interp_kernel_width != 2:
            msg = "Interpolation kernel inconsistent with interpolation kernel width!"
            raise ValueError(msg)
        kind = 0 if interp_kernel_type == "cosine" else 1
        # scalars evaluated exactly like the reference closures do (numpy scalar arithmetic on dx)
        weight_prefactor = float(
            real_t((0.25 / dx) ** dim) if kind == 0 else (0.125 / dx) ** dim
        )
        dx_pow_dim = float(dx**dim)
        dx_f, shift_f = float(dx), float(eul_grid_coord_shift)
        lib = _lib.load

        def local_eulerian_grid_support_of_lagrangian_grid_kernel(
            local_eul_grid_support_of_lag_grid: Any,
            nearest_eul_grid_index_to_lag_grid: Any,
            lag_positions: Any,
        ) -> None:
            """nearest index = floor((X - shift)/dx) and the 4^dim tap distances of every node."""
            with _lib.Staging() as s:
                sup = s.out(local_eul_grid_support_of_lag_grid)
                idx = s.out(nearest_eul_grid_index_to_lag_grid)
                pos = s.inp(lag_positions)
                if idx.dtype != torch.int64:
                    msg = "nearest_eul_grid_index_to_lag_grid must be an int64 array"
                    raise ValueError(msg)
                fs, fi, fp = _lib.field_desc(sup, dt), _lib.raw_desc(idx), _lib.raw_desc(pos)
                _lib.check(lib().sopht_ib_local_support(
                    dt, dim, ctypes.byref(fs), ctypes.byref(fi), ctypes.byref(fp), _pos_code(pos),
                    dx_f, shift_f, _lib.current_stream()))

        def interpolation_weights_kernel(
            interp_weights: Any, local_eul_grid_support_of_lag_grid: Any
        ) -> None:
            """Delta-kernel weights from the local support (which is rescaled in place, as in the reference)."""
            with _lib.Staging() as s:
                w = s.out(interp_weights)
                sup = s.out(local_eul_grid_support_of_lag_grid)
                fw, fs = _lib.field_desc(w, dt), _lib.field_desc(sup, dt)
                _lib.check(lib().sopht_ib_interpolation_weights(
                    dt, dim, kind, ctypes.byref(fw), ctypes.byref(fs), dx_f, weight_prefactor,
                    _lib.current_stream()))

        def eulerian_to_lagrangian_grid_interpolation_kernel(
            lag_grid_field: Any,
            eul_grid_field: Any,
            interp_weights: Any,
            nearest_eul_grid_index_to_lag_grid: Any,
        ) -> None:
            """Interpolate an Eulerian (scalar or vector) field onto the Lagrangian nodes."""
            with _lib.Staging() as s:
                lag, eul = s.out(lag_grid_field), s.inp(eul_grid_field)
                w, idx = s.inp(interp_weights), s.inp(nearest_eul_grid_index_to_lag_grid)
                fl, fe = _lib.field_desc(lag, dt), _lib.field_desc(eul, dt)
                fw, fi = _lib.field_desc(w, dt), _lib.raw_desc(idx)
                _lib.check(lib().sopht_ib_eulerian_to_lagrangian(
                    dt, dim, ctypes.byref(fl), ctypes.byref(fe), ctypes.byref(fw), ctypes.byref(fi),
                    dx_pow_dim, _lib.current_stream()))

        def lagrangian_to_eulerian_grid_interpolation_kernel(
            eul_grid_field: Any,
            lag_grid_field: Any,
            interp_weights: Any,
            nearest_eul_grid_index_to_lag_grid: Any,
        ) -> None:
            """Spread (accumulate) a Lagrangian field onto the Eulerian grid."""
            with _lib.Staging() as s:
                eul, lag = s.out(eul_grid_field), s.inp(lag_grid_field)
                w, idx = s.inp(interp_weights), s.inp(nearest_eul_grid_index_to_lag_grid)
                fl, fe = _lib.field_desc(lag, dt), _lib.field_desc(eul, dt)
                fw, fi = _lib.field_desc(w, dt), _lib.raw_desc(idx)
                _lib.check(lib().sopht_ib_lagrangian_to_eulerian(
                    dt, dim, ctypes.byref(fe), ctypes.byref(fl), ctypes.byref(fw), ctypes.byref(fi),
                    _lib.current_stream()))

        self.local_eulerian_grid_support_of_lagrangian_grid_kernel = (
            local_eulerian_grid_support_of_lagrangian_grid_kernel
        )
        self.interpolation_weights_kernel = interpolation_weights_kernel
        self.eulerian_to_lagrangian_grid_interpolation_kernel = (
            eulerian_to_lagrangian_grid_interpolation_kernel
        )
        self.lagrangian_to_eulerian_grid_interpolation_kernel = (
            lagrangian_to_eulerian_grid_interpolation_kernel
        )
        # constants the fused virtual-boundary kernel reuses
        self._consts = dict(dt=dt, dim=dim, dx=dx_f, shift=shift_f, weight_prefactor=weight_prefactor,
                            dx_pow_dim=dx_pow_dim, kind=kind)


class EulerianLagrangianGridCommunicator2D(_EulerianLagrangianGridCommunicator):
    """Communication between Eulerian and Lagrangian grids in 2D."""

    _dim = 2


class EulerianLagrangianGridCommunicator3D(_EulerianLagrangianGridCommunicator):
    """Communication between Eulerian and Lagrangian grids in 3D."""

    _dim = 3
