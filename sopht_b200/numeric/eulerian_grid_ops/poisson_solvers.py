"""Unbounded Poisson solver classes (2-D and 3-D).

Drop-in counterparts of UnboundedPoissonSolverPYFFTW{2,3}D
(sopht/numeric/eulerian_grid_ops/poisson_solver_3d/UnboundedPoissonSolverPYFFTW3D.py:9-172,
poisson_solver_2d/UnboundedPoissonSolverPYFFTW2D.py:8-129): same constructor arguments and
``solve`` / ``vector_field_solve`` keywords. Plans, workspaces and the Green's function spectrum live
behind an opaque C handle (sopht_poisson_create); fields stay caller-owned torch CUDA tensors (numpy
arrays are staged through the device).
"""

from __future__ import annotations

import ctypes
from typing import Any

import numpy as np
import torch

from sopht_b200 import _lib

POISSON_AUTO = 0
POISSON_FORCE_GENERIC = 1


def _reflected_axis(axis_range: float, dx: Any, grid_size: int, real_t: type) -> np.ndarray:
    """min(x, 2X - x) on the doubled axis, with the reference's numpy expressions and dtypes
    (UnboundedPoissonSolverPYFFTW3D.py:58-75)."""
    x_double = np.linspace(0, 2 * axis_range - dx, 2 * grid_size).astype(real_t)
    return np.minimum(x_double, 2 * axis_range - x_double)


class _UnboundedPoissonSolverBase:
    _dim: int

    def _create(self, flags: int) -> None:
        lib = _lib.load()
        dt = _lib.dtype_code(self.real_t)
        real_t = self.real_t
        mx = _reflected_axis(self.x_range, self.dx, self.grid_size_x, real_t)
        my = _reflected_axis(self.y_range, self.dx, self.grid_size_y, real_t)
        if self._dim == 3:
            mz = _reflected_axis(self.z_range, self.dx, self.grid_size_z, real_t)
            # Regularization term (straight from PPM), UnboundedPoissonSolverPYFFTW3D.py:79
            origin = real_t(1 / (4 * np.pi * self.dx))
            nz = self.grid_size_z
        else:
            mz = None
            # UnboundedPoissonSolverPYFFTW2D.py:64
            origin = real_t(-(2 * np.log(self.dx / np.sqrt(np.pi)) - 1) / (4 * np.pi))
            nz = 1
        if not torch.cuda.is_available():
            msg = "sopht_b200 Poisson solver needs a CUDA device (no CPU fallback)"
            raise _lib.SophtLibraryError(msg)
        handle = ctypes.c_void_p()
        _lib.check(
            lib.sopht_poisson_create(
                ctypes.byref(handle),
                dt,
                self._dim,
                nz,
                self.grid_size_y,
                self.grid_size_x,
                float(self.x_range),
                float(self.dx),
                _lib.double_array(mz) if mz is not None else None,
                _lib.double_array(my),
                _lib.double_array(mx),
                float(origin),
                flags,
                _lib.current_stream(),
            )
        )
        self._handle = handle
        self._dt = dt
        self.path = lib.sopht_poisson_path(handle).decode()

    def __del__(self) -> None:
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                _lib.load().sopht_poisson_destroy(h)
            except Exception:  # noqa: BLE001 - interpreter shutdown
                pass
            self._handle = None

    def _solve(self, solution: Any, rhs: Any) -> None:
        with _lib.Staging() as s:
            r, o = s.inp(rhs), s.out(solution)
            fo, fr = _lib.field_desc(o, self._dt), _lib.field_desc(r, self._dt)
            _lib.check(
                _lib.load().sopht_poisson_solve(
                    self._handle, ctypes.byref(fo), ctypes.byref(fr), _lib.current_stream()
                )
            )

    def solve(self, solution_field: Any, rhs_field: Any) -> None:
        """Solve -del^2(solution_field) = rhs_field on the unbounded domain (Hockney-Eastwood)."""
        self._solve(solution_field, rhs_field)

    def _greens_function_hat(self) -> torch.Tensor:
        """rfftn(G) * dx^dim on the doubled grid as a complex tensor (2nz, 2ny, nx + 1) / (2ny, nx + 1), rebuilt from
        the real, pre-normalised copy the library keeps (G is real and even, so its transform is real). The reference
        holds this array as `fourier_greens_function_times_dx_cubed` / `..._squared`
        (UnboundedPoissonSolverPYFFTW3D.py:47-49, ...2D.py:42-44) and multiplies the spectrum with it."""
        ptr = ctypes.c_void_p()
        _lib.check(_lib.load().sopht_poisson_green_hat(self._handle, ctypes.byref(ptr)))
        if not ptr.value:
            msg = "this solver path does not keep the Green's function spectrum in natural order"
            raise _lib.SophtLibraryError(msg)
        shape = ((2 * self.grid_size_z,) if self._dim == 3 else ()) + (2 * self.grid_size_y, self.grid_size_x + 1)
        count = int(np.prod(shape))

        class _Span:  # raw device memory -> torch (zero copy); the solver object keeps the memory alive
            pass

        span = _Span()
        span.owner = self
        span.__cuda_array_interface__ = {
            "shape": (count,), "typestr": np.dtype(self.real_t).str, "data": (ptr.value, False), "version": 2,
            "strides": None}
        real = torch.as_tensor(span, device=torch.device("cuda", torch.cuda.current_device())).view(*shape)
        doubled_cells = float(np.prod(shape[:-1]) * 2 * self.grid_size_x)  # shape[:-1] is already doubled
        return torch.complex(real * doubled_cells, torch.zeros_like(real))


class UnboundedPoissonSolverPYFFTW3D(_UnboundedPoissonSolverBase):
    """3-D unbounded Poisson solver (same ctor as the reference class of this name)."""

    _dim = 3

    def __init__(
        self,
        grid_size_z: int,
        grid_size_y: int,
        grid_size_x: int,
        x_range: float = 1.0,
        num_threads: int = 1,
        real_t: type = np.float64,
        flags: int = POISSON_AUTO,
    ) -> None:
        self.grid_size_z = grid_size_z
        self.grid_size_y = grid_size_y
        self.grid_size_x = grid_size_x
        self.x_range = x_range
        self.y_range = x_range * (grid_size_y / grid_size_x)
        self.z_range = x_range * (grid_size_z / grid_size_x)
        self.dx = real_t(x_range / grid_size_x)
        self.num_threads = num_threads
        self.real_t = real_t
        self.x_axis_idx, self.y_axis_idx, self.z_axis_idx = 0, 1, 2
        self._create(flags)

    def vector_field_solve(self, solution_vector_field: Any, rhs_vector_field: Any) -> None:
        """Three component solves, -del^2(solution_vector_field) = rhs_vector_field."""
        self._solve(solution_vector_field, rhs_vector_field)

    @property
    def fourier_greens_function_times_dx_cubed(self) -> torch.Tensor:
        return self._greens_function_hat()


class UnboundedPoissonSolverPYFFTW2D(_UnboundedPoissonSolverBase):
    """2-D unbounded Poisson solver (same ctor as the reference class of this name)."""

    _dim = 2

    def __init__(
        self,
        grid_size_y: int,
        grid_size_x: int,
        x_range: float = 1.0,
        num_threads: int = 1,
        real_t: type = np.float64,
        flags: int = POISSON_AUTO,
    ) -> None:
        self.grid_size_y = grid_size_y
        self.grid_size_x = grid_size_x
        self.x_range = x_range
        self.y_range = x_range * (grid_size_y / grid_size_x)
        self.dx = real_t(x_range / grid_size_x)
        self.num_threads = num_threads
        self.real_t = real_t
        self._create(flags)

    @property
    def fourier_greens_function_times_dx_squared(self) -> torch.Tensor:
        return self._greens_function_hat()


class _FastDiagPoissonSolverBase(_UnboundedPoissonSolverBase):
    """Homogeneous-Neumann Poisson solve behind the same handle type (sopht_poisson_neumann_create). The reference
    diagonalises tridiag(-1, 2, -1) / dx^2 (corner entries 1 / dx^2) per axis with numpy.linalg.eig and applies
    V diag(1 / lambda) V^-1 as dense tensordots; that product does not depend on the eigenvector scaling or order, and
    its closed form (DCT-II basis) is evaluated here as a mirror extension + FFT, so the `eig_vecs_*` /
    `inv_of_eig_vecs_*` / `inv_eig_val_matrix` / `spectral_field_buffer` arrays of the reference do not exist."""

    def _create_neumann(self) -> None:
        if not torch.cuda.is_available():
            msg = "sopht_b200 Poisson solver needs a CUDA device (no CPU fallback)"
            raise _lib.SophtLibraryError(msg)
        lib = _lib.load()
        dt = _lib.dtype_code(self.real_t)
        handle = ctypes.c_void_p()
        nz = self.grid_size_z if self._dim == 3 else 1
        _lib.check(lib.sopht_poisson_neumann_create(
            ctypes.byref(handle), dt, self._dim, nz, self.grid_size_y, self.grid_size_x, float(self.dx),
            _lib.current_stream()))
        self._handle = handle
        self._dt = dt
        self.path = lib.sopht_poisson_path(handle).decode()

    def solve(self, solution_field: Any, rhs_field: Any) -> None:
        """Solve -del^2(solution_field) = rhs_field with zero normal derivative on every wall; the mean of the
        solution is zero (the reference drops the null mode, FastDiagPoissonSolver3D.py:143-146)."""
        self._solve(solution_field, rhs_field)


class FastDiagPoissonSolver3D(_FastDiagPoissonSolverBase):
    """FastDiagPoissonSolver3D.py:12-208 (same ctor, `solve`, `vector_field_solve`)."""

    _dim = 3

    def __init__(self, grid_size_z: int, grid_size_y: int, grid_size_x: int, dx: float, real_t: type = np.float64,
                 bc_type: str = "homogenous_neumann_along_xyz") -> None:
        if bc_type != "homogenous_neumann_along_xyz":
            msg = f"unsupported bc_type {bc_type!r}: only 'homogenous_neumann_along_xyz' is defined"
            raise ValueError(msg)
        self.dx = dx
        self.grid_size_z = grid_size_z
        self.grid_size_y = grid_size_y
        self.grid_size_x = grid_size_x
        self.real_t = real_t
        self.bc_type = bc_type
        self.x_axis_idx, self.y_axis_idx, self.z_axis_idx = 0, 1, 2
        self._create_neumann()

    def vector_field_solve(self, solution_vector_field: Any, rhs_vector_field: Any) -> None:
        """Three component solves (:183-208)."""
        self._solve(solution_vector_field, rhs_vector_field)


class FastDiagPoissonSolver2D(_FastDiagPoissonSolverBase):
    """FastDiagPoissonSolver2D.py:10-119 (same ctor and `solve`)."""

    _dim = 2

    def __init__(self, grid_size_y: int, grid_size_x: int, dx: float, real_t: type = np.float64,
                 bc_type: str = "homogenous_neumann_along_xy") -> None:
        if bc_type != "homogenous_neumann_along_xy":
            msg = f"unsupported bc_type {bc_type!r}: only 'homogenous_neumann_along_xy' is defined"
            raise ValueError(msg)
        self.dx = dx
        self.grid_size_y = grid_size_y
        self.grid_size_x = grid_size_x
        self.real_t = real_t
        self.bc_type = bc_type
        self._create_neumann()


class _PeriodicPoissonSolverBase(_UnboundedPoissonSolverBase):
    """Periodic Poisson solve behind the same handle type (sopht_poisson_periodic_create). An extension for BASELINE
    config 4 (periodic Taylor-Green vortex): the reference has no periodic solver, so the
    constructor mirrors UnboundedPoissonSolverPYFFTW{2,3}D's and parity is against analytic Fourier modes
    (**parity unpinned** by the reference)."""

    def _create_periodic(self, symbol: str) -> None:
        if symbol not in ("spectral", "three_point"):
            msg = "symbol must be 'spectral' or 'three_point'"
            raise ValueError(msg)
        if not torch.cuda.is_available():
            msg = "sopht_b200 Poisson solver needs a CUDA device (no CPU fallback)"
            raise _lib.SophtLibraryError(msg)
        self.symbol = symbol
        lib = _lib.load()
        dt = _lib.dtype_code(self.real_t)
        handle = ctypes.c_void_p()
        nz = self.grid_size_z if self._dim == 3 else 1
        _lib.check(lib.sopht_poisson_periodic_create(
            ctypes.byref(handle), dt, self._dim, nz, self.grid_size_y, self.grid_size_x, float(self.dx),
            int(symbol == "three_point"), _lib.current_stream()))
        self._handle = handle
        self._dt = dt
        self.path = lib.sopht_poisson_path(handle).decode()

    def solve(self, solution_field: Any, rhs_field: Any) -> None:
        """Solve -del^2(solution_field) = rhs_field on the periodic box; the mean of the solution is zero (the mean
        of rhs_field, which a periodic problem cannot balance, is dropped)."""
        self._solve(solution_field, rhs_field)


class PeriodicPoissonSolver3D(_PeriodicPoissonSolverBase):
    """3-D periodic Poisson solver; `symbol="spectral"`: (2 pi m / L)^2, `"three_point"`: exact inverse of the
    7-point Laplacian with wrap-around neighbours."""

    _dim = 3

    def __init__(self, grid_size_z: int, grid_size_y: int, grid_size_x: int, x_range: float = 1.0,
                 num_threads: int = 1, real_t: type = np.float64, symbol: str = "spectral") -> None:
        self.grid_size_z, self.grid_size_y, self.grid_size_x = grid_size_z, grid_size_y, grid_size_x
        self.x_range = x_range
        self.y_range = x_range * (grid_size_y / grid_size_x)
        self.z_range = x_range * (grid_size_z / grid_size_x)
        self.dx = real_t(x_range / grid_size_x)
        self.num_threads = num_threads
        self.real_t = real_t
        self.x_axis_idx, self.y_axis_idx, self.z_axis_idx = 0, 1, 2
        self._create_periodic(symbol)

    def vector_field_solve(self, solution_vector_field: Any, rhs_vector_field: Any) -> None:
        self._solve(solution_vector_field, rhs_vector_field)


class PeriodicPoissonSolver2D(_PeriodicPoissonSolverBase):
    """2-D periodic Poisson solver."""

    _dim = 2

    def __init__(self, grid_size_y: int, grid_size_x: int, x_range: float = 1.0, num_threads: int = 1,
                 real_t: type = np.float64, symbol: str = "spectral") -> None:
        self.grid_size_y, self.grid_size_x = grid_size_y, grid_size_x
        self.x_range = x_range
        self.y_range = x_range * (grid_size_y / grid_size_x)
        self.dx = real_t(x_range / grid_size_x)
        self.num_threads = num_threads
        self.real_t = real_t
        self._create_periodic(symbol)
