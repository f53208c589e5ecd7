"""CPU ORACLE (test infrastructure): Eulerian <-> Lagrangian grid communicator and virtual boundary
forcing, numpy restatement.

Follows sopht/numeric/immersed_boundary_ops/EulerianLagrangianGridCommunicator3D.py (and ...2D.py,
which is the same algorithm without z) and VirtualBoundaryForcing.py. Pinned against the reference's
own numba implementation, which does run in the build container: tests/golden/make_golden.py executes
it and stores inputs + outputs in tests/golden/ib_*.npz.

Shapes (interp_kernel_width = 2 -> 4 taps per axis), N = number of Lagrangian nodes:
  lag_positions (dim, N); nearest index (dim, N) int64, row 0 = x index;
  local support (dim, [4,] 4, 4, N) with stencil axes ordered (z, y, x);
  weights ([4,] 4, 4, N).
"""

from __future__ import annotations

import numpy as np


def _offsets(dim: int, w: int) -> np.ndarray:
    """(dim, [2w,] 2w, 2w, 1) integer tap offsets; component d varies along stencil axis (dim-1-d).
    EulerianLagrangianGridCommunicator3D.py:83-100, ...2D.py:84-99."""
    x = np.arange(-w + 1, w + 1)
    grids = np.meshgrid(*([x] * dim), indexing="ij")  # ordered (z, y, x)
    return np.stack(grids[::-1]).reshape((dim,) + (2 * w,) * dim + (1,))


def local_eulerian_grid_support_of_lagrangian_grid(
    local_support, nearest_idx, lag_positions, dx, eul_grid_coord_shift, interp_kernel_width=2
):
    """EulerianLagrangianGridCommunicator3D.py:102-135. Index = floor((X - shift)/dx); support =
    (idx + offset)*dx + shift - X, evaluated in float64 (numba promotes int64*float32) and stored
    in the support array's dtype."""
    dim, n = lag_positions.shape
    nearest_idx[...] = np.floor_divide(lag_positions - eul_grid_coord_shift, dx)
    offs = _offsets(dim, interp_kernel_width)
    bshape = (dim,) + (1,) * dim + (n,)
    local_support[...] = (
        (nearest_idx.reshape(bshape) + offs) * np.float64(dx)
        + np.float64(eul_grid_coord_shift)
        - lag_positions.reshape(bshape).astype(np.float64)
    )


def cosine_interpolation_weights(interp_weights, local_support, dx):
    """...3D.py:387-412: mutates local_support (/= dx); w = (0.25/dx)^dim * prod(1 + cos(pi/2 r))."""
    t = interp_weights.dtype.type
    dim = local_support.shape[0]
    local_support /= t(dx)
    w = t((0.25 / dx) ** dim)
    for d in range(dim):
        w = w * (t(1.0) + np.cos(t(0.5 * np.pi) * local_support[d]))
    interp_weights[...] = w


def peskin_interpolation_weights(interp_weights, local_support, dx):
    """...3D.py:415-518 (Peskin 2002, eq. 6.27): mutates local_support (= |.|/dx); evaluated in
    float64 then cast, as the reference's trailing .astype(real_t) indicates."""
    dim = local_support.shape[0]
    local_support[...] = np.fabs(local_support) / dx
    w = np.float64((0.125 / dx) ** dim)
    for d in range(dim):
        r = local_support[d].astype(np.float64)
        phi = (r < 1.0) * (3.0 - 2 * r + np.sqrt(np.fabs(1 + 4 * r - 4 * r**2))) + (r >= 1.0) * (
            r < 2.0
        ) * (5.0 - 2 * r - np.sqrt(np.fabs(-7 + 12 * r - 4 * r**2)))
        w = w * phi
    interp_weights[...] = w.astype(interp_weights.dtype)


def _taps(eul_field, nearest_idx, i, w):
    dim = nearest_idx.shape[0]
    sl = tuple(
        slice(nearest_idx[dim - 1 - ax, i] - w + 1, nearest_idx[dim - 1 - ax, i] + w + 1)
        for ax in range(dim)
    )
    return eul_field[(Ellipsis,) + sl]


def eulerian_to_lagrangian_grid_interpolation(
    lag_field, eul_field, interp_weights, nearest_idx, dx, interp_kernel_width=2
):
    """...3D.py:198-295: q_i = dx^dim * sum(eul[taps] * w[..., i]); scalar (N,) or vector (dim, N)."""
    dim, n = nearest_idx.shape
    t = eul_field.dtype.type
    vol = t(dx) ** dim
    vector = eul_field.ndim == dim + 1
    for i in range(n):
        taps = _taps(eul_field, nearest_idx, i, interp_kernel_width)
        if vector:
            for c in range(eul_field.shape[0]):
                lag_field[c, i] = np.sum(taps[c] * interp_weights[..., i]) * vol
        else:
            lag_field[i] = np.sum(taps * interp_weights[..., i]) * vol


def lagrangian_to_eulerian_grid_interpolation(
    eul_field, lag_field, interp_weights, nearest_idx, interp_kernel_width=2
):
    """...3D.py:318-380: eul[taps] += F_i * w[..., i] (accumulates; no dx^dim)."""
    dim, n = nearest_idx.shape
    vector = eul_field.ndim == dim + 1
    for i in range(n):
        taps = _taps(eul_field, nearest_idx, i, interp_kernel_width)
        if vector:
            taps += lag_field[:, i].reshape((-1,) + (1,) * dim) * interp_weights[..., i]
        else:
            taps += lag_field[i] * interp_weights[..., i]


class VirtualBoundaryForcing:
    """VirtualBoundaryForcing.py:20-283 (Goldstein 1993 penalty forcing), cosine kernel, width 2."""

    def __init__(self, stiffness, damping, grid_dim, dx, num_lag_nodes, real_t, eul_grid_coord_shift=None,
                 interp_kernel_width=2, start_time=0.0, interp_kernel_type="cosine"):
        self.k, self.c = stiffness, damping
        self.dim, self.dx, self.n, self.real_t = grid_dim, dx, num_lag_nodes, real_t
        self.shift = real_t(dx / 2) if eul_grid_coord_shift is None else eul_grid_coord_shift
        self.w = interp_kernel_width
        self.time = start_time
        self.kernel_type = interp_kernel_type
        self.nearest_eul_grid_index_to_lag_grid = np.empty((grid_dim, num_lag_nodes), dtype=int)
        self.local_eul_grid_support_of_lag_grid = np.empty(
            (grid_dim,) + (2 * self.w,) * grid_dim + (num_lag_nodes,), dtype=real_t
        )
        self.interp_weights = np.empty((2 * self.w,) * grid_dim + (num_lag_nodes,), dtype=real_t)
        self.lag_grid_flow_velocity_field = np.zeros((grid_dim, num_lag_nodes), dtype=real_t)
        self.lag_grid_position_mismatch_field = np.zeros_like(self.lag_grid_flow_velocity_field)
        self.lag_grid_velocity_mismatch_field = np.zeros_like(self.lag_grid_flow_velocity_field)
        self.lag_grid_forcing_field = np.zeros_like(self.lag_grid_flow_velocity_field)

    def compute_interaction_force_on_lag_grid(self, eul_grid_velocity_field, lag_grid_position_field, lag_grid_velocity_field):
        """VirtualBoundaryForcing.py:187-230."""
        local_eulerian_grid_support_of_lagrangian_grid(
            self.local_eul_grid_support_of_lag_grid, self.nearest_eul_grid_index_to_lag_grid,
            lag_grid_position_field, self.dx, self.shift, self.w)
        weights = cosine_interpolation_weights if self.kernel_type == "cosine" else peskin_interpolation_weights
        weights(self.interp_weights, self.local_eul_grid_support_of_lag_grid, self.dx)
        eulerian_to_lagrangian_grid_interpolation(
            self.lag_grid_flow_velocity_field, eul_grid_velocity_field, self.interp_weights,
            self.nearest_eul_grid_index_to_lag_grid, self.dx, self.w)
        self.lag_grid_velocity_mismatch_field[...] = self.lag_grid_flow_velocity_field - lag_grid_velocity_field
        self.lag_grid_forcing_field[...] = (
            self.k * self.lag_grid_position_mismatch_field + self.c * self.lag_grid_velocity_mismatch_field
        )

    def compute_interaction_force_on_eul_and_lag_grid(self, eul_grid_forcing_field, eul_grid_velocity_field,
                                                      lag_grid_position_field, lag_grid_velocity_field):
        """VirtualBoundaryForcing.py:232-253."""
        self.compute_interaction_force_on_lag_grid(eul_grid_velocity_field, lag_grid_position_field, lag_grid_velocity_field)
        lagrangian_to_eulerian_grid_interpolation(
            eul_grid_forcing_field, self.lag_grid_forcing_field, self.interp_weights,
            self.nearest_eul_grid_index_to_lag_grid, self.w)

    def time_step(self, dt):
        """VirtualBoundaryForcing.py:276-283."""
        self.lag_grid_position_mismatch_field[...] = (
            self.lag_grid_position_mismatch_field + dt * self.lag_grid_velocity_mismatch_field
        )
        self.time += dt
