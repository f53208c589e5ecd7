"""Immersed-boundary parity cases against goldens produced by the reference's own numba
implementation (EulerianLagrangianGridCommunicator{2,3}D, VirtualBoundaryForcing)."""

from __future__ import annotations

import numpy as np
from conftest import load_golden, test_tol


def _atol(precision, scale=1.0):
    return test_tol(precision) * scale


def case_communicator(comm_factory, precision):
    """comm_factory(dim, dx, shift, n_lag, real_t, n_components, kernel_type) -> object exposing the
    reference's four kernels (same attribute names)."""
    gl = load_golden("ib", precision)
    real_t = np.float32 if precision == "single" else np.float64
    for dim in (3, 2):
        for kernel_type in ("cosine", "peskin"):
            g = gl.case(f"comm{dim}d_{kernel_type}")
            n = g["lag_positions"].shape[1]
            dx, shift = g["dx"][()], g["shift"][()]
            comm = comm_factory(dim, dx, shift, n, real_t, dim, kernel_type)
            comm_s = comm_factory(dim, dx, shift, n, real_t, 1, kernel_type)
            idx = np.empty((dim, n), dtype=np.int64)
            support = np.empty((dim,) + (4,) * dim + (n,), dtype=real_t)
            weights = np.empty((4,) * dim + (n,), dtype=real_t)
            comm.local_eulerian_grid_support_of_lagrangian_grid_kernel(support, idx, g["lag_positions"])
            np.testing.assert_array_equal(idx, g["nearest_idx"])  # integer index work: bit exact
            np.testing.assert_allclose(support, g["local_support"], atol=_atol(precision))
            comm.interpolation_weights_kernel(weights, support)
            # weights scale like (1/dx)^dim: compare relative to their magnitude
            wscale = float(np.max(np.abs(g["interp_weights"])))
            np.testing.assert_allclose(weights, g["interp_weights"], atol=_atol(precision, wscale))
            np.testing.assert_allclose(support, g["local_support_after_weights"], atol=_atol(precision, 10))
            lag_vec = np.zeros((dim, n), dtype=real_t)
            lag_sca = np.zeros(n, dtype=real_t)
            comm.eulerian_to_lagrangian_grid_interpolation_kernel(lag_vec, g["eul_vector_field"], weights, idx)
            comm_s.eulerian_to_lagrangian_grid_interpolation_kernel(lag_sca, g["eul_scalar_field"], weights, idx)
            np.testing.assert_allclose(lag_vec, g["lag_vector_field"], atol=_atol(precision, 10))
            np.testing.assert_allclose(lag_sca, g["lag_scalar_field"], atol=_atol(precision, 10))
            spread_vec = np.zeros_like(g["spread_vector_field"])
            spread_sca = np.zeros_like(g["spread_scalar_field"])
            comm.lagrangian_to_eulerian_grid_interpolation_kernel(spread_vec, g["lag_force_vector"], weights, idx)
            comm_s.lagrangian_to_eulerian_grid_interpolation_kernel(spread_sca, g["lag_force_scalar"], weights, idx)
            sscale = float(np.max(np.abs(g["spread_vector_field"])))
            np.testing.assert_allclose(spread_vec, g["spread_vector_field"], atol=_atol(precision, sscale))
            np.testing.assert_allclose(spread_sca, g["spread_scalar_field"], atol=_atol(precision, sscale))
            # conservation of the spread quantity (test_eulerian_lagrangian_grid_communicator_3d.py:268-300)
            np.testing.assert_allclose(
                spread_sca.sum(dtype=np.float64) * float(dx) ** dim,
                g["lag_force_scalar"].sum(dtype=np.float64), atol=_atol(precision, 100))


def case_virtual_boundary(vb_factory, precision):
    """vb_factory(dim, stiffness, damping, dx, n_lag, real_t) -> VirtualBoundaryForcing-like object."""
    gl = load_golden("ib", precision)
    real_t = np.float32 if precision == "single" else np.float64
    for dim in (3, 2):
        g = gl.case(f"virtual_boundary_{dim}d")
        n = g["lag_positions"].shape[1]
        vb = vb_factory(dim, g["stiffness"][()], g["damping"][()], g["dx"][()], n, real_t)
        eul_force = np.zeros_like(g["eul_velocity_field"])
        for step in range(3):
            vb.time_step(g["dt"][()])
            vb.compute_interaction_force_on_eul_and_lag_grid(
                eul_force, g["eul_velocity_field"], g["lag_positions"], g["lag_body_velocity"])
            fscale = float(np.max(np.abs(g[f"lag_forcing_step{step}"])))
            np.testing.assert_allclose(
                np.asarray(vb.lag_grid_forcing_field), g[f"lag_forcing_step{step}"], atol=_atol(precision, fscale))
            np.testing.assert_allclose(
                np.asarray(vb.lag_grid_position_mismatch_field), g[f"position_mismatch_step{step}"],
                atol=_atol(precision))
            escale = float(np.max(np.abs(g[f"eul_forcing_step{step}"])))
            np.testing.assert_allclose(eul_force, g[f"eul_forcing_step{step}"], atol=_atol(precision, escale))
