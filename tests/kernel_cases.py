"""Parity cases shared by the CPU (oracle vs golden) and GPU (CUDA vs golden + oracle) suites.

Each case mirrors one test of the reference's tests/test_numeric/test_eulerian_grid_ops (same seeded
inputs through tests/golden, same pre-filled outputs, same comparison region and atol = 1e3*eps).
``ops`` is an adapter from tests/adapters.py.
"""

from __future__ import annotations

import numpy as np
from conftest import load_golden, rel_l2, test_tol


def _check(actual, expected, precision, region=None):
    if region is not None:
        actual, expected = actual[region], expected[region]
    np.testing.assert_allclose(expected, actual, atol=test_tol(precision))


# ---------------------------------------------------------------------------------------------------
# 3-D stencils (golden group "stencils3d")
# ---------------------------------------------------------------------------------------------------
def case_diffusion_flux_3d(ops, precision):
    g = load_golden("stencils3d", precision).case("diffusion_flux")
    # reference: test_diffusion_flux_3d.py:75-98 (output pre-filled with ones, reset_ghost_zone=True)
    flux = np.ones_like(g["ref_field"])
    ops.diffusion_flux(flux, g["ref_field"], g["prefactor"])
    _check(flux, g["ref_diffusion_flux"], precision)
    vflux = np.ones_like(g["ref_vector_field"])
    ops.diffusion_flux(vflux, g["ref_vector_field"], g["prefactor"], vector=True)
    _check(vflux, g["ref_vector_field_diffusion_flux"], precision)


def case_diffusion_flux_3d_no_reset(ops, precision):
    g = load_golden("stencils3d", precision).case("diffusion_flux")
    flux = np.full_like(g["ref_field"], 7)
    ops.diffusion_flux(flux, g["ref_field"], g["prefactor"], reset=False)
    inner = (slice(1, -1),) * 3
    _check(flux, g["ref_diffusion_flux"], precision, inner)
    ring = np.ones(flux.shape, bool)
    ring[inner] = False
    assert np.all(flux[ring] == 7)  # untouched cells keep their old value


def case_diffusion_timestep_3d(ops, precision):
    g = load_golden("stencils3d", precision).case("diffusion_timestep")
    f = g["ref_field"].copy()
    ops.diffusion_timestep(f, np.ones_like(f), g["nu_dt_by_dx2"])
    _check(f, g["ref_new_field"], precision)
    v = g["ref_vector_field"].copy()
    ops.diffusion_timestep(v, np.ones_like(v[0]), g["nu_dt_by_dx2"], vector=True)
    _check(v, g["ref_new_vector_field"], precision)


def case_curl_3d(ops, precision):
    g = load_golden("stencils3d", precision).case("curl")
    curl = np.ones_like(g["ref_field"])
    ops.curl_3d(curl, g["ref_field"], g["prefactor"])
    _check(curl, g["ref_curl"], precision)


def case_divergence_3d(ops, precision):
    g = load_golden("stencils3d", precision).case("divergence")
    div = np.ones_like(g["ref_divergence"])
    ops.divergence_3d(div, g["ref_field"], g["inv_dx"])
    _check(div, g["ref_divergence"], precision)


def case_forcing_update_3d(ops, precision):
    g = load_golden("stencils3d", precision).case("forcing_update")
    w = g["ref_vorticity_field"].copy()
    ops.forcing_update(w, g["ref_velocity_forcing_field"], g["prefactor"])
    inner = (slice(None),) + (slice(1, -1),) * 3  # reference compares the interior only (:68-73)
    _check(w, g["ref_new_vorticity_field"], precision, inner)
    ring = np.ones(w.shape, bool)
    ring[inner] = False
    assert np.array_equal(w[ring], g["ref_vorticity_field"][ring])


def case_penalised_velocity_update_3d(ops, precision):
    # reference: test_update_vorticity_from_velocity_forcing_3d.py:97-127 (curl of the difference)
    g = load_golden("stencils3d", precision).case("forcing_update")
    rng = np.random.default_rng(7)
    real_t = g["ref_vorticity_field"].dtype
    w0 = g["ref_vorticity_field"]
    u = rng.random(w0.shape).astype(real_t)
    pu = rng.random(w0.shape).astype(real_t)
    expect = w0.copy()
    from oracle import stencils as ost

    ost.update_vorticity_from_velocity_forcing_3d(expect, pu - u, g["prefactor"])
    w = w0.copy()
    ops.penalised_velocity_update(w, pu, u, g["prefactor"])
    _check(w, expect, precision)


def case_stretching_flux_3d(ops, precision):
    g = load_golden("stencils3d", precision).case("stretching_flux")
    q = np.ones_like(g["ref_vorticity_field"])
    ops.stretching_flux(q, g["ref_vorticity_field"], g["ref_velocity_field"], g["prefactor"])
    _check(q, g["ref_vorticity_stretching_flux_field"], precision)


def case_stretching_timestep_3d(ops, precision):
    for stepper, case in (("euler_forward", "stretching_timestep_euler"), ("ssprk3", "stretching_timestep_ssprk3")):
        g = load_golden("stencils3d", precision).case(case)
        w = g["ref_vorticity_field"].copy()
        q = np.ones_like(w)
        mid = np.ones_like(w)
        ops.stretching_timestep(w, g["ref_velocity_field"], q, g["dt_by_2_dx"], stepper, mid)
        _check(w, g["ref_new_vorticity_field"], precision)


def case_advection_flux_3d(ops, precision):
    g = load_golden("stencils3d", precision).case("advection_flux")
    q = np.zeros_like(g["ref_field"])
    ops.advection_flux(q, g["ref_field"], g["ref_velocity"], g["inv_dx"])
    _check(q, g["ref_advection_flux"], precision)
    # accumulation semantics: a second call adds the same flux again
    ops.advection_flux(q, g["ref_field"], g["ref_velocity"], g["inv_dx"])
    _check(q, 2 * g["ref_advection_flux"], precision)


def case_advection_timestep_3d(ops, precision):
    g = load_golden("stencils3d", precision).case("advection_timestep")
    inner = (slice(2, -2),) * 3
    dt_by_dx = g["ref_field"].dtype.type(g["dt"] * g["inv_dx"])
    f = g["ref_field"].copy()
    ops.advection_timestep(f, np.ones_like(f), g["ref_velocity"], dt_by_dx)
    _check(f, g["ref_new_field"], precision, inner)
    v = g["ref_vector_field"].copy()
    ops.advection_timestep(v, np.ones_like(v[0]), g["ref_velocity"], dt_by_dx, vector=True)
    _check(v, g["ref_new_vector_field"], precision, (slice(None),) + inner)


def case_penalise_3d(ops, precision):
    g = load_golden("stencils3d", precision).case("penalise")
    grids = (g["x_grid_field"], g["y_grid_field"], g["z_grid_field"])
    f = g["ref_field"].copy()
    ops.penalise(f, int(g["width"]), g["dx"][()], grids)
    _check(f, g["ref_penalised_field"], precision)
    v = g["ref_vector_field"].copy()
    ops.penalise(v, int(g["width"]), g["dx"][()], grids, vector=True)
    _check(v, g["ref_penalised_vector_field"], precision)


def case_brinkmann_3d(ops, precision):
    g = load_golden("stencils3d", precision).case("brinkmann_penalise")
    o = np.ones_like(g["ref_field"])
    ops.brinkmann(o, g["ref_field"], g["ref_char_field"], g["ref_penalty_field"], g["penalty_factor"])
    _check(o, g["ref_penalised_field"], precision)
    vo = np.ones_like(g["ref_vector_field"])
    ops.brinkmann(vo, g["ref_vector_field"], g["ref_char_field"], g["ref_penalty_vector_field"],
                  g["penalty_factor"], vector=True)
    _check(vo, g["ref_penalised_vector_field"], precision)


def case_char_func_3d(ops, precision):
    g = load_golden("stencils3d", precision).case("char_func")
    o = np.ones_like(g["level_set_field"])
    ops.char_func(o, g["level_set_field"], g["blend_width"][()])
    _check(o, g["ref_char_func_field"], precision)


def case_laplacian_filter_3d(ops, precision):
    gl = load_golden("stencils3d", precision)
    for ftype in ("convolution", "multiplicative"):
        for order in (1, 2):
            g = gl.case(f"laplacian_filter_{ftype}_{order}")
            f = g["field"].copy()
            ops.laplacian_filter(f, np.zeros_like(f), np.zeros_like(f), order, ftype)
            _check(f, g["ref_field"], precision)
            v = g["vector_field"].copy()
            ops.laplacian_filter(v, np.zeros_like(v[0]), np.zeros_like(v[0]), order, ftype, vector=True)
            _check(v, g["ref_vector_field"], precision)
    # constant field is a fixed point (test_laplacian_filter_3d.py:102-131)
    c = 2 * np.ones((16, 16, 16), dtype=g["field"].dtype)
    ops.laplacian_filter(c, np.zeros_like(c), np.zeros_like(c), 2, "multiplicative")
    _check(c, 2 * np.ones_like(c), precision)


# ---------------------------------------------------------------------------------------------------
# 2-D stencils (golden group "stencils2d")
# ---------------------------------------------------------------------------------------------------
def case_diffusion_flux_2d(ops, precision):
    g = load_golden("stencils2d", precision).case("diffusion_flux")
    flux = np.ones_like(g["ref_field"])
    ops.diffusion_flux(flux, g["ref_field"], g["prefactor"])
    _check(flux, g["ref_diffusion_flux"], precision)


def case_diffusion_timestep_2d(ops, precision):
    g = load_golden("stencils2d", precision).case("diffusion_timestep")
    f = g["ref_field"].copy()
    ops.diffusion_timestep(f, np.ones_like(f), g["nu_dt_by_dx2"])
    _check(f, g["ref_new_field"], precision)


def case_advection_flux_2d(ops, precision):
    g = load_golden("stencils2d", precision).case("advection_flux")
    q = np.zeros_like(g["ref_field"])
    ops.advection_flux(q, g["ref_field"], g["ref_velocity"], g["inv_dx"])
    _check(q, g["ref_advection_flux"], precision)


def case_advection_timestep_2d(ops, precision):
    g = load_golden("stencils2d", precision).case("advection_timestep")
    f = g["ref_field"].copy()
    dt_by_dx = f.dtype.type(g["dt"] * g["inv_dx"])
    ops.advection_timestep(f, np.ones_like(f), g["ref_velocity"], dt_by_dx)
    _check(f, g["ref_new_field"], precision, (slice(2, -2),) * 2)


def case_outplane_curl_2d(ops, precision):
    g = load_golden("stencils2d", precision).case("outplane_curl")
    curl = np.ones_like(g["ref_curl"])
    ops.outplane_curl_2d(curl, g["ref_field"], g["prefactor"])
    _check(curl, g["ref_curl"], precision)


def case_inplane_curl_2d(ops, precision):
    g = load_golden("stencils2d", precision).case("inplane_curl")
    curl = np.zeros_like(g["ref_curl"])
    ops.inplane_curl_2d(curl, g["ref_field"], g["prefactor"])
    _check(curl, g["ref_curl"], precision, (slice(1, -1),) * 2)


def case_forcing_update_2d(ops, precision):
    g = load_golden("stencils2d", precision).case("forcing_update")
    w = g["ref_vorticity_field"].copy()
    ops.forcing_update(w, g["ref_velocty_forcing_field"], g["prefactor"])
    _check(w, g["ref_new_vorticity_field"], precision, (slice(1, -1),) * 2)


def case_penalised_velocity_update_2d(ops, precision):
    g = load_golden("stencils2d", precision).case("forcing_update")
    rng = np.random.default_rng(11)
    w0 = g["ref_vorticity_field"]
    shape = g["ref_velocty_forcing_field"].shape
    u = rng.random(shape).astype(w0.dtype)
    pu = rng.random(shape).astype(w0.dtype)
    from oracle import stencils as ost

    expect = w0.copy()
    ost.update_vorticity_from_velocity_forcing_2d(expect, pu - u, g["prefactor"])
    w = w0.copy()
    ops.penalised_velocity_update(w, pu, u, g["prefactor"])
    _check(w, expect, precision)


def case_penalise_2d(ops, precision):
    g = load_golden("stencils2d", precision).case("penalise")
    f = g["ref_field"].copy()
    ops.penalise(f, int(g["width"]), g["dx"][()], (g["x_grid_field"], g["y_grid_field"]))
    _check(f, g["ref_penalised_field"], precision)


def case_brinkmann_2d(ops, precision):
    g = load_golden("stencils2d", precision).case("brinkmann_penalise")
    o = np.ones_like(g["ref_field"])
    ops.brinkmann(o, g["ref_field"], g["ref_char_field"], g["ref_penalty_field"], g["penalty_factor"])
    _check(o, g["ref_penalised_field"], precision)
    vo = np.ones_like(g["ref_vector_field"])
    ops.brinkmann(vo, g["ref_vector_field"], g["ref_char_field"], g["ref_penalty_vector_field"],
                  g["penalty_factor"], vector=True)
    _check(vo, g["ref_penalised_vector_field"], precision)
    # vs fixed value == vs a constant penalty field (test_brinkmann_penalise_2d.py)
    const = np.full_like(g["ref_field"], 0.3)
    expect = np.ones_like(const)
    from oracle import stencils as ost

    ost.brinkmann_penalise(expect, g["ref_field"], g["ref_char_field"], const, g["penalty_factor"])
    o2 = np.ones_like(const)
    ops.brinkmann_vs_fixed_val(o2, g["ref_field"], g["ref_char_field"], g["penalty_factor"], 0.3)
    _check(o2, expect, precision)


def case_char_func_2d(ops, precision):
    g = load_golden("stencils2d", precision).case("char_func")
    o = np.ones_like(g["level_set_field"])
    ops.char_func(o, g["level_set_field"], g["blend_width"][()])
    _check(o, g["ref_char_func_field"], precision)


# ---------------------------------------------------------------------------------------------------
# elementwise known answers (test_elementwise_ops_{2d,3d}.py)
# ---------------------------------------------------------------------------------------------------
def case_elementwise(ops, precision):
    t = np.float32 if precision == "single" else np.float64
    for shape in ((16, 16, 16), (16, 16), (3, 16, 16, 16)):
        a = 2 * np.ones(shape, dtype=t)
        b = 3 * np.ones(shape, dtype=t)
        s = np.zeros(shape, dtype=t)
        ops.elementwise_sum(s, a, b)
        _check(s, 5 * np.ones(shape, dtype=t), precision)
        ops.saxpby(s, a, b, 2.0, 3.0)
        _check(s, 13 * np.ones(shape, dtype=t), precision)
    for shape in ((16, 16, 16), (16, 16)):
        f = np.ones(shape, dtype=t)
        ops.set_fixed_val(f, 3)
        _check(f, 3 * np.ones(shape, dtype=t), precision)
        ops.add_fixed_val(f, f, 2)  # aliasing in == out
        _check(f, 5 * np.ones(shape, dtype=t), precision)
        c = np.zeros(shape, dtype=t)
        ops.elementwise_copy(c, f)
        _check(c, f, precision)
        v = np.ones((len(shape), *shape), dtype=t)
        vals = [2.0, 3.0, 4.0][: len(shape)]
        ops.set_fixed_val_vector(v, vals)
        for k, val in enumerate(vals):
            _check(v[k], val * np.ones(shape, dtype=t), precision)
        ops.add_fixed_val_vector(v, v, [1.0] * len(shape))
        for k, val in enumerate(vals):
            _check(v[k], (val + 1) * np.ones(shape, dtype=t), precision)
        # boundary setter (test_elementwise_ops_3d.py:137-183)
        for width in (1, 2):
            f = np.ones(shape, dtype=t)
            ops.set_fixed_val_at_boundaries(f, width, 3)
            e = 3 * np.ones(shape, dtype=t)
            e[(slice(width, -width),) * len(shape)] = 1
            _check(f, e, precision)
            v = np.ones((len(shape), *shape), dtype=t)
            ops.set_fixed_val_at_boundaries_vector(v, width, vals)
            for k, val in enumerate(vals):
                e = val * np.ones(shape, dtype=t)
                e[(slice(width, -width),) * len(shape)] = 1
                _check(v[k], e, precision)
        # complex product (1+2j)(2+3j) = -4+7j (test_elementwise_ops_3d.py:118-132)
        ct = np.complex64 if precision == "single" else np.complex128
        x = (1 + 2j) * np.ones(shape, dtype=ct)
        y = (2 + 3j) * np.ones(shape, dtype=ct)
        z = np.zeros(shape, dtype=ct)
        ops.complex_product(z, x, y)
        _check(z, (-4 + 7j) * np.ones(shape, dtype=ct), precision)
    # cross product (1,2,3)x(4,5,6) = (-3,6,-3) (test_elementwise_ops_3d.py:284-309)
    a = np.ones((3, 16, 16, 16), dtype=t) * np.array([1, 2, 3], dtype=t).reshape(3, 1, 1, 1)
    b = np.ones((3, 16, 16, 16), dtype=t) * np.array([4, 5, 6], dtype=t).reshape(3, 1, 1, 1)
    r = np.zeros_like(a)
    ops.cross_product(r, a, b)
    _check(r, np.ones_like(a) * np.array([-3, 6, -3], dtype=t).reshape(3, 1, 1, 1), precision)


# ---------------------------------------------------------------------------------------------------
# Poisson (golden group "poisson")
# ---------------------------------------------------------------------------------------------------
def case_poisson_3d(ops, precision, **solver_kw):
    gl = load_golden("poisson", precision)
    for name in ("poisson3d_16", "poisson3d_8x12x20"):
        g = gl.case(name)
        grid = (int(g["grid_size_z"]), int(g["grid_size_y"]), int(g["grid_size_x"]))
        solver = ops.poisson_solver(grid, g["x_range"][()], **solver_kw)
        sol = np.zeros_like(g["rhs_field"])
        solver.solve(solution_field=sol, rhs_field=g["rhs_field"])
        _check(sol, g["ref_solution_field"], precision)
        vsol = np.zeros_like(g["rhs_vector_field"])
        solver.vector_field_solve(solution_vector_field=vsol, rhs_vector_field=g["rhs_vector_field"])
        _check(vsol, g["ref_solution_vector_field"], precision)
        assert rel_l2(vsol, g["ref_solution_vector_field"]) < (1e-5 if precision == "single" else 1e-12)


def case_poisson_2d(ops, precision, **solver_kw):
    g = load_golden("poisson", precision).case("poisson2d_16")
    grid = (int(g["grid_size_y"]), int(g["grid_size_x"]))
    solver = ops.poisson_solver(grid, g["x_range"][()], **solver_kw)
    sol = np.zeros_like(g["rhs_field"])
    solver.solve(solution_field=sol, rhs_field=g["rhs_field"])
    _check(sol, g["ref_solution_field"], precision)


STENCIL_CASES = [
    case_diffusion_flux_3d,
    case_diffusion_flux_3d_no_reset,
    case_diffusion_timestep_3d,
    case_curl_3d,
    case_divergence_3d,
    case_forcing_update_3d,
    case_penalised_velocity_update_3d,
    case_stretching_flux_3d,
    case_stretching_timestep_3d,
    case_advection_flux_3d,
    case_advection_timestep_3d,
    case_penalise_3d,
    case_brinkmann_3d,
    case_char_func_3d,
    case_laplacian_filter_3d,
    case_diffusion_flux_2d,
    case_diffusion_timestep_2d,
    case_advection_flux_2d,
    case_advection_timestep_2d,
    case_outplane_curl_2d,
    case_inplane_curl_2d,
    case_forcing_update_2d,
    case_penalised_velocity_update_2d,
    case_penalise_2d,
    case_brinkmann_2d,
    case_char_func_2d,
    case_elementwise,
]
POISSON_CASES = [case_poisson_3d, case_poisson_2d]
