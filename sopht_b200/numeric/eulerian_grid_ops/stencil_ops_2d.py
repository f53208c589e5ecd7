"""2-D stencil kernel factories.

Drop-in counterparts of sopht/numeric/eulerian_grid_ops/stencil_ops_2d/*.py (same names, keyword
names, defaults, ValueErrors; in-place outputs). See stencil_ops_3d.py for conventions.
"""

from __future__ import annotations

from collections.abc import Callable
from typing import Any, Literal

from sopht_b200 import _lib

from .elementwise_ops import _check_field_type
from .stencil_ops_3d import _sine_ramps, _to_numpy


def gen_diffusion_flux_pyst_kernel_2d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int] | bool = False,
    reset_ghost_zone: bool = True,
) -> Callable:
    """2D diffusion flux kernel generator (diffusion_flux_2d.py:13-72)."""
    dt = _lib.dtype_code(real_t)
    reset = 0 if reset_ghost_zone is False else 1

    def diffusion_flux_pyst_kernel_2d(diffusion_flux: Any, field: Any, prefactor: float) -> None:
        """diffusion_flux = prefactor * 5-point Laplacian(field) on the ring-1 interior."""
        with _lib.Staging() as s:
            f, o = s.inp(field), s.out(diffusion_flux)
            _lib.call("sopht_diffusion_flux_2d", dt, o, f, prefactor, reset)

    return diffusion_flux_pyst_kernel_2d


def gen_diffusion_timestep_euler_forward_pyst_kernel_2d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int] | bool = False,
) -> Callable:
    """2D diffusion Euler-forward timestep generator (diffusion_timestep_2d.py:10-43)."""
    dt = _lib.dtype_code(real_t)

    def diffusion_timestep_euler_forward_pyst_kernel_2d(
        field: Any, diffusion_flux: Any, nu_dt_by_dx2: float
    ) -> None:
        with _lib.Staging() as s:
            f, q = s.out(field), s.out(diffusion_flux)
            _lib.call("sopht_diffusion_flux_2d", dt, q, f, nu_dt_by_dx2, 1)
            _lib.call("sopht_elementwise_sum", dt, f, f, q)

    return diffusion_timestep_euler_forward_pyst_kernel_2d


def gen_advection_flux_conservative_eno3_pyst_kernel_2d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int] | bool = False,
) -> Callable:
    """2D conservative ENO3 advection flux generator (advection_flux_2d.py:12-165); accumulates."""
    dt = _lib.dtype_code(real_t)

    def advection_flux_conservative_eno3_pyst_kernel_2d(
        advection_flux: Any, field: Any, velocity: Any, inv_dx: float
    ) -> None:
        with _lib.Staging() as s:
            f, v = s.inp(field), s.inp(velocity)
            _lib.call("sopht_advection_flux_eno3_2d", dt, s.out(advection_flux), f, v, inv_dx)

    return advection_flux_conservative_eno3_pyst_kernel_2d


def gen_advection_timestep_euler_forward_conservative_eno3_pyst_kernel_2d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int] | bool = False,
) -> Callable:
    """2D ENO3 advection Euler-forward timestep generator (advection_timestep_2d.py:10-55)."""
    dt = _lib.dtype_code(real_t)

    def advection_timestep_euler_forward_conservative_eno3_pyst_kernel_2d(
        field: Any, advection_flux: Any, velocity: Any, dt_by_dx: float
    ) -> None:
        with _lib.Staging() as s:
            v = s.inp(velocity)
            f, q = s.out(field), s.out(advection_flux)
            _lib.call("sopht_set_fixed_val", dt, q, 0.0)
            _lib.call("sopht_advection_flux_eno3_2d", dt, q, f, v, -float(dt_by_dx))
            _lib.call("sopht_elementwise_sum", dt, f, f, q)

    return advection_timestep_euler_forward_conservative_eno3_pyst_kernel_2d


def gen_outplane_field_curl_pyst_kernel_2d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int] | bool = False,
    reset_ghost_zone: bool = True,
) -> Callable:
    """psi -> velocity curl generator (outplane_field_curl_2d.py:13-100)."""
    dt = _lib.dtype_code(real_t)
    reset = 1 if reset_ghost_zone else 0

    def outplane_field_curl_pyst_kernel_2d(curl: Any, field: Any, prefactor: float) -> None:
        with _lib.Staging() as s:
            f = s.inp(field)
            _lib.call("sopht_outplane_field_curl_2d", dt, s.out(curl), f, prefactor, reset)

    return outplane_field_curl_pyst_kernel_2d


def gen_inplane_field_curl_pyst_kernel_2d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int] | bool = False,
) -> Callable:
    """velocity -> vorticity curl generator, no ring reset (inplane_field_curl_2d.py:10-50)."""
    dt = _lib.dtype_code(real_t)

    def inplane_field_curl_pyst_kernel_2d(curl: Any, field: Any, prefactor: float) -> None:
        with _lib.Staging() as s:
            f = s.inp(field)
            _lib.call("sopht_inplane_field_curl_2d", dt, s.out(curl), f, prefactor)

    return inplane_field_curl_pyst_kernel_2d


def gen_update_vorticity_from_velocity_forcing_pyst_kernel_2d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int] | bool = False,
) -> Callable:
    """vorticity += prefactor * curl(forcing) (update_vorticity_from_velocity_forcing_2d.py:12-71)."""
    dt = _lib.dtype_code(real_t)

    def update_vorticity_from_velocity_forcing_pyst_kernel_2d(
        vorticity_field: Any, velocity_forcing_field: Any, prefactor: float
    ) -> None:
        with _lib.Staging() as s:
            f = s.inp(velocity_forcing_field)
            _lib.call(
                "sopht_update_vorticity_from_velocity_forcing_2d",
                dt,
                s.out(vorticity_field),
                f,
                prefactor,
            )

    return update_vorticity_from_velocity_forcing_pyst_kernel_2d


def gen_update_vorticity_from_penalised_velocity_pyst_kernel_2d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int] | bool = False,
) -> Callable:
    """vorticity += prefactor * curl(penalised_velocity - velocity) (…_2d.py:74-147)."""
    dt = _lib.dtype_code(real_t)

    def update_vorticity_from_penalised_velocity_pyst_kernel_2d(
        vorticity_field: Any, penalised_velocity_field: Any, velocity_field: Any, prefactor: float
    ) -> None:
        with _lib.Staging() as s:
            p, u = s.inp(penalised_velocity_field), s.inp(velocity_field)
            _lib.call(
                "sopht_update_vorticity_from_penalised_velocity_2d",
                dt,
                s.out(vorticity_field),
                p,
                u,
                prefactor,
            )

    return update_vorticity_from_penalised_velocity_pyst_kernel_2d


def gen_penalise_field_boundary_pyst_kernel_2d(
    width: int,
    dx: float,
    x_grid_field: Any,
    y_grid_field: Any,
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int] | bool = False,
) -> Callable:
    """2D penalise field boundary kernel generator (penalise_field_boundary_2d.py:12-142)."""
    if not isinstance(width, int) or width < 0:
        msg = "Invalid width for boundary zone, must be a non-negative integer"
        raise ValueError(msg)
    if width == 0:

        def penalise_field_boundary_pyst_kernel_2d(field: Any) -> None:
            pass

        return penalise_field_boundary_pyst_kernel_2d

    dt = _lib.dtype_code(real_t)
    ramp_x = _sine_ramps(_to_numpy(x_grid_field[0, :]), width, dx, real_t)
    ramp_y = _sine_ramps(_to_numpy(y_grid_field[:, 0]), width, dx, real_t)

    def penalise_field_boundary_pyst_kernel_2d(field: Any) -> None:  # noqa: F811
        with _lib.Staging() as s:
            _lib.call("sopht_penalise_field_boundary_2d", dt, s.out(field), width, ramp_x, ramp_y)

    return penalise_field_boundary_pyst_kernel_2d


def gen_brinkmann_penalise_pyst_kernel_2d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int] | bool = False,
    field_type: Literal["scalar", "vector"] = "scalar",
) -> Callable:
    """Brinkmann penalisation 2D kernel generator (brinkmann_penalise_2d.py:13-76)."""
    dt = _lib.dtype_code(real_t)
    _check_field_type(field_type)
    if field_type == "scalar":

        def brinkmann_penalise_pyst_kernel_2d(
            penalised_field: Any, field: Any, char_field: Any, penalty_field: Any, penalty_factor: float
        ) -> None:
            with _lib.Staging() as s:
                f, chi, pen = s.inp(field), s.inp(char_field), s.inp(penalty_field)
                _lib.call(
                    "sopht_brinkmann_penalise", dt, s.out(penalised_field), f, chi, pen, penalty_factor
                )

        return brinkmann_penalise_pyst_kernel_2d

    def brinkmann_penalise_vector_field_pyst_kernel_2d(
        penalised_vector_field: Any,
        penalty_factor: float,
        char_field: Any,
        penalty_vector_field: Any,
        vector_field: Any,
    ) -> None:
        with _lib.Staging() as s:
            f, chi, pen = s.inp(vector_field), s.inp(char_field), s.inp(penalty_vector_field)
            o = s.out(penalised_vector_field)
            for c in range(2):
                _lib.call("sopht_brinkmann_penalise", dt, o[c], f[c], chi, pen[c], penalty_factor)

    return brinkmann_penalise_vector_field_pyst_kernel_2d


def gen_brinkmann_penalise_vs_fixed_val_pyst_kernel_2d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int] | bool = False,
    field_type: Literal["scalar", "vector"] = "scalar",
) -> Callable:
    """Brinkmann penalisation against a fixed value (brinkmann_penalise_2d.py:79-141)."""
    dt = _lib.dtype_code(real_t)
    _check_field_type(field_type)
    if field_type == "scalar":

        def brinkmann_penalise_vs_fixed_val_pyst_kernel_2d(
            penalised_field: Any, field: Any, char_field: Any, penalty_factor: float, penalty_val: float
        ) -> None:
            with _lib.Staging() as s:
                f, chi = s.inp(field), s.inp(char_field)
                _lib.call(
                    "sopht_brinkmann_penalise_vs_fixed_val",
                    dt,
                    s.out(penalised_field),
                    f,
                    chi,
                    penalty_val,
                    penalty_factor,
                )

        return brinkmann_penalise_vs_fixed_val_pyst_kernel_2d

    def brinkmann_penalise_vector_field_vs_fixed_val_pyst_kernel_2d(
        penalised_vector_field: Any,
        penalty_factor: float,
        char_field: Any,
        penalty_val: Any,
        vector_field: Any,
    ) -> None:
        with _lib.Staging() as s:
            f, chi = s.inp(vector_field), s.inp(char_field)
            o = s.out(penalised_vector_field)
            for c in range(2):
                _lib.call(
                    "sopht_brinkmann_penalise_vs_fixed_val",
                    dt,
                    o[c],
                    f[c],
                    chi,
                    penalty_val[c],
                    penalty_factor,
                )

    return brinkmann_penalise_vector_field_vs_fixed_val_pyst_kernel_2d


def gen_char_func_from_level_set_via_sine_heaviside_pyst_kernel_2d(
    blend_width: float,
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int] | bool = False,
) -> Callable:
    """Smooth sine-Heaviside characteristic function (char_func_from_level_set_2d.py:12-51)."""
    dt = _lib.dtype_code(real_t)

    def char_func_from_level_set_via_sine_heaviside_pyst_kernel_2d(
        char_func_field: Any, level_set_field: Any
    ) -> None:
        with _lib.Staging() as s:
            ls = s.inp(level_set_field)
            _lib.call("sopht_char_func_from_level_set", dt, s.out(char_func_field), ls, blend_width)

    return char_func_from_level_set_via_sine_heaviside_pyst_kernel_2d
