"""3-D stencil kernel factories.

Drop-in counterparts of the factories in sopht/numeric/eulerian_grid_ops/stencil_ops_3d/*.py — same
names, keyword names, defaults and ValueErrors; the returned callables mutate their output argument
in place. Where the reference composes several generated kernels in a Python closure (flux + ring
reset, flux + sum, ...) the same composition order is kept, but each stage is one CUDA launch with
the ring handling folded in.
"""

from __future__ import annotations

from collections.abc import Callable
from typing import Any, Literal

import numpy as np
import torch

from sopht_b200 import _lib

from .elementwise_ops import _check_field_type


def _to_numpy(a: Any) -> np.ndarray:
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy()
    return np.asarray(a)


# ---- diffusion -------------------------------------------------------------------------------------------
def gen_diffusion_flux_pyst_kernel_3d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int, int] | bool = False,
    field_type: Literal["scalar", "vector"] = "scalar",
    reset_ghost_zone: bool = True,
) -> Callable:
    """3D diffusion flux kernel generator (diffusion_flux_3d.py:14-115)."""
    dt = _lib.dtype_code(real_t)
    _check_field_type(field_type)
    reset = 0 if reset_ghost_zone is False else 1

    if field_type == "scalar":

        def diffusion_flux_pyst_kernel_3d(diffusion_flux: Any, field: Any, prefactor: float) -> None:
            """diffusion_flux = prefactor * 7-point Laplacian(field) on the ring-1 interior."""
            with _lib.Staging() as s:
                f, o = s.inp(field), s.out(diffusion_flux)
                _lib.call("sopht_diffusion_flux_3d", dt, o, f, prefactor, reset)

        return diffusion_flux_pyst_kernel_3d

    def vector_field_diffusion_flux_pyst_kernel_3d(
        vector_field_diffusion_flux: Any, vector_field: Any, prefactor: float
    ) -> None:
        """Vector Laplacian flux of a (3, nz, ny, nx) field."""
        with _lib.Staging() as s:
            f, o = s.inp(vector_field), s.out(vector_field_diffusion_flux)
            _lib.call("sopht_diffusion_flux_3d", dt, o, f, prefactor, reset)

    return vector_field_diffusion_flux_pyst_kernel_3d


def gen_diffusion_timestep_euler_forward_pyst_kernel_3d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int, int] | bool = False,
    field_type: Literal["scalar", "vector"] = "scalar",
) -> Callable | None:
    """3D diffusion Euler-forward timestep generator (diffusion_timestep_3d.py:12-80).

    Like the reference, an unknown ``field_type`` silently yields ``None`` (no ``case _`` there).
    """
    dt = _lib.dtype_code(real_t)

    def diffusion_timestep_euler_forward_pyst_kernel_3d(
        field: Any, diffusion_flux: Any, nu_dt_by_dx2: float
    ) -> None:
        """field += nu_dt_by_dx2 * Laplacian(field); diffusion_flux holds the flux afterwards."""
        with _lib.Staging() as s:
            f, q = s.out(field), s.out(diffusion_flux)
            _lib.call("sopht_diffusion_flux_3d", dt, q, f, nu_dt_by_dx2, 1)
            _lib.call("sopht_elementwise_sum", dt, f, f, q)

    if field_type == "scalar":
        return diffusion_timestep_euler_forward_pyst_kernel_3d
    if field_type == "vector":

        def vector_field_diffusion_timestep_euler_forward_pyst_kernel_3d(
            vector_field: Any, diffusion_flux: Any, nu_dt_by_dx2: float
        ) -> None:
            """Component-by-component diffusion step through ONE scalar flux buffer."""
            with _lib.Staging() as s:
                v, q = s.out(vector_field), s.out(diffusion_flux)
                for c in range(3):
                    _lib.call("sopht_diffusion_flux_3d", dt, q, v[c], nu_dt_by_dx2, 1)
                    _lib.call("sopht_elementwise_sum", dt, v[c], v[c], q)

        return vector_field_diffusion_timestep_euler_forward_pyst_kernel_3d
    return None


# ---- curl / divergence -----------------------------------------------------------------------------------
def gen_curl_pyst_kernel_3d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int, int] | bool = False,
    reset_ghost_zone: bool = True,
) -> Callable:
    """3D curl kernel generator (curl_3d.py:13-132)."""
    dt = _lib.dtype_code(real_t)
    reset = 1 if reset_ghost_zone else 0

    def curl_pyst_kernel_3d(curl: Any, field: Any, prefactor: float) -> None:
        """curl = prefactor * centred-difference curl(field); ring-1 zeroed when reset_ghost_zone."""
        with _lib.Staging() as s:
            f, o = s.inp(field), s.out(curl)
            _lib.call("sopht_curl_3d", dt, o, f, prefactor, reset)

    return curl_pyst_kernel_3d


def gen_divergence_pyst_kernel_3d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int, int] | bool = False,
    reset_ghost_zone: bool = True,
) -> Callable:
    """3D divergence kernel generator (divergence_3d.py:13-96)."""
    dt = _lib.dtype_code(real_t)
    reset = 0 if reset_ghost_zone is False else 1

    def divergence_pyst_kernel_3d(divergence: Any, field: Any, inv_dx: float) -> None:
        """divergence = 0.5 * inv_dx * centred-difference div(field)."""
        with _lib.Staging() as s:
            f, o = s.inp(field), s.out(divergence)
            _lib.call("sopht_divergence_3d", dt, o, f, inv_dx, reset)

    return divergence_pyst_kernel_3d


# ---- vorticity updates -----------------------------------------------------------------------------------
def gen_update_vorticity_from_velocity_forcing_pyst_kernel_3d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int, int] | bool = False,
) -> Callable:
    """vorticity += prefactor * curl(velocity_forcing) (update_vorticity_from_velocity_forcing_3d.py:12-132)."""
    dt = _lib.dtype_code(real_t)

    def update_vorticity_from_velocity_forcing_pyst_kernel_3d(
        vorticity_field: Any, velocity_forcing_field: Any, prefactor: float
    ) -> None:
        with _lib.Staging() as s:
            f, w = s.inp(velocity_forcing_field), s.out(vorticity_field)
            _lib.call("sopht_update_vorticity_from_velocity_forcing_3d", dt, w, f, prefactor)

    return update_vorticity_from_velocity_forcing_pyst_kernel_3d


def gen_update_vorticity_from_penalised_velocity_pyst_kernel_3d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int, int] | bool = False,
) -> Callable:
    """vorticity += prefactor * curl(penalised_velocity - velocity) (…_3d.py:135-291)."""
    dt = _lib.dtype_code(real_t)

    def update_vorticity_from_penalised_velocity_pyst_kernel_3d(
        vorticity_field: Any, penalised_velocity_field: Any, velocity_field: Any, prefactor: float
    ) -> None:
        with _lib.Staging() as s:
            p, u = s.inp(penalised_velocity_field), s.inp(velocity_field)
            w = s.out(vorticity_field)
            _lib.call("sopht_update_vorticity_from_penalised_velocity_3d", dt, w, p, u, prefactor)

    return update_vorticity_from_penalised_velocity_pyst_kernel_3d


# ---- vorticity stretching --------------------------------------------------------------------------------
def gen_vorticity_stretching_flux_pyst_kernel_3d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int, int] | bool = False,
) -> Callable:
    """flux_c = prefactor * (omega . grad) u_c, ring zeroed (vorticity_stretching_flux_3d.py:13-107)."""
    dt = _lib.dtype_code(real_t)

    def vorticity_stretching_flux_pyst_kernel_3d(
        vorticity_stretching_flux_field: Any,
        vorticity_field: Any,
        velocity_field: Any,
        prefactor: float,
    ) -> None:
        with _lib.Staging() as s:
            w, u = s.inp(vorticity_field), s.inp(velocity_field)
            q = s.out(vorticity_stretching_flux_field)
            _lib.call("sopht_vorticity_stretching_flux_3d", dt, q, w, u, prefactor)

    return vorticity_stretching_flux_pyst_kernel_3d


def gen_vorticity_stretching_timestep_euler_forward_pyst_kernel_3d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int, int] | bool = False,
) -> Callable:
    """Euler-forward vortex stretching step (vorticity_stretching_timestep_3d.py:10-54)."""
    dt = _lib.dtype_code(real_t)

    def vorticity_stretching_timestep_euler_forward_pyst_kernel_3d(
        vorticity_field: Any,
        velocity_field: Any,
        vorticity_stretching_flux_field: Any,
        dt_by_2_dx: float,
    ) -> None:
        with _lib.Staging() as s:
            u = s.inp(velocity_field)
            w, q = s.out(vorticity_field), s.out(vorticity_stretching_flux_field)
            _lib.call("sopht_vorticity_stretching_flux_3d", dt, q, w, u, dt_by_2_dx)
            _lib.call("sopht_elementwise_sum", dt, w, w, q)

    return vorticity_stretching_timestep_euler_forward_pyst_kernel_3d


def gen_vorticity_stretching_timestep_ssprk3_pyst_kernel_3d(
    real_t: type,
    midstep_buffer_vector_field: Any,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int, int] | bool = False,
) -> Callable:
    """SSP-RK3 vortex stretching step (vorticity_stretching_timestep_3d.py:57-158)."""
    dt = _lib.dtype_code(real_t)

    def vorticity_stretching_timestep_ssprk3_pyst_kernel_3d(
        vorticity_field: Any,
        velocity_field: Any,
        vorticity_stretching_flux_field: Any,
        dt_by_2_dx: float,
    ) -> None:
        with _lib.Staging() as s:
            u = s.inp(velocity_field)
            w, q = s.out(vorticity_field), s.out(vorticity_stretching_flux_field)
            mid = s.out(midstep_buffer_vector_field)
            # stage 1: w1 = w + L(w)
            _lib.call("sopht_vorticity_stretching_flux_3d", dt, q, w, u, dt_by_2_dx)
            _lib.call("sopht_elementwise_sum", dt, mid, w, q)
            # stage 2: w2 = 3/4 w + 1/4 (w1 + L(w1))
            _lib.call("sopht_vorticity_stretching_flux_3d", dt, q, mid, u, dt_by_2_dx)
            _lib.call("sopht_elementwise_sum", dt, mid, mid, q)
            _lib.call("sopht_elementwise_saxpby", dt, mid, w, mid, 0.75, 0.25)
            # stage 3: w = 1/3 w + 2/3 (w2 + 1/2 L(w2))
            _lib.call("sopht_vorticity_stretching_flux_3d", dt, q, mid, u, dt_by_2_dx * 0.5)
            _lib.call("sopht_elementwise_sum", dt, mid, mid, q)
            _lib.call("sopht_elementwise_saxpby", dt, w, w, mid, 1.0 / 3.0, 2.0 / 3.0)

    return vorticity_stretching_timestep_ssprk3_pyst_kernel_3d


# ---- ENO3 advection --------------------------------------------------------------------------------------
def gen_advection_flux_conservative_eno3_pyst_kernel_3d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int, int] | bool = False,
) -> Callable:
    """3D conservative ENO3 advection flux generator (advection_flux_3d.py:12-233).

    The six accumulating face kernels of the reference are one launch; the accumulation order into
    ``advection_flux`` (x front, x back, y front, y back, z front, z back) is kept.
    """
    dt = _lib.dtype_code(real_t)

    def advection_flux_conservative_eno3_pyst_kernel_3d(
        advection_flux: Any, field: Any, velocity: Any, inv_dx: float
    ) -> None:
        with _lib.Staging() as s:
            f, v = s.inp(field), s.inp(velocity)
            q = s.out(advection_flux)
            _lib.call("sopht_advection_flux_eno3_3d", dt, q, f, v, inv_dx)

    return advection_flux_conservative_eno3_pyst_kernel_3d


def gen_advection_timestep_euler_forward_conservative_eno3_pyst_kernel_3d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int, int] | bool = False,
    field_type: Literal["scalar", "vector"] = "scalar",
) -> Callable:
    """3D ENO3 advection Euler-forward timestep generator (advection_timestep_3d.py:12-99)."""
    dt = _lib.dtype_code(real_t)
    _check_field_type(field_type)

    def _scalar_step(f: torch.Tensor, q: torch.Tensor, v: torch.Tensor, dt_by_dx: float) -> None:
        _lib.call("sopht_set_fixed_val", dt, q, 0.0)
        _lib.call("sopht_advection_flux_eno3_3d", dt, q, f, v, -float(dt_by_dx))
        _lib.call("sopht_elementwise_sum", dt, f, f, q)

    if field_type == "scalar":

        def advection_timestep_euler_forward_conservative_eno3_pyst_kernel_3d(
            field: Any, advection_flux: Any, velocity: Any, dt_by_dx: float
        ) -> None:
            with _lib.Staging() as s:
                v = s.inp(velocity)
                _scalar_step(s.out(field), s.out(advection_flux), v, dt_by_dx)

        return advection_timestep_euler_forward_conservative_eno3_pyst_kernel_3d

    def vector_field_advection_timestep_euler_forward_conservative_eno3_pyst_kernel_3d(
        vector_field: Any, advection_flux: Any, velocity: Any, dt_by_dx: float
    ) -> None:
        with _lib.Staging() as s:
            v = s.inp(velocity)
            w, q = s.out(vector_field), s.out(advection_flux)
            for c in range(3):
                _scalar_step(w[c], q, v, dt_by_dx)

    return vector_field_advection_timestep_euler_forward_conservative_eno3_pyst_kernel_3d


# ---- boundary penalisation -------------------------------------------------------------------------------
def _sine_ramps(coords: np.ndarray, width: int, dx: float, real_t: type) -> list[float]:
    """Front and back sine ramps along one axis, evaluated in ``real_t`` like the generated kernels."""
    coords = np.asarray(coords).astype(real_t)
    sine_prefactor = real_t((np.pi / 2) / (width * dx))
    start, end = coords[0], coords[-1]
    front = np.sin(sine_prefactor * (coords[:width] - start))
    back = np.sin(sine_prefactor * (end - coords[-width:]))
    return [float(v) for v in front] + [float(v) for v in back]


def gen_penalise_field_boundary_pyst_kernel_3d(
    width: int,
    dx: float,
    x_grid_field: Any,
    y_grid_field: Any,
    z_grid_field: Any,
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int, int] | bool = False,
    field_type: Literal["scalar", "vector"] = "scalar",
) -> Callable:
    """3D penalise field boundary kernel generator (penalise_field_boundary_3d.py:13-240).

    The reference's six broadcast copies + six sliced kernels per component (applied x, then y, then
    z) are restated in separable form and run as three small launches over the shell of the inner box.
    """
    if not isinstance(width, int) or width < 0:
        msg = "Invalid width for boundary zone, must be a non-negative integer"
        raise ValueError(msg)
    _check_field_type(field_type)
    if width == 0:
        if field_type == "scalar":

            def penalise_field_boundary_pyst_kernel_3d(field: Any) -> None:
                pass

            return penalise_field_boundary_pyst_kernel_3d

        def penalise_vector_field_boundary_pyst_kernel_3d(vector_field: Any) -> None:
            pass

        return penalise_vector_field_boundary_pyst_kernel_3d

    dt = _lib.dtype_code(real_t)
    ramp_x = _sine_ramps(_to_numpy(x_grid_field[0, 0, :]), width, dx, real_t)
    ramp_y = _sine_ramps(_to_numpy(y_grid_field[0, :, 0]), width, dx, real_t)
    ramp_z = _sine_ramps(_to_numpy(z_grid_field[:, 0, 0]), width, dx, real_t)

    if field_type == "scalar":

        def penalise_field_boundary_pyst_kernel_3d(field: Any) -> None:  # noqa: F811
            with _lib.Staging() as s:
                _lib.call(
                    "sopht_penalise_field_boundary_3d", dt, s.out(field), width, ramp_x, ramp_y, ramp_z
                )

        return penalise_field_boundary_pyst_kernel_3d

    def penalise_vector_field_boundary_pyst_kernel_3d(vector_field: Any) -> None:  # noqa: F811
        with _lib.Staging() as s:
            _lib.call(
                "sopht_penalise_field_boundary_3d",
                dt,
                s.out(vector_field),
                width,
                ramp_x,
                ramp_y,
                ramp_z,
            )

    return penalise_vector_field_boundary_pyst_kernel_3d


# ---- Brinkman penalisation / characteristic function --------------------------------------------------------
def gen_brinkmann_penalise_pyst_kernel_3d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int, int] | bool = False,
    field_type: Literal["scalar", "vector"] = "scalar",
) -> Callable:
    """Brinkmann penalisation 3D kernel generator (brinkmann_penalise_3d.py:13-83)."""
    dt = _lib.dtype_code(real_t)
    _check_field_type(field_type)
    if field_type == "scalar":

        def brinkmann_penalise_pyst_kernel_3d(
            penalised_field: Any, field: Any, char_field: Any, penalty_field: Any, penalty_factor: float
        ) -> None:
            with _lib.Staging() as s:
                f, chi, pen = s.inp(field), s.inp(char_field), s.inp(penalty_field)
                o = s.out(penalised_field)
                _lib.call("sopht_brinkmann_penalise", dt, o, f, chi, pen, penalty_factor)

        return brinkmann_penalise_pyst_kernel_3d

    def brinkmann_penalise_vector_field_pyst_kernel_3d(
        penalised_vector_field: Any,
        penalty_factor: float,
        char_field: Any,
        penalty_vector_field: Any,
        vector_field: Any,
    ) -> None:
        with _lib.Staging() as s:
            f, chi, pen = s.inp(vector_field), s.inp(char_field), s.inp(penalty_vector_field)
            o = s.out(penalised_vector_field)
            for c in range(3):
                _lib.call("sopht_brinkmann_penalise", dt, o[c], f[c], chi, pen[c], penalty_factor)

    return brinkmann_penalise_vector_field_pyst_kernel_3d


def gen_char_func_from_level_set_via_sine_heaviside_pyst_kernel_3d(
    blend_width: float,
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int, int] | bool = False,
) -> Callable:
    """Smooth sine-Heaviside characteristic function of a level set (char_func_from_level_set_3d.py:12-51)."""
    dt = _lib.dtype_code(real_t)

    def char_func_from_level_set_via_sine_heaviside_pyst_kernel_3d(
        char_func_field: Any, level_set_field: Any
    ) -> None:
        with _lib.Staging() as s:
            ls = s.inp(level_set_field)
            _lib.call("sopht_char_func_from_level_set", dt, s.out(char_func_field), ls, blend_width)

    return char_func_from_level_set_via_sine_heaviside_pyst_kernel_3d


# ---- Laplacian filter ------------------------------------------------------------------------------------
def gen_laplacian_filter_kernel_3d(
    filter_order: int,
    filter_flux_buffer: Any,
    field_buffer: Any,
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int, int] | bool = False,
    field_type: Literal["scalar", "vector"] = "scalar",
    filter_type: Literal["multiplicative", "convolution"] = "multiplicative",
    filter_flux_buffer_boundary_width: int = 1,
) -> Callable:
    """Laplacian filter kernel generator (laplacian_filter_3d.py:13-194).

    Same sub-kernel order and scratch-buffer usage as the reference closures (:95-163).
    """
    if not isinstance(filter_order, int) or filter_order < 0:
        msg = "Invalid filter order, must be a non-negative integer"
        raise ValueError(msg)
    if (
        not isinstance(filter_flux_buffer_boundary_width, int)
        or filter_flux_buffer_boundary_width <= 0
    ):
        msg = "Invalid value for filter flux buffer boundary zone, must be a positive integer"
        raise ValueError(msg)
    dt = _lib.dtype_code(real_t)
    bw = filter_flux_buffer_boundary_width
    axis_x, axis_y, axis_z = 0, 1, 2

    def _filter_pass(flux: torch.Tensor, buf: torch.Tensor, axis: int) -> None:
        _lib.call("sopht_laplacian_filter_flux_3d", dt, flux, buf, axis)
        _lib.call("sopht_elementwise_copy", dt, buf, flux)

    def _multiplicative(f: torch.Tensor, flux: torch.Tensor, buf: torch.Tensor) -> None:
        _lib.call("sopht_set_fixed_val_at_boundaries", dt, flux, bw, [0.0], 0)
        _lib.call("sopht_elementwise_copy", dt, buf, f)
        for _ in range(filter_order):
            _filter_pass(flux, buf, axis_x)
            _filter_pass(flux, buf, axis_y)
            _filter_pass(flux, buf, axis_z)
        _lib.call("sopht_elementwise_saxpby", dt, f, f, flux, 1.0, -1.0)

    def _convolution(f: torch.Tensor, flux: torch.Tensor, buf: torch.Tensor) -> None:
        _lib.call("sopht_set_fixed_val_at_boundaries", dt, flux, bw, [0.0], 0)
        for axis in (axis_x, axis_y, axis_z):
            _lib.call("sopht_elementwise_copy", dt, buf, f)
            for _ in range(filter_order):
                _filter_pass(flux, buf, axis)
            _lib.call("sopht_elementwise_saxpby", dt, f, f, flux, 1.0, -1.0)

    def _fusable(f: torch.Tensor, buf: torch.Tensor) -> bool:
        # order 0 subtracts whatever the flux buffer holds (reference behaviour): only the pass-by-pass path has it
        return (1 <= filter_order <= 64 and f.stride(-1) == 1 and buf.stride(-1) == 1 and min(f.shape[-3:]) >= 3
                and 3 * f.shape[-1] * f.element_size() <= 96 * 1024)  # a row and its two work copies in shared memory

    def _convolution_fused_or_not(f: torch.Tensor, flux: torch.Tensor, buf: torch.Tensor) -> None:
        """One line kernel per direction (csrc/filter3d.cu) instead of 5 + 4 order passes; f is a scalar field or
        the whole vector field."""
        if _fusable(f, buf):
            _lib.call("sopht_laplacian_filter_convolution_3d", dt, f, buf, filter_order)
        elif f.dim() == 4:
            for c in range(3):
                _convolution(f[c], flux, buf)
        else:
            _convolution(f, flux, buf)

    fused_vector_impl = None
    if filter_type == "multiplicative":
        scalar_impl = _multiplicative
    elif filter_type == "convolution":
        scalar_impl = _convolution_fused_or_not
        fused_vector_impl = _convolution_fused_or_not
    else:
        msg = "Invalid filter type"
        raise ValueError(msg)

    if field_type == "scalar":

        def scalar_field_filter_kernel_3d(scalar_field: Any) -> None:
            with _lib.Staging() as s:
                flux, buf = s.out(filter_flux_buffer), s.out(field_buffer)
                scalar_impl(s.out(scalar_field), flux, buf)

        return scalar_field_filter_kernel_3d
    if field_type == "vector":

        def vector_field_filter_kernel_3d(vector_field: Any) -> None:
            with _lib.Staging() as s:
                flux, buf = s.out(filter_flux_buffer), s.out(field_buffer)
                v = s.out(vector_field)
                if fused_vector_impl is not None:
                    fused_vector_impl(v, flux, buf)
                    return
                for c in range(3):
                    scalar_impl(v[c], flux, buf)

        return vector_field_filter_kernel_3d
    msg = "Invalid field type"
    raise ValueError(msg)


__all__ = [
    "gen_advection_flux_conservative_eno3_pyst_kernel_3d",
    "gen_advection_timestep_euler_forward_conservative_eno3_pyst_kernel_3d",
    "gen_brinkmann_penalise_pyst_kernel_3d",
    "gen_char_func_from_level_set_via_sine_heaviside_pyst_kernel_3d",
    "gen_curl_pyst_kernel_3d",
    "gen_diffusion_flux_pyst_kernel_3d",
    "gen_diffusion_timestep_euler_forward_pyst_kernel_3d",
    "gen_divergence_pyst_kernel_3d",
    "gen_laplacian_filter_kernel_3d",
    "gen_penalise_field_boundary_pyst_kernel_3d",
    "gen_update_vorticity_from_penalised_velocity_pyst_kernel_3d",
    "gen_update_vorticity_from_velocity_forcing_pyst_kernel_3d",
    "gen_vorticity_stretching_flux_pyst_kernel_3d",
    "gen_vorticity_stretching_timestep_euler_forward_pyst_kernel_3d",
    "gen_vorticity_stretching_timestep_ssprk3_pyst_kernel_3d",
]
