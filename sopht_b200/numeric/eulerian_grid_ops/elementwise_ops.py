"""Elementwise kernel factories, 2-D and 3-D.

Drop-in counterparts of sopht/numeric/eulerian_grid_ops/stencil_ops_3d/elementwise_ops_3d.py and
stencil_ops_2d/elementwise_ops_2d.py: same factory names, keyword names and ValueErrors. Each
returned callable mutates its output argument in place and returns None; it enqueues ONE CUDA kernel
(per component for the per-component vector closures) on torch's current stream through the C ABI.
``num_threads`` and ``fixed_grid_size`` are accepted and ignored (the reference's ``fixed_grid_size``
test is dead code, elementwise_ops_3d.py:28-30).
"""

from __future__ import annotations

from collections.abc import Callable
from typing import Any, Literal

from sopht_b200 import _lib

_INVALID_FIELD_TYPE = "Invalid field type"


def _check_field_type(field_type: str) -> None:
    if field_type not in ("scalar", "vector"):
        raise ValueError(_INVALID_FIELD_TYPE)


def _gen_elementwise_sum(real_t: type, field_type: str) -> Callable:
    dt = _lib.dtype_code(real_t)
    _check_field_type(field_type)

    def elementwise_sum_kernel(sum_field: Any, field_1: Any, field_2: Any) -> None:
        """sum_field = field_1 + field_2 (elementwise_ops_3d.py:13-57)."""
        with _lib.Staging() as s:
            a, b, o = s.inp(field_1), s.inp(field_2), s.out(sum_field)
            _lib.call("sopht_elementwise_sum", dt, o, a, b)

    return elementwise_sum_kernel


def _gen_set_fixed_val(real_t: type, field_type: str) -> Callable:
    dt = _lib.dtype_code(real_t)
    _check_field_type(field_type)
    if field_type == "scalar":

        def set_fixed_val_kernel(field: Any, fixed_val: float) -> None:
            """field[...] = fixed_val (elementwise_ops_3d.py:77-86)."""
            with _lib.Staging() as s:
                _lib.call("sopht_set_fixed_val", dt, s.out(field), fixed_val)

        return set_fixed_val_kernel

    def vector_field_set_fixed_val_kernel(vector_field: Any, fixed_vals: Any) -> None:
        """vector_field[c] = fixed_vals[c] (elementwise_ops_3d.py:95-115)."""
        with _lib.Staging() as s:
            v = s.out(vector_field)
            _lib.call("sopht_set_fixed_vals_vector", dt, v, list(fixed_vals), len(fixed_vals))

    return vector_field_set_fixed_val_kernel


def _gen_elementwise_copy(real_t: type) -> Callable:
    dt = _lib.dtype_code(real_t)

    def elementwise_copy_kernel(field: Any, rhs_field: Any) -> None:
        """field[...] = rhs_field (elementwise_ops_3d.py:122-143)."""
        with _lib.Staging() as s:
            r, f = s.inp(rhs_field), s.out(field)
            _lib.call("sopht_elementwise_copy", dt, f, r)

    return elementwise_copy_kernel


def _gen_elementwise_complex_product(real_t: type) -> Callable:
    dt = _lib.dtype_code(real_t)

    def elementwise_complex_product_kernel(product_field: Any, field_1: Any, field_2: Any) -> None:
        """Complex product of two complex fields (elementwise_ops_3d.py:146-197)."""
        with _lib.Staging() as s:
            a, b, o = s.inp(field_1), s.inp(field_2), s.out(product_field)
            _lib.call("sopht_elementwise_complex_product", dt, o, a, b)

    return elementwise_complex_product_kernel


def _gen_set_fixed_val_at_boundaries(real_t: type, width: int, field_type: str) -> Callable:
    if not isinstance(width, int) or width <= 0:
        msg = "Invalid width for boundary zone, must be a positive integer"
        raise ValueError(msg)
    dt = _lib.dtype_code(real_t)
    _check_field_type(field_type)
    if field_type == "scalar":

        def set_fixed_val_at_boundaries_kernel(field: Any, fixed_val: float) -> None:
            """Ring of `width` cells <- fixed_val (elementwise_ops_3d.py:219-233)."""
            with _lib.Staging() as s:
                _lib.call(
                    "sopht_set_fixed_val_at_boundaries", dt, s.out(field), width, [fixed_val], 0
                )

        return set_fixed_val_at_boundaries_kernel

    def vector_field_set_fixed_val_at_boundaries_kernel(vector_field: Any, fixed_vals: Any) -> None:
        """Ring of `width` cells of every component <- fixed_vals[c] (elementwise_ops_3d.py:238-263)."""
        with _lib.Staging() as s:
            _lib.call(
                "sopht_set_fixed_val_at_boundaries",
                dt,
                s.out(vector_field),
                width,
                list(fixed_vals),
                1,
            )

    return vector_field_set_fixed_val_at_boundaries_kernel


def _gen_add_fixed_val(real_t: type, field_type: str) -> Callable:
    dt = _lib.dtype_code(real_t)
    _check_field_type(field_type)
    if field_type == "scalar":

        def add_fixed_val_kernel(sum_field: Any, field: Any, fixed_val: float) -> None:
            """sum_field = field + fixed_val (elementwise_ops_3d.py:288-297)."""
            with _lib.Staging() as s:
                f, o = s.inp(field), s.out(sum_field)
                _lib.call("sopht_add_fixed_val", dt, o, f, fixed_val)

        return add_fixed_val_kernel

    def vector_field_add_fixed_val_kernel(sum_field: Any, vector_field: Any, fixed_vals: Any) -> None:
        """sum_field[c] = vector_field[c] + fixed_vals[c] (elementwise_ops_3d.py:305-329)."""
        with _lib.Staging() as s:
            f, o = s.inp(vector_field), s.out(sum_field)
            _lib.call("sopht_add_fixed_vals_vector", dt, o, f, list(fixed_vals), len(fixed_vals))

    return vector_field_add_fixed_val_kernel


def _gen_elementwise_saxpby(real_t: type, field_type: str) -> Callable:
    dt = _lib.dtype_code(real_t)
    _check_field_type(field_type)

    def elementwise_saxpby_kernel(
        sum_field: Any, field_1: Any, field_2: Any, field_1_prefac: float, field_2_prefac: float
    ) -> None:
        """sum_field = field_1_prefac * field_1 + field_2_prefac * field_2 (elementwise_ops_3d.py:337-387)."""
        with _lib.Staging() as s:
            a, b, o = s.inp(field_1), s.inp(field_2), s.out(sum_field)
            _lib.call("sopht_elementwise_saxpby", dt, o, a, b, field_1_prefac, field_2_prefac)

    return elementwise_saxpby_kernel


# ---- public factories: 3-D ---------------------------------------------------------------------------
def gen_elementwise_sum_pyst_kernel_3d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int, int] | bool = False,
    field_type: Literal["scalar", "vector"] = "scalar",
) -> Callable:
    """3D elementwise sum kernel generator."""
    return _gen_elementwise_sum(real_t, field_type)


def gen_set_fixed_val_pyst_kernel_3d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int, int] | bool = False,
    field_type: Literal["scalar", "vector"] = "scalar",
) -> Callable:
    """3D set field to fixed value kernel generator."""
    return _gen_set_fixed_val(real_t, field_type)


def gen_elementwise_copy_pyst_kernel_3d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int, int] | bool = False,
) -> Callable:
    """3D elementwise copy one field to another kernel generator."""
    return _gen_elementwise_copy(real_t)


def gen_elementwise_complex_product_pyst_kernel_3d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int, int] | bool = False,
) -> Callable:
    """3D elementwise complex number product kernel generator."""
    return _gen_elementwise_complex_product(real_t)


def gen_set_fixed_val_at_boundaries_pyst_kernel_3d(
    real_t: type,
    width: int,
    num_threads: bool | int = False,
    field_type: Literal["scalar", "vector"] = "scalar",
) -> Callable:
    """3D set field to fixed value at boundaries kernel generator."""
    return _gen_set_fixed_val_at_boundaries(real_t, width, field_type)


def gen_add_fixed_val_pyst_kernel_3d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int, int] | bool = False,
    field_type: Literal["scalar", "vector"] = "scalar",
) -> Callable:
    """3D add a fixed value to a field kernel generator."""
    return _gen_add_fixed_val(real_t, field_type)


def gen_elementwise_saxpby_pyst_kernel_3d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int, int] | bool = False,
    field_type: Literal["scalar", "vector"] = "scalar",
) -> Callable:
    """3D elementwise saxpby (s = a * x + b * y) kernel generator."""
    return _gen_elementwise_saxpby(real_t, field_type)


def gen_elementwise_cross_product_pyst_kernel_3d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int, int] | bool = False,
) -> Callable:
    """3D elementwise cross product kernel generator (elementwise_ops_3d.py:390-449)."""
    dt = _lib.dtype_code(real_t)

    def elementwise_cross_product_pyst_kernel_3d(result_field: Any, field_1: Any, field_2: Any) -> None:
        """Elementwise cross product of two (3, nz, ny, nx) vector fields, all three components in one launch."""
        with _lib.Staging() as s:
            a, b, o = s.inp(field_1), s.inp(field_2), s.out(result_field)
            _lib.call("sopht_elementwise_cross_product_3d", dt, o, a, b)

    return elementwise_cross_product_pyst_kernel_3d


# ---- public factories: 2-D ---------------------------------------------------------------------------
def gen_elementwise_sum_pyst_kernel_2d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int] | bool = False,
    field_type: Literal["scalar", "vector"] = "scalar",
) -> Callable:
    """2D elementwise sum kernel generator."""
    return _gen_elementwise_sum(real_t, field_type)


def gen_set_fixed_val_pyst_kernel_2d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int] | bool = False,
    field_type: Literal["scalar", "vector"] = "scalar",
) -> Callable:
    """2D set field to fixed value kernel generator."""
    return _gen_set_fixed_val(real_t, field_type)


def gen_elementwise_copy_pyst_kernel_2d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int] | bool = False,
) -> Callable:
    """2D elementwise copy one field to another kernel generator."""
    return _gen_elementwise_copy(real_t)


def gen_elementwise_complex_product_pyst_kernel_2d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int] | bool = False,
) -> Callable:
    """2D elementwise complex number product kernel generator."""
    return _gen_elementwise_complex_product(real_t)


def gen_set_fixed_val_at_boundaries_pyst_kernel_2d(
    real_t: type,
    width: int,
    num_threads: bool | int = False,
    field_type: Literal["scalar", "vector"] = "scalar",
) -> Callable:
    """2D set field to fixed value at boundaries kernel generator."""
    return _gen_set_fixed_val_at_boundaries(real_t, width, field_type)


def gen_add_fixed_val_pyst_kernel_2d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int] | bool = False,
    field_type: Literal["scalar", "vector"] = "scalar",
) -> Callable:
    """2D add a fixed value to a field kernel generator."""
    return _gen_add_fixed_val(real_t, field_type)


def gen_elementwise_saxpby_pyst_kernel_2d(
    real_t: type,
    num_threads: bool | int = False,
    fixed_grid_size: tuple[int, int] | bool = False,
    field_type: Literal["scalar", "vector"] = "scalar",
) -> Callable:
    """2D elementwise saxpby (s = a * x + b * y) kernel generator."""
    return _gen_elementwise_saxpby(real_t, field_type)
