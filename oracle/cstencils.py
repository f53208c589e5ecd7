"""CPU ORACLE (test infrastructure, never on the product path): the same kernel vocabulary as oracle/stencils.py,
executed by the C / OpenMP restatement in oracle/c/ref_kernels.c - one OpenMP loop nest per pystencils kernel of the
reference, composed here in the reference's own launch order (file:line cited per function). This is the CPU baseline
bench.py reports beside the GPU number ("one OpenMP C++ loop nest per reference kernel, same unfused
passes"); tests/test_oracle_golden.py pins it to the same golden vectors as the numpy oracle.

Anything this module does not define (or a call with arrays the C kernels do not take: non-contiguous views, mixed
dtypes) falls through to the numpy restatement in oracle/stencils.py, so `cstencils.<name>` exists for every name
there.
"""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from . import stencils as _np_impl

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsopht_ref_kernels.so")
_lib = None


def load():
    """Load (building it first if necessary - gcc is part of the image) the C / OpenMP kernel library."""
    global _lib
    if _lib is not None:
        return _lib
    src = [os.path.join(_HERE, "c", n) for n in ("ref_kernels.c", "ref_kernels_body.inc")]
    stale = not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if stale:
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    try:
        _lib = ctypes.CDLL(_SO)
    except OSError:  # built on a different machine: rebuild here
        subprocess.run(["make", "-C", _HERE, "clean", "all"], check=True, capture_output=True)
        _lib = ctypes.CDLL(_SO)
    assert _lib.sopht_ref_kernels_abi() == 1
    return _lib


def set_num_threads(n: int) -> None:
    """OpenMP thread count of the kernels (libgomp reads OMP_NUM_THREADS at load; this sets it at run time)."""
    load()
    gomp = ctypes.CDLL("libgomp.so.1")
    gomp.omp_set_num_threads(int(n))


_P = ctypes.c_void_p
_I64 = ctypes.c_int64
_INT = ctypes.c_int


def _suffix(a: np.ndarray) -> str | None:
    if a.dtype == np.float32:
        return "_f32"
    if a.dtype == np.float64:
        return "_f64"
    return None


def _real(a: np.ndarray, v):
    return (ctypes.c_float if a.dtype == np.float32 else ctypes.c_double)(float(a.dtype.type(v)))


def _ok(*arrays: np.ndarray) -> bool:
    d = arrays[0].dtype
    return all(isinstance(a, np.ndarray) and a.dtype == d and a.flags.c_contiguous for a in arrays) and \
        _suffix(arrays[0]) is not None


def _fn(name: str, a: np.ndarray):
    return getattr(load(), name + _suffix(a))


def _ptr(a: np.ndarray) -> ctypes.c_void_p:
    return _P(a.ctypes.data)


def _dims(a: np.ndarray):
    """(nz, ny, nx) with nz = 1 for 2-D arrays."""
    return (1, *a.shape) if a.ndim == 2 else a.shape


# ---- elementwise ---------------------------------------------------------------------------------------------------
def elementwise_sum(sum_field, field_1, field_2):  # elementwise_ops_3d.py:13-57
    if not _ok(sum_field, field_1, field_2):
        return _np_impl.elementwise_sum(sum_field, field_1, field_2)
    _fn("ew_sum", sum_field)(_ptr(sum_field), _ptr(field_1), _ptr(field_2), _I64(sum_field.size))


def set_fixed_val(field, fixed_val):  # :60-119
    if not _ok(field):
        return _view_set(field, fixed_val)
    _fn("ew_set", field)(_ptr(field), _real(field, fixed_val), _I64(field.size))


def set_fixed_val_vector(vector_field, fixed_vals):  # :95-115 (one launch per component)
    for c in range(vector_field.shape[0]):
        set_fixed_val(vector_field[c], fixed_vals[c])


def elementwise_copy(field, rhs_field):  # :122-143
    if not _ok(field, rhs_field):
        return _view_copy(field, rhs_field)
    _fn("ew_copy", field)(_ptr(field), _ptr(rhs_field), _I64(field.size))


def _row_view(a: np.ndarray):
    """(pointer, sz, sy, nz, ny, nx) of a view whose rows are contiguous, or None."""
    if a.ndim not in (2, 3) or _suffix(a) is None or a.strides[-1] != a.itemsize:
        return None
    st = [s // a.itemsize for s in a.strides]
    if a.ndim == 2:
        return _ptr(a), _I64(0), _I64(st[0]), _INT(1), _INT(a.shape[0]), _INT(a.shape[1])
    return _ptr(a), _I64(st[0]), _I64(st[1]), _INT(a.shape[0]), _INT(a.shape[1]), _INT(a.shape[2])


def _view_set(field, fixed_val):
    v = _row_view(field)
    if v is None:
        return _np_impl.set_fixed_val(field, fixed_val)
    p, sz, sy, nz, ny, nx = v
    _fn("view_set", field)(p, sz, sy, _real(field, fixed_val), nz, ny, nx)


def _view_copy(field, rhs_field):
    a, b = _row_view(field), _row_view(rhs_field)
    if a is None or b is None or field.dtype != rhs_field.dtype or field.shape != rhs_field.shape:
        return _np_impl.elementwise_copy(field, rhs_field)
    _fn("view_copy", field)(a[0], a[1], a[2], b[0], b[1], b[2], a[3], a[4], a[5])


def elementwise_complex_product(product_field, field_1, field_2):  # :146-197
    arrs = (product_field, field_1, field_2)
    if not all(a.flags.c_contiguous and a.dtype == product_field.dtype for a in arrs) or \
            product_field.dtype not in (np.complex64, np.complex128):
        return _np_impl.elementwise_complex_product(product_field, field_1, field_2)
    name = "ew_complex_product" + ("_f32" if product_field.dtype == np.complex64 else "_f64")
    getattr(load(), name)(_ptr(product_field), _ptr(field_1), _ptr(field_2), _I64(product_field.size))


def set_fixed_val_at_boundaries(field, width, fixed_val):  # :200-233: one launch per boundary slab
    for ax in range(field.ndim):
        front = [slice(None)] * field.ndim
        back = [slice(None)] * field.ndim
        front[ax] = slice(None, width)
        back[ax] = slice(-width, None)
        _view_set(field[tuple(front)], fixed_val)
        _view_set(field[tuple(back)], fixed_val)


def set_fixed_val_at_boundaries_vector(vector_field, width, fixed_vals):  # :236-263
    for c in range(vector_field.shape[0]):
        set_fixed_val_at_boundaries(vector_field[c], width, fixed_vals[c])


def add_fixed_val(sum_field, field, fixed_val):  # :271-334
    if not _ok(sum_field, field):
        return _np_impl.add_fixed_val(sum_field, field, fixed_val)
    _fn("ew_add_val", field)(_ptr(sum_field), _ptr(field), _real(field, fixed_val), _I64(field.size))


def add_fixed_val_vector(sum_field, vector_field, fixed_vals):
    for c in range(vector_field.shape[0]):
        add_fixed_val(sum_field[c], vector_field[c], fixed_vals[c])


def elementwise_saxpby(sum_field, field_1, field_2, field_1_prefac, field_2_prefac):  # :337-387
    if not _ok(sum_field, field_1, field_2):
        return _np_impl.elementwise_saxpby(sum_field, field_1, field_2, field_1_prefac, field_2_prefac)
    _fn("ew_saxpby", sum_field)(_ptr(sum_field), _ptr(field_1), _ptr(field_2), _real(sum_field, field_1_prefac),
                                _real(sum_field, field_2_prefac), _I64(sum_field.size))


def elementwise_cross_product(result_field, field_1, field_2):  # :390-449 (three launches)
    if not _ok(result_field, field_1, field_2):
        return _np_impl.elementwise_cross_product(result_field, field_1, field_2)
    a, b, r = field_1, field_2, result_field
    f = _fn("ew_cross_component", r)
    n = _I64(r[0].size)
    f(_ptr(r[0]), _ptr(a[1]), _ptr(b[2]), _ptr(b[1]), _ptr(a[2]), n)
    f(_ptr(r[1]), _ptr(a[2]), _ptr(b[0]), _ptr(b[2]), _ptr(a[0]), n)
    f(_ptr(r[2]), _ptr(a[0]), _ptr(b[1]), _ptr(b[0]), _ptr(a[1]), n)


# ---- diffusion -----------------------------------------------------------------------------------------------------
def diffusion_flux(diffusion_flux_, field, prefactor, reset_ghost_zone=True):
    """diffusion_flux_3d.py:14-115 / diffusion_flux_2d.py:13-72."""
    if not _ok(diffusion_flux_, field):
        return _np_impl.diffusion_flux(diffusion_flux_, field, prefactor, reset_ghost_zone)
    nz, ny, nx = _dims(field)
    _fn("diffusion_flux", field)(_ptr(diffusion_flux_), _ptr(field), _real(field, prefactor), _INT(nz), _INT(ny),
                                 _INT(nx))
    if reset_ghost_zone:
        set_fixed_val_at_boundaries(diffusion_flux_, 1, 0)


def diffusion_flux_vector(vector_flux, vector_field, prefactor, reset_ghost_zone=True):
    for c in range(vector_field.shape[0]):
        diffusion_flux(vector_flux[c], vector_field[c], prefactor, reset_ghost_zone)


def diffusion_timestep_euler_forward(field, diffusion_flux_, nu_dt_by_dx2):
    """diffusion_timestep_3d.py:31-44 / diffusion_timestep_2d.py:28-43."""
    diffusion_flux(diffusion_flux_, field, nu_dt_by_dx2, True)
    elementwise_sum(field, field, diffusion_flux_)


def diffusion_timestep_euler_forward_vector(vector_field, diffusion_flux_, nu_dt_by_dx2):
    for c in range(vector_field.shape[0]):  # diffusion_timestep_3d.py:62-78
        diffusion_timestep_euler_forward(vector_field[c], diffusion_flux_, nu_dt_by_dx2)


# ---- curl / forcing update -----------------------------------------------------------------------------------------
def _curl_launches(out, f, p, accumulate):
    """Three launches (x, y, z components) of the centred curl, curl_3d.py:37-66."""
    nz, ny, nx = f.shape[1:]
    sx, sy, sz = _I64(1), _I64(nx), _I64(ny * nx)
    fn = _fn("curl_component", f)
    pr, acc, d = _real(f, p), _INT(accumulate), (_INT(nz), _INT(ny), _INT(nx))
    fn(_ptr(out[0]), _ptr(f[2]), sy, _ptr(f[1]), sz, pr, acc, *d)  # c_x = dy f_z - dz f_y
    fn(_ptr(out[1]), _ptr(f[0]), sz, _ptr(f[2]), sx, pr, acc, *d)  # c_y = dz f_x - dx f_z
    fn(_ptr(out[2]), _ptr(f[1]), sx, _ptr(f[0]), sy, pr, acc, *d)  # c_z = dx f_y - dy f_x


def curl_3d(curl, field, prefactor, reset_ghost_zone=True):  # curl_3d.py:13-132
    if not _ok(curl, field):
        return _np_impl.curl_3d(curl, field, prefactor, reset_ghost_zone)
    _curl_launches(curl, field, prefactor, 0)
    if reset_ghost_zone:
        set_fixed_val_at_boundaries_vector(curl, 1, [0, 0, 0])


def update_vorticity_from_velocity_forcing_3d(vorticity_field, velocity_forcing_field, prefactor):
    """update_vorticity_from_velocity_forcing_3d.py:12-132."""
    if not _ok(vorticity_field, velocity_forcing_field):
        return _np_impl.update_vorticity_from_velocity_forcing_3d(vorticity_field, velocity_forcing_field, prefactor)
    _curl_launches(vorticity_field, velocity_forcing_field, prefactor, 1)


def update_vorticity_from_velocity_forcing_2d(vorticity_field, velocity_forcing_field, prefactor):
    """update_vorticity_from_velocity_forcing_2d.py:12-72."""
    if not _ok(vorticity_field, velocity_forcing_field):
        return _np_impl.update_vorticity_from_velocity_forcing_2d(vorticity_field, velocity_forcing_field, prefactor)
    ny, nx = vorticity_field.shape
    f = velocity_forcing_field
    _fn("forcing_update_2d", f)(_ptr(vorticity_field), _ptr(f[0]), _ptr(f[1]), _real(f, prefactor), _INT(ny), _INT(nx))


def outplane_field_curl_2d(curl, field, prefactor, reset_ghost_zone=True):  # outplane_field_curl_2d.py:13-100
    if not _ok(curl, field):
        return _np_impl.outplane_field_curl_2d(curl, field, prefactor, reset_ghost_zone)
    ny, nx = field.shape
    _fn("outplane_curl_2d_x", field)(_ptr(curl[0]), _ptr(field), _real(field, prefactor), _INT(ny), _INT(nx))
    _fn("outplane_curl_2d_y", field)(_ptr(curl[1]), _ptr(field), _real(field, prefactor), _INT(ny), _INT(nx))
    if reset_ghost_zone:
        set_fixed_val_at_boundaries_vector(curl, 1, [0, 0])


# ---- ENO3 conservative advection -----------------------------------------------------------------------------------
def advection_flux_conservative_eno3(advection_flux, field, velocity, inv_dx):
    """advection_flux_3d.py:12-233 / advection_flux_2d.py:12-165: front and back face kernels per axis, x first."""
    if not _ok(advection_flux, field, velocity):
        return _np_impl.advection_flux_conservative_eno3(advection_flux, field, velocity, inv_dx)
    nz, ny, nx = _dims(field)
    strides = [1, nx, ny * nx]
    fn = _fn("eno3_face", field)
    for comp in range(field.ndim):
        for front in (1, 0):
            fn(_ptr(advection_flux), _ptr(field), _ptr(velocity[comp]), _I64(strides[comp]), _INT(front),
               _real(field, inv_dx), _INT(nz), _INT(ny), _INT(nx))


def advection_timestep_euler_forward_conservative_eno3(field, advection_flux, velocity, dt_by_dx):
    """advection_timestep_3d.py:37-56 / advection_timestep_2d.py:36-55."""
    set_fixed_val(advection_flux, 0)
    advection_flux_conservative_eno3(advection_flux, field, velocity, -field.dtype.type(dt_by_dx))
    elementwise_sum(field, field, advection_flux)


def advection_timestep_euler_forward_conservative_eno3_vector(vector_field, advection_flux, velocity, dt_by_dx):
    for c in range(vector_field.shape[0]):  # advection_timestep_3d.py:62-95
        advection_timestep_euler_forward_conservative_eno3(vector_field[c], advection_flux, velocity, dt_by_dx)


# ---- Laplacian filter ----------------------------------------------------------------------------------------------
def _filter_flux(flux, field, axis):
    nz, ny, nx = field.shape
    s = [ny * nx, nx, 1][axis]
    _fn("filter_flux", field)(_ptr(flux), _ptr(field), _I64(s), _INT(nz), _INT(ny), _INT(nx))


def laplacian_filter_3d(scalar_field, filter_flux_buffer, field_buffer, filter_order, filter_type="multiplicative",
                        boundary_width=1):
    """laplacian_filter_3d.py:95-163: the same pass sequence as the numpy oracle, every pass a C kernel."""
    if not _ok(scalar_field, filter_flux_buffer, field_buffer) or boundary_width != 1:
        return _np_impl.laplacian_filter_3d(scalar_field, filter_flux_buffer, field_buffer, filter_order, filter_type,
                                            boundary_width)
    f, flux, buf = scalar_field, filter_flux_buffer, field_buffer
    set_fixed_val_at_boundaries(flux, boundary_width, 0)  # the scratch ring is held at 0 (laplacian_filter_3d.py:99)
    if filter_type == "multiplicative":
        elementwise_copy(buf, f)
        for _ in range(filter_order):
            for axis in (2, 1, 0):  # x, y, z
                _filter_flux(flux, buf, axis)
                elementwise_copy(buf, flux)
        elementwise_saxpby(f, f, flux, 1.0, -1.0)
    elif filter_type == "convolution":
        for axis in (2, 1, 0):
            elementwise_copy(buf, f)
            for _ in range(filter_order):
                _filter_flux(flux, buf, axis)
                elementwise_copy(buf, flux)
            elementwise_saxpby(f, f, flux, 1.0, -1.0)
    else:
        msg = "Invalid filter type"
        raise ValueError(msg)


def laplacian_filter_3d_vector(vector_field, filter_flux_buffer, field_buffer, filter_order,
                               filter_type="multiplicative", boundary_width=1):
    for c in range(vector_field.shape[0]):
        laplacian_filter_3d(vector_field[c], filter_flux_buffer, field_buffer, filter_order, filter_type,
                            boundary_width)


def abs_sum_max(velocity_magnitude_field, velocity_field):
    """passive_transport_flow_simulators.py:150-151: buf = sum_c |u_c| (side effect) and its maximum."""
    if not _ok(velocity_magnitude_field, velocity_field):
        velocity_magnitude_field[...] = np.sum(np.fabs(velocity_field), axis=0)
        return np.amax(velocity_magnitude_field)
    fn = _fn("abs_sum_max", velocity_field)
    fn.restype = ctypes.c_float if velocity_field.dtype == np.float32 else ctypes.c_double
    return velocity_field.dtype.type(
        fn(_ptr(velocity_magnitude_field), _ptr(velocity_field), _INT(velocity_field.shape[0]),
           _I64(velocity_magnitude_field.size)))


def __getattr__(name):  # everything else: the numpy restatement
    return getattr(_np_impl, name)
