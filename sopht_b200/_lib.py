"""ctypes binding of libsopht_b200.so (the C ABI declared in include/sopht_b200.h).

Host-side plumbing only: turns torch CUDA tensors (or numpy arrays, staged through the device) into
``sopht_field_t`` descriptors and forwards the call on torch's current CUDA stream. There is no CPU
fallback: if the shared library is missing every kernel call raises.
"""

from __future__ import annotations

import ctypes
import os
from typing import Any

import numpy as np
import torch

MAX_DIMS = 5
SOPHT_F32 = 0
SOPHT_F64 = 1

# SOPHT_B200_LIB: an experiment build of the same library (Makefile: EXTRA= / LIBNAME=)
_LIB_PATH = os.environ.get("SOPHT_B200_LIB") or os.path.join(
    os.path.dirname(os.path.abspath(__file__)), "lib", "libsopht_b200.so")


class SophtField(ctypes.Structure):
    _fields_ = [
        ("data", ctypes.c_void_p),
        ("ndim", ctypes.c_int32),
        ("shape", ctypes.c_int64 * MAX_DIMS),
        ("stride", ctypes.c_int64 * MAX_DIMS),
    ]


_F = ctypes.POINTER(SophtField)
_D = ctypes.c_double
_I = ctypes.c_int
_P = ctypes.c_void_p
_PD = ctypes.POINTER(ctypes.c_double)
_PI64 = ctypes.POINTER(ctypes.c_int64)

# name -> argument kinds after the leading `int dtype`; every function ends with `void* stream`
# (appended automatically) unless listed in _NO_STREAM.
_SIGNATURES: dict[str, list[Any]] = {
    "sopht_set_fixed_val": [_F, _D],
    "sopht_set_fixed_vals_vector": [_F, _PD, _I],
    "sopht_elementwise_copy": [_F, _F],
    "sopht_elementwise_sum": [_F, _F, _F],
    "sopht_elementwise_saxpby": [_F, _F, _F, _D, _D],
    "sopht_add_fixed_val": [_F, _F, _D],
    "sopht_add_fixed_vals_vector": [_F, _F, _PD, _I],
    "sopht_elementwise_complex_product": [_F, _F, _F],
    "sopht_elementwise_cross_product_3d": [_F, _F, _F],
    "sopht_set_fixed_val_at_boundaries": [_F, _I, _PD, _I],
    "sopht_brinkmann_penalise": [_F, _F, _F, _F, _D],
    "sopht_brinkmann_penalise_vs_fixed_val": [_F, _F, _F, _D, _D],
    "sopht_char_func_from_level_set": [_F, _F, _D],
    "sopht_abs_sum_max": [_F, _F, _P],
    "sopht_wrap_z_halos": [_F],
    "sopht_diffusion_flux_3d": [_F, _F, _D, _I],
    "sopht_curl_3d": [_F, _F, _D, _I],
    "sopht_divergence_3d": [_F, _F, _D, _I],
    "sopht_update_vorticity_from_velocity_forcing_3d": [_F, _F, _D],
    "sopht_update_vorticity_from_penalised_velocity_3d": [_F, _F, _F, _D],
    "sopht_vorticity_stretching_flux_3d": [_F, _F, _F, _D],
    "sopht_advection_flux_eno3_3d": [_F, _F, _F, _D],
    "sopht_laplacian_filter_flux_3d": [_F, _F, _I],
    "sopht_laplacian_filter_convolution_3d": [_F, _F, _I],
    "sopht_penalise_field_boundary_3d": [_F, _I, _PD, _PD, _PD],
    "sopht_penalise_field_boundary_3d_slab": [_F, _I, _PD, _PD, _PD, _I],
    "sopht_diffusion_flux_2d": [_F, _F, _D, _I],
    "sopht_advection_flux_eno3_2d": [_F, _F, _F, _D],
    "sopht_outplane_field_curl_2d": [_F, _F, _D, _I],
    "sopht_inplane_field_curl_2d": [_F, _F, _D],
    "sopht_update_vorticity_from_velocity_forcing_2d": [_F, _F, _D],
    "sopht_update_vorticity_from_penalised_velocity_2d": [_F, _F, _F, _D],
    "sopht_penalise_field_boundary_2d": [_F, _I, _PD, _PD],
}

_lib: ctypes.CDLL | None = None


class SophtLibraryError(RuntimeError):
    """The CUDA extension is missing or failed; there is no fallback path."""


def lib_path() -> str:
    return _LIB_PATH


def load() -> ctypes.CDLL:
    """Load libsopht_b200.so once; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise SophtLibraryError(
            f"{_LIB_PATH} not found: build it with `make` (or __graft_entry__.build()). "
            "sopht_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(_LIB_PATH)
    lib.sopht_last_error.restype = ctypes.c_char_p
    lib.sopht_last_error.argtypes = []
    lib.sopht_version.restype = ctypes.c_int
    lib.sopht_launch_count.restype = ctypes.c_int64
    lib.sopht_profile_enable.restype = ctypes.c_int
    lib.sopht_profile_enable.argtypes = [ctypes.c_int]
    lib.sopht_profile_report.restype = ctypes.c_char_p
    lib.sopht_profile_report.argtypes = []
    lib.sopht_profile_range_begin.restype = ctypes.c_int
    lib.sopht_profile_range_begin.argtypes = [ctypes.c_char_p, ctypes.c_void_p]
    lib.sopht_profile_range_end.restype = ctypes.c_int
    lib.sopht_profile_range_end.argtypes = [ctypes.c_int, ctypes.c_void_p]
    for name, kinds in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_int, *kinds, ctypes.c_void_p]
    _declare_handle_api(lib)
    _lib = lib
    return lib


def _declare_handle_api(lib: ctypes.CDLL) -> None:
    """argtypes of the handle-based entry points (Poisson, IB, fused step); filled in by those modules."""
    for name, (restype, argtypes) in _HANDLE_SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes


_HANDLE_SIGNATURES: dict[str, tuple[Any, list[Any]]] = {
    # sopht_poisson_create(handle*, dtype, dim, nz, ny, nx, x_range, dx, mz, my, mx, origin, flags, stream)
    "sopht_poisson_create": (
        ctypes.c_int,
        [ctypes.POINTER(_P), _I, _I, _I, _I, _I, _D, _D, _PD, _PD, _PD, _D, _I, _P],
    ),
    "sopht_poisson_neumann_create": (
        ctypes.c_int, [ctypes.POINTER(_P), _I, _I, _I, _I, _I, ctypes.c_double, _P]),
    "sopht_poisson_periodic_create": (
        ctypes.c_int, [ctypes.POINTER(_P), _I, _I, _I, _I, _I, ctypes.c_double, _I, _P]),
    "sopht_poisson_solve": (ctypes.c_int, [_P, _F, _F, _P]),
    "sopht_poisson_green_hat": (ctypes.c_int, [_P, ctypes.POINTER(_P)]),
    "sopht_poisson_path": (ctypes.c_char_p, [_P]),
    "sopht_poisson_destroy": (ctypes.c_int, [_P]),
    # z-slab decomposed Poisson solve (local phases)
    "sopht_poisson_slab_create": (
        ctypes.c_int,
        [ctypes.POINTER(_P), _I, _I, _I, _I, _I, _I, _D, _PD, _PD, _PD, _D, _P],
    ),
    "sopht_poisson_slab_create_periodic": (
        ctypes.c_int, [ctypes.POINTER(_P), _I, _I, _I, _I, _I, _I, _D, _I, _P]),
    "sopht_poisson_slab_forward_x": (ctypes.c_int, [_P, _F, _P, _P, _P]),
    "sopht_poisson_slab_pipe_forward_x": (ctypes.c_int, [_P, _F, _I, _P, _P]),
    "sopht_poisson_slab_pipe_transpose": (ctypes.c_int, [_P, _I, _I, ctypes.POINTER(_P), _I, _P]),
    "sopht_poisson_slab_pipe_yz": (ctypes.c_int, [_P, _I, _P, _P, _P]),
    "sopht_poisson_slab_pipe_inverse_x": (ctypes.c_int, [_P, _F, _I, _P, _P]),
    "sopht_poisson_slab_yz": (ctypes.c_int, [_P, _P, _P, _P, _P, _P]),
    "sopht_poisson_slab_inverse_x": (ctypes.c_int, [_P, _F, _P, _P, _P]),
    "sopht_poisson_slab_enable_peer_exchange": (ctypes.c_int, [_P, _P]),
    "sopht_poisson_slab_open_peers": (ctypes.c_int, [_P, _P]),
    "sopht_poisson_slab_destroy": (ctypes.c_int, [_P]),
    # plain rfftn / irfftn plans (FFTPyFFTW{2,3}D)
    "sopht_fft_create": (ctypes.c_int, [ctypes.POINTER(_P), _I, _I, _I, _I, _I]),
    "sopht_fft_forward": (ctypes.c_int, [_P, _F, _F, _P]),
    "sopht_fft_inverse": (ctypes.c_int, [_P, _F, _F, _P]),
    "sopht_fft_destroy": (ctypes.c_int, [_P]),
    # rigid-body forcing grids
    "sopht_rigid_forcing_grid_kinematics": (ctypes.c_int, [_I, _F, _F, _F, _F, _PD, _PD, _PD, _PD, _P]),
    "sopht_rigid_forcing_grid_force_sums": (ctypes.c_int, [_I, _I, _F, _F, _P, _P]),
    "sopht_rod_state_doubles": (ctypes.c_int64, [ctypes.c_int64]),
    "sopht_rod_forcing_grid_kinematics": (ctypes.c_int, [_I, _I, ctypes.c_int64, _P, _F, _F, _F, _P, _P, _P, _P]),
    "sopht_rod_forcing_grid_transfer": (ctypes.c_int, [_I, _I, _I, ctypes.c_int64, _P, _F, _F, _P, _P, _P]),
    # peer-memory arena (halo exchange / barrier over NVLink)
    "sopht_peer_arena_create": (ctypes.c_int, [ctypes.POINTER(_P), ctypes.c_size_t, _I, _I, _P]),
    "sopht_peer_arena_open": (ctypes.c_int, [_P, _P]),
    "sopht_peer_arena_payload": (_P, [_P]),
    "sopht_peer_halo_exchange": (
        ctypes.c_int,
        [_P, _I, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int), _I, _I,
         ctypes.c_int64, _P],
    ),
    "sopht_peer_barrier": (ctypes.c_int, [_P, _P]),
    "sopht_peer_arena_status": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_int)]),
    "sopht_peer_arena_set_periodic": (ctypes.c_int, [_P, _I]),
    "sopht_peer_arena_destroy": (ctypes.c_int, [_P]),
    # fused 3-D Navier-Stokes passes
    "sopht_ns3d_advect_rotational": (ctypes.c_int, [_I, _F, _F, _F, _D, _P]),
    "sopht_ns3d_diffuse": (ctypes.c_int, [_I, _F, _F, _D, _F, _P]),
    "sopht_ns3d_diffuse_penalise": (ctypes.c_int, [_I, _F, _F, _D, _F, _P, _P, _P, _P]),
    "sopht_ns3d_velocity_from_stream_function": (ctypes.c_int, [_I, _F, _F, _D, _PD, _P, _P]),
    "sopht_ns3d_advect_rotational_periodic_xy": (ctypes.c_int, [_I, _F, _F, _F, _D, _P]),
    "sopht_ns3d_diffuse_periodic_xy": (ctypes.c_int, [_I, _F, _F, _D, _F, _P]),
    "sopht_ns3d_velocity_from_stream_function_periodic_xy": (ctypes.c_int, [_I, _F, _F, _D, _PD, _P, _P]),
    # immersed boundary (int dtype, int dim, ...)
    "sopht_ib_local_support": (ctypes.c_int, [_I, _I, _F, _F, _F, _I, _D, _D, _P]),
    "sopht_ib_interpolation_weights": (ctypes.c_int, [_I, _I, _I, _F, _F, _D, _D, _P]),
    "sopht_ib_eulerian_to_lagrangian": (ctypes.c_int, [_I, _I, _F, _F, _F, _F, _D, _P]),
    "sopht_ib_lagrangian_to_eulerian": (ctypes.c_int, [_I, _I, _F, _F, _F, _F, _P]),
    "sopht_ib_spread_stragglers": (ctypes.c_int, [ctypes.POINTER(ctypes.c_ulonglong)]),
    "sopht_ib_virtual_boundary_forcing": (
        ctypes.c_int,
        [_I, _I, _F, _F, _F, _F, _I, _F, _F, _F, _F, _F, _F, _F, _D, _D, _D, _D, _D, _D, _P],
    ),
}


def exported_symbols() -> list[str]:
    """Every symbol include/sopht_b200.h declares (used by the CPU-side ABI test)."""
    return [
        "sopht_last_error",
        "sopht_version",
        "sopht_launch_count",
        "sopht_profile_enable",
        "sopht_profile_report",
        "sopht_profile_range_begin",
        "sopht_profile_range_end",
        *_SIGNATURES.keys(),
        *_HANDLE_SIGNATURES.keys(),
    ]


_replayed_launches = 0  # kernel launches executed through CUDA-graph replays (the library counts at capture time only)


def launch_count() -> int:
    return int(load().sopht_launch_count()) + _replayed_launches


class StepGraph:
    """A sequence of library calls captured once into a CUDA graph and replayed with one launch.

    For the small-grid regime (BASELINE configs[0-2]: a step is 15-40 kernels of 5-40 us each, the host enqueues them
    more slowly than the device runs them): `g = StepGraph(fn, state)`; `g()` replays. `fn` must be a fixed sequence of
    library calls on fixed tensors with fixed scalar arguments (dt, free stream ...: they are baked into the graph) and
    must not synchronise. Capture needs every lazy initialisation (kernel attributes, occupancy queries, FFT plans)
    behind it, so `fn` is run once for real first; the tensors in `state` are saved before that run and restored after
    it, so constructing the graph does not advance the simulation."""

    def __init__(self, fn, state=()) -> None:
        saved = [t.clone() for t in state]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for t, v in zip(state, saved):
            t.copy_(v)
        n0 = int(load().sopht_launch_count())
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            fn()
        self.launches_per_replay = int(load().sopht_launch_count()) - n0
        global _replayed_launches
        _replayed_launches -= self.launches_per_replay  # the capture itself executed nothing
        for t, v in zip(state, saved):
            t.copy_(v)

    def __call__(self) -> None:
        global _replayed_launches
        self.graph.replay()
        _replayed_launches += self.launches_per_replay


_profile_on = False


def profile_enable(on: bool) -> None:
    """Switch the library's per-kernel CUDA-event timers on or off (include/sopht_b200.h)."""
    global _profile_on
    _profile_on = bool(on)
    load().sopht_profile_enable(1 if on else 0)


class profile_range:
    """Times host-enqueued work (collectives) on the current stream with the library's event timers."""

    def __init__(self, label: str) -> None:
        self.label = label.encode()
        self.token = -1

    def __enter__(self) -> "profile_range":
        if _profile_on:
            self.token = load().sopht_profile_range_begin(self.label, current_stream())
        return self

    def __exit__(self, *exc) -> None:
        if self.token >= 0:
            load().sopht_profile_range_end(self.token, current_stream())


def profile_report() -> dict:
    """{label: {"launches": n, "ms": total}} of the launches recorded since the last report."""
    import json

    return json.loads(load().sopht_profile_report().decode())


# -------------------------------------------------------------------------------------------------
_VALUE_ERRORS = {-1, -2, -3, -4}


def check(rc: int) -> None:
    if rc == 0:
        return
    msg = load().sopht_last_error().decode("utf-8", "replace")
    if rc in _VALUE_ERRORS:
        raise ValueError(msg)
    raise SophtLibraryError(f"libsopht_b200 error {rc}: {msg}")


def dtype_code(real_t: Any) -> int:
    """Map the reference's ``real_t`` (np.float32 / np.float64) to the ABI dtype.

    Mirrors sopht/utils/pyst_kernel_config.py:5-12 (raises ValueError("Invalid real type")).
    """
    if real_t == np.float32 or real_t is torch.float32:
        return SOPHT_F32
    if real_t == np.float64 or real_t is torch.float64:
        return SOPHT_F64
    msg = "Invalid real type"
    raise ValueError(msg)


_TORCH_REAL = {SOPHT_F32: torch.float32, SOPHT_F64: torch.float64}
_TORCH_COMPLEX = {SOPHT_F32: torch.complex64, SOPHT_F64: torch.complex128}


def torch_dtype(real_t: Any) -> torch.dtype:
    return _TORCH_REAL[dtype_code(real_t)]


def torch_complex_dtype(real_t: Any) -> torch.dtype:
    return _TORCH_COMPLEX[dtype_code(real_t)]


# descriptors of recently seen views (a simulator hands the same few tensors over every step); read-only by contract
_DESC_CACHE: dict[tuple, SophtField] = {}


def field_desc(t: torch.Tensor, dt: int, *, is_complex: bool = False) -> SophtField:
    """Describe a CUDA tensor view for the C ABI (strides in elements of its own dtype)."""
    try:
        key = (t.data_ptr(), t.shape, t.stride(), t.dtype, dt, is_complex)
        return _DESC_CACHE[key]
    except (KeyError, AttributeError):
        pass
    f = _field_desc_uncached(t, dt, is_complex)
    if len(_DESC_CACHE) > 4096:
        _DESC_CACHE.clear()
    _DESC_CACHE[key] = f
    return f


def _field_desc_uncached(t: torch.Tensor, dt: int, is_complex: bool) -> SophtField:
    if not isinstance(t, torch.Tensor):
        msg = f"expected a torch.Tensor, got {type(t).__name__}"
        raise TypeError(msg)
    if not t.is_cuda:
        msg = "sopht_b200 kernels need CUDA tensors (no CPU fallback)"
        raise SophtLibraryError(msg)
    want = _TORCH_COMPLEX[dt] if is_complex else _TORCH_REAL[dt]
    if t.dtype != want:
        msg = f"field dtype {t.dtype} does not match kernel dtype {want}"
        raise ValueError(msg)
    if t.dim() < 1 or t.dim() > MAX_DIMS:
        msg = f"unsupported field rank {t.dim()}"
        raise ValueError(msg)
    f = SophtField()
    f.data = t.data_ptr()
    f.ndim = t.dim()
    for d in range(t.dim()):
        f.shape[d] = t.shape[d]
        f.stride[d] = t.stride(d)
    return f


def raw_desc(t: torch.Tensor) -> SophtField:
    """Describe a CUDA tensor of any dtype (int64 index arrays, float64 Lagrangian positions)."""
    if not isinstance(t, torch.Tensor):
        msg = f"expected a torch.Tensor, got {type(t).__name__}"
        raise TypeError(msg)
    if not t.is_cuda:
        msg = "sopht_b200 kernels need CUDA tensors (no CPU fallback)"
        raise SophtLibraryError(msg)
    if t.dim() < 1 or t.dim() > MAX_DIMS:
        msg = f"unsupported field rank {t.dim()}"
        raise ValueError(msg)
    f = SophtField()
    f.data = t.data_ptr()
    f.ndim = t.dim()
    for d in range(t.dim()):
        f.shape[d] = t.shape[d]
        f.stride[d] = t.stride(d)
    return f


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def current_stream() -> ctypes.c_void_p:
    """torch's current CUDA stream on the current device as a cudaStream_t."""
    if _raw_stream is not None:  # one C call instead of building a torch.cuda.Stream object
        return ctypes.c_void_p(_raw_stream(torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def double_array(vals: Any, n: int | None = None) -> Any:
    seq = [float(v) for v in (vals.tolist() if hasattr(vals, "tolist") else vals)]
    if n is not None and len(seq) < n:
        msg = f"expected at least {n} values, got {len(seq)}"
        raise ValueError(msg)
    return (ctypes.c_double * len(seq))(*seq)


def call(name: str, dt: int, *args: Any) -> None:
    """Invoke ``name(dtype, *args, stream)``; tensors -> descriptors, numbers -> C scalars."""
    lib = load()
    kinds = _SIGNATURES[name]
    if len(args) != len(kinds):
        msg = f"{name}: expected {len(kinds)} arguments, got {len(args)}"
        raise TypeError(msg)
    cargs: list[Any] = []
    keep: list[Any] = []
    is_complex = name == "sopht_elementwise_complex_product"
    for a, k in zip(args, kinds):
        if k is _F:
            d = field_desc(a, dt, is_complex=is_complex)
            keep.append(d)
            cargs.append(ctypes.byref(d))
        elif k is _D:
            cargs.append(ctypes.c_double(float(a)))
        elif k is _I:
            cargs.append(ctypes.c_int(int(a)))
        elif k is _PD:
            arr = a if isinstance(a, ctypes.Array) else double_array(a)
            keep.append(arr)
            cargs.append(arr)
        elif k is _P:
            cargs.append(ctypes.c_void_p(int(a)))
        else:  # pragma: no cover
            raise TypeError(k)
    check(getattr(lib, name)(dt, *cargs, current_stream()))


# -------------------------------------------------------------------------------------------------
class Staging:
    """Lets the reference-shaped callables take numpy arrays as well as CUDA tensors.

    numpy arguments are copied to the device on entry and outputs copied back in place on exit, so the
    reference's own call sites and tests (which own numpy arrays) work unchanged. The same numpy view
    passed twice (``sum_field=field, field_1=field``) maps to one device tensor, preserving aliasing.
    CUDA tensors pass straight through (zero copy).
    """

    def __init__(self) -> None:
        self._map: dict[tuple, torch.Tensor] = {}
        self._arrays: list[tuple[tuple, np.ndarray, bool]] = []  # (key, array, is_output) of every staged view
        self._writeback: list[tuple[np.ndarray, torch.Tensor]] = []

    def __enter__(self) -> "Staging":
        return self

    def __exit__(self, exc_type, exc, tb) -> None:
        if exc_type is None:
            for arr, t in self._writeback:
                arr[...] = t.cpu().numpy()
        self._map.clear()
        self._arrays.clear()
        self._writeback.clear()

    def _check_overlap(self, key: tuple, a: np.ndarray, out: bool) -> None:
        """Two DIFFERENT views of one host buffer (a field and its slice, a vector field and one component) become
        independent device copies, and the later write-back would silently clobber the earlier one - the reference's
        in-place numpy semantics cannot be kept, so refuse instead of returning wrong data."""
        for k2, b, out2 in self._arrays:
            if k2 != key and (out or out2) and np.shares_memory(a, b):
                msg = ("numpy arguments that partially overlap (different views of one buffer, one of them an output) "
                       "cannot be staged through the device; pass CUDA tensors (views alias there) or disjoint arrays")
                raise ValueError(msg)
        self._arrays.append((key, a, out))

    def _get(self, a: Any, out: bool) -> torch.Tensor:
        if isinstance(a, torch.Tensor):
            return a
        if isinstance(a, np.ndarray):
            key = (a.__array_interface__["data"][0], a.shape, a.strides, a.dtype.str)
            self._check_overlap(key, a, out)
            t = self._map.get(key)
            if t is None:
                if not torch.cuda.is_available():
                    msg = "sopht_b200 kernels need a CUDA device (no CPU fallback)"
                    raise SophtLibraryError(msg)
                t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
                self._map[key] = t
                if out:
                    self._writeback.append((a, t))
            elif out and not any(w[0] is a for w in self._writeback):
                self._writeback.append((a, t))
            return t
        msg = f"expected a torch.Tensor or numpy.ndarray, got {type(a).__name__}"
        raise TypeError(msg)

    def inp(self, a: Any) -> torch.Tensor:
        return self._get(a, out=False)

    def out(self, a: Any) -> torch.Tensor:
        return self._get(a, out=True)
