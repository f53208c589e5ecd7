"""FFTPyFFTW{2,3}D and the scipy-named rfftn / irfftn helpers of the reference, on the device.

ref: sopht/numeric/eulerian_grid_ops/poisson_solver_3d/FFTPyFFTW3D.py:7-67, poisson_solver_2d/FFTPyFFTW2D.py:7-65
(buffers, `fft_plan` / `ifft_plan` called as `plan(input_array=, output_array=)`, normalised inverse that destroys its
input), poisson_solver_{2,3}d/scipy_fft_{2,3}d.py (the helper the reference ships as its own cross-check). The
transforms run through libsopht_b200 (`sopht_fft_*`: cuFFT batched 2-D + strided 1-D plans, normalisation kernel);
arguments are torch CUDA tensors, numpy arrays are staged through the device like everywhere else in this package.
"""

from __future__ import annotations

import ctypes
from typing import Any

import numpy as np
import torch

from sopht_b200 import _lib


class _Plan:
    """One direction of an FFT handle, callable like a pyfftw.FFTW object."""

    def __init__(self, owner: "_FFTBase", forward: bool) -> None:
        self._owner, self._forward = owner, forward

    def __call__(self, input_array: Any = None, output_array: Any = None) -> Any:
        o = self._owner
        src = o.field_pyfftw_buffer if self._forward else o.fourier_field_pyfftw_buffer
        dst = o.fourier_field_pyfftw_buffer if self._forward else o.field_pyfftw_buffer
        input_array = src if input_array is None else input_array
        output_array = dst if output_array is None else output_array
        with _lib.Staging() as stage:
            # the inverse destroys its input: a numpy spectrum is written back too, as pyfftw would leave it
            a = stage.inp(input_array) if self._forward else stage.out(input_array)
            b = stage.out(output_array)
            if not a.is_contiguous() or not b.is_contiguous():
                msg = "FFT plans take contiguous arrays of the plan's shape (as pyfftw.FFTW does)"
                raise ValueError(msg)
            lib = _lib.load()
            if self._forward:
                fr, fc = _lib.field_desc(a, o._dt), _lib.field_desc(b, o._dt, is_complex=True)
                _lib.check(lib.sopht_fft_forward(o._handle, ctypes.byref(fr), ctypes.byref(fc), _lib.current_stream()))
            else:
                fc, fr = _lib.field_desc(a, o._dt, is_complex=True), _lib.field_desc(b, o._dt)
                _lib.check(lib.sopht_fft_inverse(o._handle, ctypes.byref(fc), ctypes.byref(fr), _lib.current_stream()))
        return output_array


class _FFTBase:
    _dim: int

    def _create(self, shape: tuple[int, ...], num_threads: int, real_t: type) -> None:
        self.num_threads = num_threads  # accepted and ignored: parallelism is the device's
        self.real_t = real_t
        self.complex_dtype = np.complex64 if real_t == np.float32 else np.complex128
        self._dt = _lib.dtype_code(real_t)
        if not torch.cuda.is_available():
            msg = "sopht_b200 FFT plans need a CUDA device (no CPU fallback)"
            raise _lib.SophtLibraryError(msg)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.field_pyfftw_buffer = torch.zeros(shape, dtype=_lib.torch_dtype(real_t), device=dev)
        self.fourier_field_pyfftw_buffer = torch.zeros(
            (*shape[:-1], shape[-1] // 2 + 1), dtype=_lib.torch_complex_dtype(real_t), device=dev)
        handle = ctypes.c_void_p()
        nz = shape[0] if self._dim == 3 else 1
        _lib.check(_lib.load().sopht_fft_create(ctypes.byref(handle), self._dt, self._dim, nz, shape[-2], shape[-1]))
        self._handle = handle
        self.fft_plan = _Plan(self, forward=True)
        self.ifft_plan = _Plan(self, forward=False)

    # the reference's pyfftw.builders objects (never called inside SophT): out-of-place rfftn / normalised irfftn
    def pyfftw_fftn(self, field: Any) -> Any:
        out = torch.zeros_like(self.fourier_field_pyfftw_buffer)
        self.fft_plan(input_array=field, output_array=out)
        return out

    def pyfftw_ifftn(self, fourier_field: Any) -> Any:
        out = torch.zeros_like(self.field_pyfftw_buffer)
        scratch = fourier_field.clone() if isinstance(fourier_field, torch.Tensor) else np.array(fourier_field)
        self.ifft_plan(input_array=scratch, output_array=out)
        return out

    def fft_ifft_plan_kernel(self, fourier_field: Any, inv_fourier_field: Any, field: Any) -> None:
        """Forward and backward transform (FFTPyFFTW3D.py:57-67; only used to benchmark the pair)."""
        self.fft_plan(input_array=field, output_array=fourier_field)
        self.ifft_plan(input_array=fourier_field, output_array=inv_fourier_field)

    def __del__(self) -> None:
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                _lib.load().sopht_fft_destroy(h)
            except Exception:  # noqa: BLE001 - interpreter shutdown
                pass
            self._handle = None


class FFTPyFFTW3D(_FFTBase):
    """FFTPyFFTW3D.py:7-67."""

    _dim = 3

    def __init__(self, grid_size_z: int, grid_size_y: int, grid_size_x: int, num_threads: int = 1,
                 real_t: type = np.float64) -> None:
        self.grid_size_z, self.grid_size_y, self.grid_size_x = grid_size_z, grid_size_y, grid_size_x
        self._create((grid_size_z, grid_size_y, grid_size_x), num_threads, real_t)


class FFTPyFFTW2D(_FFTBase):
    """FFTPyFFTW2D.py:7-65."""

    _dim = 2

    def __init__(self, grid_size_y: int, grid_size_x: int, num_threads: int = 1, real_t: type = np.float64) -> None:
        self.grid_size_y, self.grid_size_x = grid_size_y, grid_size_x
        self._create((grid_size_y, grid_size_x), num_threads, real_t)


def _fft_ifft(dim: int, fourier_field: Any, inv_fourier_field: Any, field: Any) -> None:
    shape = tuple(field.shape)
    real_t = np.float32 if str(field.dtype).endswith("float32") else np.float64
    plan = (FFTPyFFTW3D if dim == 3 else FFTPyFFTW2D)(*shape, real_t=real_t)
    plan.fft_plan(input_array=field, output_array=fourier_field)
    # the inverse consumes its input: work on a copy so that fourier_field keeps the spectrum, like scipy's irfftn
    scratch = fourier_field.clone() if isinstance(fourier_field, torch.Tensor) else np.array(fourier_field)
    plan.ifft_plan(input_array=scratch, output_array=inv_fourier_field)


def fft_ifft_via_scipy_kernel_3d(fourier_field: Any, inv_fourier_field: Any, field: Any, num_threads: int = 1) -> None:
    """fourier_field = rfftn(field); inv_fourier_field = irfftn(fourier_field)  (scipy_fft_3d.py:7-15)."""
    _fft_ifft(3, fourier_field, inv_fourier_field, field)


def fft_ifft_via_scipy_kernel_2d(fourier_field: Any, inv_fourier_field: Any, field: Any, num_threads: int = 1) -> None:
    """2-D twin (poisson_solver_2d/scipy_fft_2d.py)."""
    _fft_ifft(2, fourier_field, inv_fourier_field, field)
