"""Periodic 3-D Navier-Stokes flow simulator - an EXTENSION for BASELINE config 4 (periodic Taylor-Green vortex).

The reference has no periodic case, so there is no reference class to mirror and **parity is
unpinned**; attribute names (`vorticity_field`, `velocity_field`, `stream_func_field`, `position_field`, `dx`,
`time`) and the method names follow UnboundedNavierStokesFlowSimulator3D. The sub-steps and their order are the
reference's unbounded step (navier_stokes_flow_simulators.py:449-485) without the boundary penalisation.

`step_mode="fused"` (default when rows are 16-byte multiples): three register-marching kernels whose x / y neighbour
loads wrap around inside the kernel (csrc/fused_step3d.cu, `*_periodic_xy`), on arrays that carry one halo plane per
z side - the z wrap is a copy of two planes (one GPU) or the neighbour rank's planes (slab decomposition), so the same
kernels serve both: advect (w + p curl(u x w), cross product never stored) -> diffuse -> periodic Poisson solve ->
velocity = curl(psi) with the max-reduction for dt. 180 B/cell/step algorithmic.
`step_mode="unfused"`: the first, pass-by-pass version (public factory kernels on one-cell halo-padded arrays whose
halos are refilled by wrap-around before every stencil), kept as the independent composition the fused step is
tested against.
"""

from __future__ import annotations

import ctypes

import numpy as np
import torch

import sopht_b200.numeric.eulerian_grid_ops as spne
from sopht_b200 import _lib

from .navier_stokes_flow_simulators import compute_advection_diffusion_stable_timestep, stable_timestep_from_max


def wrap_halos(field: torch.Tensor) -> None:
    """Periodic ghost cells of a halo-1 padded (..., nz+2, ny+2, nx+2) tensor, axis by axis so edges and corners
    follow (six strided device copies)."""
    for axis in (-3, -2, -1):
        f = field.movedim(axis, 0)
        f[0] = f[-2]
        f[-1] = f[1]


def wrap_z_halos(*fields: torch.Tensor) -> None:
    """z halo planes of (C, nz + 2, ny, nx) arrays from the opposite owned planes (periodic box on one GPU)."""
    for f in fields:
        _lib.call("sopht_wrap_z_halos", _lib.dtype_code(f.dtype), f)


class PeriodicNavierStokesFlowSimulator3D:
    def __new__(cls, grid_size, *args, step_mode: str = "auto", **kwargs):
        if cls is PeriodicNavierStokesFlowSimulator3D:
            real_t = kwargs.get("real_t", np.float32)
            row_ok = (grid_size[2] * np.dtype(real_t).itemsize) % 16 == 0
            if step_mode not in ("auto", "fused", "unfused"):
                msg = "step_mode must be 'auto', 'fused' or 'unfused'"
                raise ValueError(msg)
            if step_mode == "fused" and not row_ok:
                msg = "the fused periodic step needs rows that are multiples of 16 bytes"
                raise ValueError(msg)
            if step_mode == "unfused" or not row_ok:
                return object.__new__(_PassByPassPeriodicSimulator3D)
        return object.__new__(cls)

    def __init__(self, grid_size: tuple[int, int, int], x_range: float, kinematic_viscosity: float, cfl: float = 0.1,
                 real_t: type = np.float32, num_threads: int = 1, time: float = 0.0,
                 poisson_symbol: str = "spectral", step_mode: str = "auto") -> None:
        if not torch.cuda.is_available():
            msg = "sopht_b200 flow simulators need a CUDA device (no CPU fallback)"
            raise _lib.SophtLibraryError(msg)
        self.grid_dim = 3
        self.grid_size = tuple(grid_size)
        self.x_range, self.real_t, self.kinematic_viscosity, self.cfl, self.time = (
            x_range, real_t, kinematic_viscosity, cfl, time)
        self.step_mode = "fused"
        nz, ny, nx = self.grid_size
        self.dx = real_t(x_range / nx)
        self.device = torch.device("cuda", torch.cuda.current_device())
        tt = _lib.torch_dtype(real_t)
        self._dc = _lib.dtype_code(real_t)
        coords = [np.linspace(self.dx / 2.0, x_range * n / nx - self.dx / 2.0, n).astype(real_t) for n in (nz, ny, nx)]
        mesh = np.flipud(np.array(np.meshgrid(*coords, indexing="ij")))
        self.position_field = torch.from_numpy(np.ascontiguousarray(mesh)).to(self.device)
        padded = (3, nz + 2, ny, nx)  # one halo plane per z side; x and y wrap inside the kernels
        self._w = torch.zeros(padded, dtype=tt, device=self.device)
        self._u, self._buf, self._psi = (torch.zeros_like(self._w) for _ in range(3))
        self.vorticity_field = self._w[:, 1:-1]
        self.velocity_field = self._u[:, 1:-1]
        self.stream_func_field = self._psi[:, 1:-1]
        self._poisson = spne.PeriodicPoissonSolver3D(nz, ny, nx, x_range=x_range, real_t=real_t,
                                                     symbol=poisson_symbol)
        self._vel_absmax = torch.zeros(1, dtype=tt, device=self.device)
        self._vel_absmax_version = None

    def compute_velocity_from_vorticity(self) -> None:
        lib, fd, st = _lib.load(), _lib.field_desc, _lib.current_stream()
        self._poisson.vector_field_solve(solution_vector_field=self.stream_func_field,
                                         rhs_vector_field=self.vorticity_field)
        wrap_z_halos(self._psi)
        fu, fpsi = fd(self._u, self._dc), fd(self._psi, self._dc)
        _lib.check(lib.sopht_ns3d_velocity_from_stream_function_periodic_xy(
            self._dc, ctypes.byref(fu), ctypes.byref(fpsi), float(self.real_t(0.5 / self.dx)), None,
            ctypes.c_void_p(self._vel_absmax.data_ptr()), st))
        self._vel_absmax_version = (self._u.data_ptr(), self._u._version)

    def time_step(self, dt: float) -> None:
        rt = self.real_t
        lib, fd, st = _lib.load(), _lib.field_desc, _lib.current_stream()
        wrap_z_halos(self._w, self._u)
        fw, fu, fb = fd(self._w, self._dc), fd(self._u, self._dc), fd(self._buf, self._dc)
        _lib.check(lib.sopht_ns3d_advect_rotational_periodic_xy(
            self._dc, ctypes.byref(fb), ctypes.byref(fw), ctypes.byref(fu), float(rt(dt / (2 * self.dx))), st))
        wrap_z_halos(self._buf)
        _lib.check(lib.sopht_ns3d_diffuse_periodic_xy(
            self._dc, ctypes.byref(fw), ctypes.byref(fb),
            float(rt(self.kinematic_viscosity * dt / self.dx / self.dx)), None, st))
        self.compute_velocity_from_vorticity()
        self.time += dt

    def compute_stable_timestep(self, dt_prefac: float = 1.0) -> float:
        if self._vel_absmax_version == (self._u.data_ptr(), self._u._version):
            dt = stable_timestep_from_max(self.real_t(self._vel_absmax.item()), 3, self.dx, self.cfl,
                                          self.kinematic_viscosity, self.real_t)
            return dt * dt_prefac
        dt = compute_advection_diffusion_stable_timestep(
            velocity_field=self.velocity_field, velocity_magnitude_field=self._buf[0, 1:-1], grid_dim=3, dx=self.dx,
            cfl=self.cfl, kinematic_viscosity=self.kinematic_viscosity, real_t=self.real_t)
        return dt * dt_prefac


class _PassByPassPeriodicSimulator3D(PeriodicNavierStokesFlowSimulator3D):
    """`step_mode="unfused"`: public factory kernels on arrays padded by one cell in every direction."""

    def __init__(self, grid_size: tuple[int, int, int], x_range: float, kinematic_viscosity: float, cfl: float = 0.1,
                 real_t: type = np.float32, num_threads: int = 1, time: float = 0.0,
                 poisson_symbol: str = "spectral", step_mode: str = "unfused") -> None:
        self.step_mode = "unfused"
        if not torch.cuda.is_available():
            msg = "sopht_b200 flow simulators need a CUDA device (no CPU fallback)"
            raise _lib.SophtLibraryError(msg)
        self.grid_dim = 3
        self.grid_size = tuple(grid_size)
        self.x_range = x_range
        self.real_t = real_t
        self.kinematic_viscosity = kinematic_viscosity
        self.cfl = cfl
        self.time = time
        nz, ny, nx = self.grid_size
        self.dx = real_t(x_range / nx)
        self.device = torch.device("cuda", torch.cuda.current_device())
        tt = _lib.torch_dtype(real_t)
        coords = [np.linspace(self.dx / 2.0, x_range * n / nx - self.dx / 2.0, n).astype(real_t) for n in (nz, ny, nx)]
        mesh = np.flipud(np.array(np.meshgrid(*coords, indexing="ij")))
        self.position_field = torch.from_numpy(np.ascontiguousarray(mesh)).to(self.device)
        padded = (3, nz + 2, ny + 2, nx + 2)
        self._w = torch.zeros(padded, dtype=tt, device=self.device)
        self._u, self._buf, self._psi = (torch.zeros_like(self._w) for _ in range(3))
        self.vorticity_field = self._w[:, 1:-1, 1:-1, 1:-1]
        self.velocity_field = self._u[:, 1:-1, 1:-1, 1:-1]
        self.stream_func_field = self._psi[:, 1:-1, 1:-1, 1:-1]
        gs = padded[1:]
        self._cross = spne.gen_elementwise_cross_product_pyst_kernel_3d(
            real_t=real_t, num_threads=num_threads, fixed_grid_size=gs)
        self._advect = spne.gen_update_vorticity_from_velocity_forcing_pyst_kernel_3d(
            real_t=real_t, fixed_grid_size=gs, num_threads=num_threads)
        self._diffuse = spne.gen_diffusion_timestep_euler_forward_pyst_kernel_3d(
            real_t=real_t, fixed_grid_size=gs, num_threads=num_threads, field_type="vector")
        self._curl = spne.gen_curl_pyst_kernel_3d(real_t=real_t, num_threads=num_threads, fixed_grid_size=gs)
        self._poisson = spne.PeriodicPoissonSolver3D(nz, ny, nx, x_range=x_range, real_t=real_t,
                                                     symbol=poisson_symbol)

    def compute_velocity_from_vorticity(self) -> None:
        self._poisson.vector_field_solve(solution_vector_field=self.stream_func_field,
                                         rhs_vector_field=self.vorticity_field)
        wrap_halos(self._psi)
        self._curl(curl=self._u, field=self._psi, prefactor=self.real_t(0.5 / self.dx))
        wrap_halos(self._u)

    def time_step(self, dt: float) -> None:
        rt = self.real_t
        wrap_halos(self._w)
        wrap_halos(self._u)
        self._cross(result_field=self._buf, field_1=self._u, field_2=self._w)
        self._advect(vorticity_field=self._w, velocity_forcing_field=self._buf, prefactor=rt(dt / (2 * self.dx)))
        wrap_halos(self._w)
        self._diffuse(vector_field=self._w, diffusion_flux=self._buf[0],
                      nu_dt_by_dx2=rt(self.kinematic_viscosity * dt / self.dx / self.dx))
        self.compute_velocity_from_vorticity()
        self.time += dt

    def compute_stable_timestep(self, dt_prefac: float = 1.0) -> float:
        dt = compute_advection_diffusion_stable_timestep(
            velocity_field=self._u, velocity_magnitude_field=self._buf[0], grid_dim=3, dx=self.dx, cfl=self.cfl,
            kinematic_viscosity=self.kinematic_viscosity, real_t=self.real_t)
        return dt * dt_prefac
