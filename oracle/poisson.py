"""CPU ORACLE (test infrastructure): unbounded Poisson solve, scipy.fft restatement.

Follows sopht/numeric/eulerian_grid_ops/poisson_solver_3d/UnboundedPoissonSolverPYFFTW3D.py:51-149 and
poisson_solver_2d/UnboundedPoissonSolverPYFFTW2D.py:47-129 with scipy's rfftn/irfftn standing in for
the two pyFFTW plans — exactly the substitution the reference's own tests make
(tests/.../test_unbounded_poisson_solver_3d.py:55-90, SophT's scipy_fft_3d.py:7-15).
"""

from __future__ import annotations

import numpy as np
from scipy.fft import irfftn, rfftn


class UnboundedPoissonSolver3D:
    def __init__(self, grid_size_z, grid_size_y, grid_size_x, x_range=1.0, real_t=np.float64, workers=1,
                 kernels=None):
        self._k = kernels  # None: numpy slice arithmetic; oracle.cstencils: the reference's elementwise kernel passes
        self.nz, self.ny, self.nx = grid_size_z, grid_size_y, grid_size_x
        self.real_t = real_t
        self.workers = workers
        self.x_range = x_range
        self.y_range = x_range * (grid_size_y / grid_size_x)
        self.z_range = x_range * (grid_size_z / grid_size_x)
        self.dx = real_t(x_range / grid_size_x)
        self.domain_doubled_buffer = np.zeros((2 * self.nz, 2 * self.ny, 2 * self.nx), dtype=real_t)
        self.fourier_greens_function_times_dx_cubed = self._greens() * (self.dx**3)

    def _greens(self):  # UnboundedPoissonSolverPYFFTW3D.py:51-83
        real_t = self.real_t
        x = np.linspace(0, 2 * self.x_range - self.dx, 2 * self.nx).astype(real_t)
        y = np.linspace(0, 2 * self.y_range - self.dx, 2 * self.ny).astype(real_t)
        z = np.linspace(0, 2 * self.z_range - self.dx, 2 * self.nz).astype(real_t)
        zg, yg, xg = np.meshgrid(z, y, x, indexing="ij")
        r = np.sqrt(
            np.minimum(xg, 2 * self.x_range - xg) ** 2
            + np.minimum(yg, 2 * self.y_range - yg) ** 2
            + np.minimum(zg, 2 * self.z_range - zg) ** 2
        )
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            g = (1 / r) / (4 * np.pi)
        g[0, 0, 0] = 1 / (4 * np.pi * self.dx)
        return rfftn(g, workers=self.workers)

    def solve(self, solution_field, rhs_field):  # :111-149
        buf = self.domain_doubled_buffer
        corner = (slice(None, self.nz), slice(None, self.ny), slice(None, self.nx))
        if self._k is not None:  # the reference's own passes: set_fixed_val, copy, rfft, complex product, irfft, copy
            k = self._k
            k.set_fixed_val(buf, 0)
            k.elementwise_copy(buf[corner], rhs_field)
            spec = rfftn(buf, workers=self.workers)
            k.elementwise_complex_product(spec, spec, self.fourier_greens_function_times_dx_cubed)
            back = irfftn(spec, s=buf.shape, workers=self.workers, overwrite_x=True)
            k.elementwise_copy(solution_field, back[corner])
            return
        buf[...] = 0
        buf[corner] = rhs_field
        spec = rfftn(buf, workers=self.workers)
        spec = spec * self.fourier_greens_function_times_dx_cubed
        solution_field[...] = irfftn(spec, s=buf.shape, workers=self.workers)[corner]

    def vector_field_solve(self, solution_vector_field, rhs_vector_field):  # :151-172
        for c in range(3):
            self.solve(solution_vector_field[c], rhs_vector_field[c])


class UnboundedPoissonSolver2D:
    def __init__(self, grid_size_y, grid_size_x, x_range=1.0, real_t=np.float64, workers=1, kernels=None):
        self._k = kernels
        self.ny, self.nx = grid_size_y, grid_size_x
        self.real_t = real_t
        self.workers = workers
        self.x_range = x_range
        self.y_range = x_range * (grid_size_y / grid_size_x)
        self.dx = real_t(x_range / grid_size_x)
        self.domain_doubled_buffer = np.zeros((2 * self.ny, 2 * self.nx), dtype=real_t)
        self.fourier_greens_function_times_dx_squared = self._greens() * (self.dx**2)

    def _greens(self):  # UnboundedPoissonSolverPYFFTW2D.py:47-68
        real_t = self.real_t
        x = np.linspace(0, 2 * self.x_range - self.dx, 2 * self.nx).astype(real_t)
        y = np.linspace(0, 2 * self.y_range - self.dx, 2 * self.ny).astype(real_t)
        xg, yg = np.meshgrid(x, y)
        r = np.sqrt(
            np.minimum(xg, 2 * self.x_range - xg) ** 2 + np.minimum(yg, 2 * self.y_range - yg) ** 2
        )
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            g = -np.log(r) / (2 * np.pi)
        g[0, 0] = -(2 * np.log(self.dx / np.sqrt(np.pi)) - 1) / (4 * np.pi)
        return rfftn(g, workers=self.workers)

    def solve(self, solution_field, rhs_field):  # :95-129
        buf = self.domain_doubled_buffer
        corner = (slice(None, self.ny), slice(None, self.nx))
        if self._k is not None:
            k = self._k
            k.set_fixed_val(buf, 0)
            k.elementwise_copy(buf[corner], rhs_field)
            spec = rfftn(buf, workers=self.workers)
            k.elementwise_complex_product(spec, spec, self.fourier_greens_function_times_dx_squared)
            back = irfftn(spec, s=buf.shape, workers=self.workers, overwrite_x=True)
            k.elementwise_copy(solution_field, back[corner])
            return
        buf[...] = 0
        buf[corner] = rhs_field
        spec = rfftn(buf, workers=self.workers)
        spec = spec * self.fourier_greens_function_times_dx_squared
        solution_field[...] = irfftn(spec, s=buf.shape, workers=self.workers)[corner]


class FastDiagPoissonSolver:
    """Restatement of FastDiagPoissonSolver{2,3}D (poisson_solver_3d/FastDiagPoissonSolver3D.py:15-181,
    poisson_solver_2d/FastDiagPoissonSolver2D.py:13-119), homogeneous Neumann walls: per axis the matrix
    tridiag(-1, 2, -1) / dx^2 with both corner entries 1 / dx^2 (:51-100), its eigen-decomposition sorted by decreasing
    eigenvalue (:108-133), the mean mode's eigenvalue set to inf (:143-146), then forward transform with V^-1 along
    every axis, division by the eigenvalue sums, backward transform with V (:150-181). Pinned against the reference
    classes themselves, run in the build container (tests/golden/fastdiag_*.npz)."""

    def __init__(self, grid_size, dx, real_t=np.float64):
        self.grid_size = tuple(grid_size)  # (nz, ny, nx) or (ny, nx)
        self.real_t = real_t
        inv_dx2 = real_t(1 / dx / dx)
        self.eig_vecs, self.inv_eig_vecs, vals = [], [], []
        for n in self.grid_size:
            mat = (2.0 * np.eye(n) - np.eye(n, k=1) - np.eye(n, k=-1)).astype(real_t) * inv_dx2
            mat[0, 0] = inv_dx2
            mat[-1, -1] = inv_dx2
            w, v = np.linalg.eig(mat)
            order = w.argsort()[::-1]
            self.eig_vecs.append(v[:, order])
            self.inv_eig_vecs.append(np.linalg.inv(v[:, order]))
            vals.append(w[order])
        dim = len(self.grid_size)
        total = sum(w.reshape([-1 if a == d else 1 for a in range(dim)]) for d, w in enumerate(vals))
        total = np.array(np.broadcast_to(total, self.grid_size))
        total[(-1,) * dim] = np.inf
        self.inv_eig_val_matrix = (real_t(1) / total).astype(real_t)

    def _apply(self, mats, field):
        for axis, m in enumerate(mats):
            field = np.moveaxis(np.tensordot(m, field, axes=(1, axis)), 0, axis)
        return field

    def solve(self, solution_field, rhs_field):
        spectral = self._apply(self.inv_eig_vecs, rhs_field) * self.inv_eig_val_matrix
        solution_field[...] = self._apply(self.eig_vecs, spectral)

    def vector_field_solve(self, solution_vector_field, rhs_vector_field):
        for c in range(len(self.grid_size)):
            self.solve(solution_vector_field[c], rhs_vector_field[c])


class PeriodicPoissonSolver:
    """numpy restatement of the periodic Poisson solve (-del^2 psi = rhs, zero-mean solution) - an EXTENSION for
    BASELINE config 4. **Parity unpinned**: the reference has no periodic solver and no test of one (SURVEY.md fact
    2); this class is pinned on analytic Fourier modes only (tests/test_periodic_poisson.py)."""

    def __init__(self, grid_size, dx, symbol="spectral"):
        self.grid_size = tuple(grid_size)
        dim = len(self.grid_size)
        lam = np.zeros(self.grid_size)
        for axis, n in enumerate(self.grid_size):
            k = np.arange(n)
            if symbol == "spectral":
                m = np.where(k <= n // 2, k, k - n)
                sym = (2.0 * np.pi * m / (n * dx)) ** 2
            else:
                sym = 4.0 * np.sin(np.pi * k / n) ** 2 / (dx * dx)
            lam = lam + sym.reshape([-1 if a == axis else 1 for a in range(dim)])
        lam[(0,) * dim] = np.inf
        self.inv_symbol = 1.0 / lam

    def solve(self, solution_field, rhs_field):
        solution_field[...] = np.fft.ifftn(np.fft.fftn(rhs_field.astype(np.float64)) * self.inv_symbol).real
