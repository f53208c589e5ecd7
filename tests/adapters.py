"""Two implementations of one kernel vocabulary, so every parity case runs unchanged against
(a) the CPU oracle and (b) the CUDA path behind the reference-shaped factories (through the C ABI).

All methods take numpy arrays and write outputs in place.
"""

from __future__ import annotations

import numpy as np


class OracleOps:
    name = "oracle"

    def __init__(self, real_t):
        from oracle import ib, poisson, stencils

        self.real_t = real_t
        self.s = stencils
        self.p = poisson
        self.ib = ib

    # elementwise
    def elementwise_sum(self, o, a, b):
        self.s.elementwise_sum(o, a, b)

    def set_fixed_val(self, f, v):
        self.s.set_fixed_val(f, v)

    def set_fixed_val_vector(self, f, vals):
        self.s.set_fixed_val_vector(f, vals)

    def elementwise_copy(self, f, r):
        self.s.elementwise_copy(f, r)

    def complex_product(self, o, a, b):
        self.s.elementwise_complex_product(o, a, b)

    def set_fixed_val_at_boundaries(self, f, width, v):
        self.s.set_fixed_val_at_boundaries(f, width, v)

    def set_fixed_val_at_boundaries_vector(self, f, width, vals):
        self.s.set_fixed_val_at_boundaries_vector(f, width, vals)

    def add_fixed_val(self, o, f, v):
        self.s.add_fixed_val(o, f, v)

    def add_fixed_val_vector(self, o, f, vals):
        self.s.add_fixed_val_vector(o, f, vals)

    def saxpby(self, o, a, b, pa, pb):
        self.s.elementwise_saxpby(o, a, b, pa, pb)

    def cross_product(self, o, a, b):
        self.s.elementwise_cross_product(o, a, b)

    # stencils (dim taken from the arrays)
    def diffusion_flux(self, flux, field, p, reset=True, vector=False):
        (self.s.diffusion_flux_vector if vector else self.s.diffusion_flux)(flux, field, p, reset)

    def diffusion_timestep(self, field, flux, p, vector=False):
        if vector:
            self.s.diffusion_timestep_euler_forward_vector(field, flux, p)
        else:
            self.s.diffusion_timestep_euler_forward(field, flux, p)

    def curl_3d(self, curl, field, p, reset=True):
        self.s.curl_3d(curl, field, p, reset)

    def divergence_3d(self, div, field, inv_dx, reset=True):
        self.s.divergence_3d(div, field, inv_dx, reset)

    def forcing_update(self, w, f, p):
        if w.ndim == 2:
            self.s.update_vorticity_from_velocity_forcing_2d(w, f, p)
        else:
            self.s.update_vorticity_from_velocity_forcing_3d(w, f, p)

    def penalised_velocity_update(self, w, pv, u, p):
        if w.ndim == 2:
            self.s.update_vorticity_from_penalised_velocity_2d(w, pv, u, p)
        else:
            self.s.update_vorticity_from_penalised_velocity_3d(w, pv, u, p)

    def stretching_flux(self, q, w, u, p):
        self.s.vorticity_stretching_flux_3d(q, w, u, p)

    def stretching_timestep(self, w, u, q, p, stepper="euler_forward", midstep=None):
        if stepper == "euler_forward":
            self.s.vorticity_stretching_timestep_euler_forward_3d(w, u, q, p)
        else:
            self.s.vorticity_stretching_timestep_ssprk3_3d(w, u, q, p, midstep)

    def advection_flux(self, q, f, v, inv_dx):
        self.s.advection_flux_conservative_eno3(q, f, v, inv_dx)

    def advection_timestep(self, f, q, v, dt_by_dx, vector=False):
        if vector:
            self.s.advection_timestep_euler_forward_conservative_eno3_vector(f, q, v, dt_by_dx)
        else:
            self.s.advection_timestep_euler_forward_conservative_eno3(f, q, v, dt_by_dx)

    def penalise(self, f, width, dx, grids, vector=False):
        """grids: full coordinate arrays ordered (x_grid, y_grid[, z_grid]) like the factory kwargs."""
        nd = grids[0].ndim
        coords = []
        for ax in range(nd):  # array axis ax <-> coordinate component (nd-1-ax)
            g = grids[nd - 1 - ax]
            idx = [0] * nd
            idx[ax] = slice(None)
            coords.append(g[tuple(idx)])
        if vector:
            self.s.penalise_field_boundary_vector(f, width, dx, coords)
        else:
            self.s.penalise_field_boundary(f, width, dx, coords)

    def brinkmann(self, o, f, chi, pen, factor, vector=False):
        if vector:
            for c in range(f.shape[0]):
                self.s.brinkmann_penalise(o[c], f[c], chi, pen[c], factor)
        else:
            self.s.brinkmann_penalise(o, f, chi, pen, factor)

    def brinkmann_vs_fixed_val(self, o, f, chi, factor, val, vector=False):
        if vector:
            for c in range(f.shape[0]):
                self.s.brinkmann_penalise_vs_fixed_val(o[c], f[c], chi, factor, val[c])
        else:
            self.s.brinkmann_penalise_vs_fixed_val(o, f, chi, factor, val)

    def char_func(self, o, ls, blend_width):
        self.s.char_func_from_level_set_via_sine_heaviside(o, ls, blend_width)

    def laplacian_filter(self, f, flux_buf, field_buf, order, ftype, vector=False):
        if vector:
            self.s.laplacian_filter_3d_vector(f, flux_buf, field_buf, order, ftype)
        else:
            self.s.laplacian_filter_3d(f, flux_buf, field_buf, order, ftype)

    def outplane_curl_2d(self, curl, f, p, reset=True):
        self.s.outplane_field_curl_2d(curl, f, p, reset)

    def inplane_curl_2d(self, curl, f, p):
        self.s.inplane_field_curl_2d(curl, f, p)

    # poisson
    def poisson_solver(self, grid, x_range):
        if len(grid) == 3:
            return self.p.UnboundedPoissonSolver3D(*grid, x_range=x_range, real_t=self.real_t)
        return self.p.UnboundedPoissonSolver2D(*grid, x_range=x_range, real_t=self.real_t)


class CudaOps:
    """The product path: reference-named factories from sopht_b200 (numpy args are staged via the GPU)."""

    name = "cuda"

    def __init__(self, real_t):
        import sopht_b200.numeric.eulerian_grid_ops as spne

        self.real_t = real_t
        self.spne = spne

    def _k(self, base, nd, **kw):
        return getattr(self.spne, f"{base}_{nd}d")(real_t=self.real_t, **kw)

    @staticmethod
    def _gdim(a, vector=False):
        return a.ndim - (1 if vector else 0)

    def elementwise_sum(self, o, a, b):
        nd = min(o.ndim, 3)
        ft = "vector" if o.ndim == 4 else "scalar"
        self._k("gen_elementwise_sum_pyst_kernel", nd, field_type=ft)(sum_field=o, field_1=a, field_2=b)

    def set_fixed_val(self, f, v):
        self._k("gen_set_fixed_val_pyst_kernel", f.ndim)(field=f, fixed_val=v)

    def set_fixed_val_vector(self, f, vals):
        self._k("gen_set_fixed_val_pyst_kernel", f.ndim - 1, field_type="vector")(vector_field=f, fixed_vals=vals)

    def elementwise_copy(self, f, r):
        self._k("gen_elementwise_copy_pyst_kernel", f.ndim)(field=f, rhs_field=r)

    def complex_product(self, o, a, b):
        self._k("gen_elementwise_complex_product_pyst_kernel", o.ndim)(product_field=o, field_1=a, field_2=b)

    def set_fixed_val_at_boundaries(self, f, width, v):
        self._k("gen_set_fixed_val_at_boundaries_pyst_kernel", f.ndim, width=width)(field=f, fixed_val=v)

    def set_fixed_val_at_boundaries_vector(self, f, width, vals):
        self._k("gen_set_fixed_val_at_boundaries_pyst_kernel", f.ndim - 1, width=width, field_type="vector")(
            vector_field=f, fixed_vals=vals)

    def add_fixed_val(self, o, f, v):
        self._k("gen_add_fixed_val_pyst_kernel", f.ndim)(sum_field=o, field=f, fixed_val=v)

    def add_fixed_val_vector(self, o, f, vals):
        self._k("gen_add_fixed_val_pyst_kernel", f.ndim - 1, field_type="vector")(
            sum_field=o, vector_field=f, fixed_vals=vals)

    def saxpby(self, o, a, b, pa, pb):
        nd = min(o.ndim, 3)
        ft = "vector" if o.ndim == 4 else "scalar"
        self._k("gen_elementwise_saxpby_pyst_kernel", nd, field_type=ft)(
            sum_field=o, field_1=a, field_2=b, field_1_prefac=pa, field_2_prefac=pb)

    def cross_product(self, o, a, b):
        self._k("gen_elementwise_cross_product_pyst_kernel", 3)(result_field=o, field_1=a, field_2=b)

    def diffusion_flux(self, flux, field, p, reset=True, vector=False):
        nd = self._gdim(field, vector)
        if nd == 3:
            k = self._k("gen_diffusion_flux_pyst_kernel", 3, field_type="vector" if vector else "scalar",
                        reset_ghost_zone=reset)
            if vector:
                k(vector_field_diffusion_flux=flux, vector_field=field, prefactor=p)
            else:
                k(diffusion_flux=flux, field=field, prefactor=p)
        else:
            self._k("gen_diffusion_flux_pyst_kernel", 2, reset_ghost_zone=reset)(
                diffusion_flux=flux, field=field, prefactor=p)

    def diffusion_timestep(self, field, flux, p, vector=False):
        nd = self._gdim(field, vector)
        if nd == 3:
            k = self._k("gen_diffusion_timestep_euler_forward_pyst_kernel", 3,
                        field_type="vector" if vector else "scalar")
            if vector:
                k(vector_field=field, diffusion_flux=flux, nu_dt_by_dx2=p)
            else:
                k(field=field, diffusion_flux=flux, nu_dt_by_dx2=p)
        else:
            self._k("gen_diffusion_timestep_euler_forward_pyst_kernel", 2)(
                field=field, diffusion_flux=flux, nu_dt_by_dx2=p)

    def curl_3d(self, curl, field, p, reset=True):
        self._k("gen_curl_pyst_kernel", 3, reset_ghost_zone=reset)(curl=curl, field=field, prefactor=p)

    def divergence_3d(self, div, field, inv_dx, reset=True):
        self._k("gen_divergence_pyst_kernel", 3, reset_ghost_zone=reset)(divergence=div, field=field, inv_dx=inv_dx)

    def forcing_update(self, w, f, p):
        nd = 2 if w.ndim == 2 else 3
        self._k("gen_update_vorticity_from_velocity_forcing_pyst_kernel", nd)(
            vorticity_field=w, velocity_forcing_field=f, prefactor=p)

    def penalised_velocity_update(self, w, pv, u, p):
        nd = 2 if w.ndim == 2 else 3
        self._k("gen_update_vorticity_from_penalised_velocity_pyst_kernel", nd)(
            vorticity_field=w, penalised_velocity_field=pv, velocity_field=u, prefactor=p)

    def stretching_flux(self, q, w, u, p):
        self._k("gen_vorticity_stretching_flux_pyst_kernel", 3)(
            vorticity_stretching_flux_field=q, vorticity_field=w, velocity_field=u, prefactor=p)

    def stretching_timestep(self, w, u, q, p, stepper="euler_forward", midstep=None):
        if stepper == "euler_forward":
            k = self._k("gen_vorticity_stretching_timestep_euler_forward_pyst_kernel", 3)
        else:
            k = self._k("gen_vorticity_stretching_timestep_ssprk3_pyst_kernel", 3,
                        midstep_buffer_vector_field=midstep)
        k(vorticity_field=w, velocity_field=u, vorticity_stretching_flux_field=q, dt_by_2_dx=p)

    def advection_flux(self, q, f, v, inv_dx):
        self._k("gen_advection_flux_conservative_eno3_pyst_kernel", f.ndim)(
            advection_flux=q, field=f, velocity=v, inv_dx=inv_dx)

    def advection_timestep(self, f, q, v, dt_by_dx, vector=False):
        nd = self._gdim(f, vector)
        if nd == 3:
            k = self._k("gen_advection_timestep_euler_forward_conservative_eno3_pyst_kernel", 3,
                        field_type="vector" if vector else "scalar")
            if vector:
                k(vector_field=f, advection_flux=q, velocity=v, dt_by_dx=dt_by_dx)
            else:
                k(field=f, advection_flux=q, velocity=v, dt_by_dx=dt_by_dx)
        else:
            self._k("gen_advection_timestep_euler_forward_conservative_eno3_pyst_kernel", 2)(
                field=f, advection_flux=q, velocity=v, dt_by_dx=dt_by_dx)

    def penalise(self, f, width, dx, grids, vector=False):
        nd = grids[0].ndim
        if nd == 3:
            k = self.spne.gen_penalise_field_boundary_pyst_kernel_3d(
                width=width, dx=dx, x_grid_field=grids[0], y_grid_field=grids[1], z_grid_field=grids[2],
                real_t=self.real_t, field_type="vector" if vector else "scalar")
            if vector:
                k(vector_field=f)
            else:
                k(field=f)
        else:
            self.spne.gen_penalise_field_boundary_pyst_kernel_2d(
                width=width, dx=dx, x_grid_field=grids[0], y_grid_field=grids[1], real_t=self.real_t)(field=f)

    def brinkmann(self, o, f, chi, pen, factor, vector=False):
        nd = self._gdim(f, vector)
        k = self._k("gen_brinkmann_penalise_pyst_kernel", nd, field_type="vector" if vector else "scalar")
        if vector:
            k(penalised_vector_field=o, penalty_factor=factor, char_field=chi, penalty_vector_field=pen,
              vector_field=f)
        else:
            k(penalised_field=o, field=f, char_field=chi, penalty_field=pen, penalty_factor=factor)

    def brinkmann_vs_fixed_val(self, o, f, chi, factor, val, vector=False):
        k = self._k("gen_brinkmann_penalise_vs_fixed_val_pyst_kernel", 2,
                    field_type="vector" if vector else "scalar")
        if vector:
            k(penalised_vector_field=o, penalty_factor=factor, char_field=chi, penalty_val=val, vector_field=f)
        else:
            k(penalised_field=o, field=f, char_field=chi, penalty_factor=factor, penalty_val=val)

    def char_func(self, o, ls, blend_width):
        fac = getattr(self.spne, f"gen_char_func_from_level_set_via_sine_heaviside_pyst_kernel_{ls.ndim}d")
        fac(blend_width=blend_width, real_t=self.real_t)(char_func_field=o, level_set_field=ls)

    def laplacian_filter(self, f, flux_buf, field_buf, order, ftype, vector=False):
        k = self.spne.gen_laplacian_filter_kernel_3d(
            filter_order=order, filter_flux_buffer=flux_buf, field_buffer=field_buf, real_t=self.real_t,
            field_type="vector" if vector else "scalar", filter_type=ftype)
        k(f)

    def outplane_curl_2d(self, curl, f, p, reset=True):
        self._k("gen_outplane_field_curl_pyst_kernel", 2, reset_ghost_zone=reset)(curl=curl, field=f, prefactor=p)

    def inplane_curl_2d(self, curl, f, p):
        self._k("gen_inplane_field_curl_pyst_kernel", 2)(curl=curl, field=f, prefactor=p)

    def poisson_solver(self, grid, x_range, flags=0):
        if len(grid) == 3:
            return self.spne.UnboundedPoissonSolverPYFFTW3D(*grid, x_range=x_range, real_t=self.real_t, flags=flags)
        return self.spne.UnboundedPoissonSolverPYFFTW2D(*grid, x_range=x_range, real_t=self.real_t, flags=flags)


class COracleOps(OracleOps):
    """The oracle with its stencil / elementwise passes executed by the C + OpenMP restatement
    (oracle/c/ref_kernels.c through oracle/cstencils.py) - the CPU baseline bench.py times."""

    name = "coracle"

    def __init__(self, real_t):
        super().__init__(real_t)
        from oracle import cstencils

        cstencils.load()
        self.s = cstencils

    def poisson_solver(self, grid, x_range):
        if len(grid) == 3:
            return self.p.UnboundedPoissonSolver3D(*grid, x_range=x_range, real_t=self.real_t, kernels=self.s)
        return self.p.UnboundedPoissonSolver2D(*grid, x_range=x_range, real_t=self.real_t, kernels=self.s)


def make_ops(kind: str, precision: str):
    real_t = np.float32 if precision == "single" else np.float64
    if kind == "coracle":
        return COracleOps(real_t)
    return OracleOps(real_t) if kind == "oracle" else CudaOps(real_t)
