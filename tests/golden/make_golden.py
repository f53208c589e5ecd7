"""Generate the golden fixtures in tests/golden/*.npz FROM THE REFERENCE (run in the build container only).

    NUMBA_CACHE_DIR=/tmp/numba_cache python tests/golden/make_golden.py [/root/reference [fastdiag]]

What is executed:
  * stencils / Poisson: the numpy reference functions and seeded ``*Solution`` classes that live inside
    the reference's own test modules (tests/test_numeric/test_eulerian_grid_ops/**), imported unmodified.
    Those modules import ``sopht.numeric.eulerian_grid_ops`` whose dependencies pystencils / pyfftw /
    h5py / matplotlib are not installed here, so those third-party names are stubbed at import time —
    none of the stubbed code is executed, only the pure-numpy references are.
  * immersed boundary: the reference's OWN implementation (numba) —
    EulerianLagrangianGridCommunicator{2,3}D and VirtualBoundaryForcing — is run on seeded inputs.
  * Neumann (fast-diagonalisation) Poisson: the reference's OWN FastDiagPoissonSolver{2,3}D (numpy + scipy.sparse).
Every fixture stores inputs and expected outputs; tests/test_oracle_golden.py pins oracle/ against them
and the gpu tests pin the CUDA path against them. /root/reference is NOT needed to run the tests.
"""

from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import os
import sys
from unittest.mock import MagicMock

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
SEED = 42  # tests/conftest.py:21-24 of the reference
N = 16  # the reference tests' grid size


class _StubLoader(importlib.abc.Loader):
    def create_module(self, spec):
        m = MagicMock(name=spec.name)
        m.__path__ = []
        m.__name__ = spec.name
        m.__spec__ = spec
        return m

    def exec_module(self, module):
        pass


class _StubFinder(importlib.abc.MetaPathFinder):
    ROOTS = {"pystencils", "pyfftw", "h5py", "matplotlib", "elastica", "ffmpeg", "magneto_pyelastica", "click"}

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.ROOTS:
            return importlib.machinery.ModuleSpec(name, _StubLoader(), is_package=True)
        return None


def _setup_imports():
    sys.meta_path.append(_StubFinder())
    sys.path.insert(0, REF)
    base = os.path.join(REF, "tests", "test_numeric")
    for sub in (
        "test_eulerian_grid_ops/test_stencil_ops_3d",
        "test_eulerian_grid_ops/test_stencil_ops_2d",
        "test_eulerian_grid_ops/test_poisson_solver_3d",
        "test_eulerian_grid_ops/test_poisson_solver_2d",
        "test_immersed_boundary_ops",
    ):
        sys.path.insert(0, os.path.join(base, sub))


def _arrays_of(obj, skip=("domain_doubled", "fourier")) -> dict:
    out = {}
    for k, v in vars(obj).items():
        if any(s in k for s in skip):
            continue
        if isinstance(v, np.ndarray):
            out[k] = v
        elif isinstance(v, (np.floating, np.integer, float, int)) and not isinstance(v, bool):
            out[k] = np.asarray(v)
    return out


def _save(name: str, cases: dict) -> None:
    flat = {}
    for case, arrays in cases.items():
        for k, v in arrays.items():
            flat[f"{case}/{k}"] = v
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **flat)
    print(f"wrote {path}: {len(flat)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


SOLUTION_CASES_3D = [
    # (module, class, case name, extra kwargs)
    ("test_advection_flux_3d", "AdvectionSolution", "advection_flux", {"flux_type": "conservative_eno3"}),
    ("test_advection_timestep_3d", "AdvectionTimestepEulerForwardSolution", "advection_timestep", {}),
    ("test_brinkmann_penalise_3d", "BrinkmannPenalisationSolution", "brinkmann_penalise", {}),
    ("test_char_func_from_level_set_3d", "CharFuncFromLevelSetFuncSolution", "char_func", {}),
    ("test_curl_3d", "CurlSolution", "curl", {}),
    ("test_diffusion_flux_3d", "DiffusionFluxSolution", "diffusion_flux", {}),
    ("test_diffusion_timestep_3d", "DiffusionTimestepEulerForwardSolution", "diffusion_timestep", {}),
    ("test_divergence_3d", "DivergenceSolution", "divergence", {}),
    ("test_penalise_field_boundary_3d", "PenaliseFieldBoundarySolution", "penalise", {}),
    ("test_update_vorticity_from_velocity_forcing_3d", "UpdateVorticityFromVelocityForcingSolution", "forcing_update", {}),
    ("test_vorticity_stretching_flux_3d", "VorticityStretchingFluxSolution", "stretching_flux", {}),
    ("test_vorticity_stretching_timestep_3d", "VorticityStretchingTimestepSolution", "stretching_timestep_euler", {"time_stepper": "euler_forward"}),
    ("test_vorticity_stretching_timestep_3d", "VorticityStretchingTimestepSolution", "stretching_timestep_ssprk3", {"time_stepper": "ssprk3"}),
]

SOLUTION_CASES_2D = [
    ("test_advection_flux_2d", "AdvectionFluxSolution", "advection_flux", {}),
    ("test_advection_timestep_2d", "AdvectionTimestepSolution", "advection_timestep", {}),
    ("test_brinkmann_penalise_2d", "BrinkmannPenalisationSolution", "brinkmann_penalise", {}),
    ("test_char_func_from_level_set_2d", "CharFuncFromLevelSetFuncSolution", "char_func", {}),
    ("test_diffusion_flux_2d", "DiffusionFluxSolution", "diffusion_flux", {}),
    ("test_diffusion_timestep_2d", "DiffusionTimestepSolution", "diffusion_timestep", {}),
    ("test_inplane_field_curl_2d", "InplaneCurlSolution", "inplane_curl", {}),
    ("test_outplane_field_curl_2d", "OutplaneCurlSolution", "outplane_curl", {}),
    ("test_penalise_field_boundary_2d", "PenaliseFieldBoundarySolution", "penalise", {}),
    ("test_update_vorticity_from_velocity_forcing_2d", "UpdateVorticityFromVelocityForcingSolution", "forcing_update", {}),
]


def _find_class(mod, wanted: str):
    if hasattr(mod, wanted):
        return getattr(mod, wanted)
    cands = [n for n in dir(mod) if n.endswith("Solution")]
    if len(cands) == 1:
        return getattr(mod, cands[0])
    for n in cands:
        if wanted.lower()[:12] in n.lower():
            return getattr(mod, n)
    raise AttributeError(f"{mod.__name__}: no class {wanted}; candidates {cands}")


def stencil_goldens(cases, tag):
    for precision in ("single", "double"):
        out = {}
        for modname, clsname, case, extra in cases:
            mod = importlib.import_module(modname)
            try:
                cls = _find_class(mod, clsname)
            except AttributeError as e:
                print("skip:", e)
                continue
            rng = np.random.default_rng(SEED)
            sol = cls(n_samples=N, rng_generator=rng, precision=precision, **extra)
            out[case] = _arrays_of(sol)
        if tag == "stencils3d":
            lf = importlib.import_module("test_laplacian_filter_3d")
            real_t = np.float32 if precision == "single" else np.float64
            for filter_type in ("convolution", "multiplicative"):
                for order in (1, 2):
                    rng = np.random.default_rng(SEED)
                    f = rng.random((N, N, N)).astype(real_t)
                    v = rng.random((3, N, N, N)).astype(real_t)
                    fo, vo = f.copy(), v.copy()
                    lf.scalar_laplacian_filter(scalar_field=fo, filter_order=order, filter_type=filter_type)
                    lf.vector_laplacian_filter(vector_field=vo, filter_order=order, filter_type=filter_type)
                    out[f"laplacian_filter_{filter_type}_{order}"] = {
                        "field": f, "vector_field": v, "ref_field": fo, "ref_vector_field": vo,
                    }
        _save(f"{tag}_{precision}", out)


def poisson_goldens():
    p3 = importlib.import_module("test_unbounded_poisson_solver_3d")
    p2 = importlib.import_module("test_unbounded_poisson_solver_2d")
    for precision in ("single", "double"):
        real_t = np.float32 if precision == "single" else np.float64
        out = {}
        rng = np.random.default_rng(SEED)
        s3 = p3.UnboundedPoissonSolverSolution3D(
            grid_size_z=N, grid_size_y=N, grid_size_x=N, x_range=real_t(2.0), rng_generator=rng, precision=precision)
        out["poisson3d_16"] = _arrays_of(s3)
        rng = np.random.default_rng(SEED)
        s3b = p3.UnboundedPoissonSolverSolution3D(
            grid_size_z=8, grid_size_y=12, grid_size_x=20, x_range=real_t(1.0), rng_generator=rng, precision=precision)
        out["poisson3d_8x12x20"] = _arrays_of(s3b)
        rng = np.random.default_rng(SEED)
        cls2 = [getattr(p2, n) for n in dir(p2) if n.endswith("Solution2D") or n.endswith("Solution")][0]
        import inspect

        params = inspect.signature(cls2.__init__).parameters
        kw = {"rng_generator": rng, "precision": precision}
        if "grid_size_y" in params:
            kw.update(grid_size_y=N, grid_size_x=N, x_range=real_t(2.0))
        else:
            kw.update(n_samples=N)
        s2 = cls2(**kw)
        out["poisson2d_16"] = _arrays_of(s2)
        _save(f"poisson_{precision}", out)


def _load_ref_module(relpath: str, name: str):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ib_goldens():
    """Run the reference numba communicator + VirtualBoundaryForcing on seeded inputs."""
    import importlib.util  # noqa: F401

    from sopht.numeric.immersed_boundary_ops.EulerianLagrangianGridCommunicator2D import (
        EulerianLagrangianGridCommunicator2D,
    )
    from sopht.numeric.immersed_boundary_ops.EulerianLagrangianGridCommunicator3D import (
        EulerianLagrangianGridCommunicator3D,
    )
    from sopht.numeric.immersed_boundary_ops.VirtualBoundaryForcing import VirtualBoundaryForcing

    for precision in ("single", "double"):
        real_t = np.float32 if precision == "single" else np.float64
        out = {}
        for dim, grid in ((3, (20, 24, 28)), (2, (40, 48))):
            for kernel_type in ("cosine", "peskin"):
                rng = np.random.default_rng(SEED)
                nx = grid[-1]
                dx = real_t(1.0 / nx)
                shift = real_t(dx / 2)
                n_lag = 37
                ranges = [dx * n for n in grid[::-1]]  # x, y, (z) extents
                # float64 positions, as forcing grids deliver them; keep the 4-tap support inside the grid
                pos = np.stack([rng.uniform(3 * dx, r - 3 * dx, n_lag) for r in ranges]).astype(np.float64)
                Comm = EulerianLagrangianGridCommunicator3D if dim == 3 else EulerianLagrangianGridCommunicator2D
                comm = Comm(dx=dx, eul_grid_coord_shift=shift, num_lag_nodes=n_lag, interp_kernel_width=2,
                            real_t=real_t, n_components=dim, interp_kernel_type=kernel_type)
                comm_s = Comm(dx=dx, eul_grid_coord_shift=shift, num_lag_nodes=n_lag, interp_kernel_width=2,
                              real_t=real_t, n_components=1, interp_kernel_type=kernel_type)
                idx = np.empty((dim, n_lag), dtype=int)
                support = np.empty((dim,) + (4,) * dim + (n_lag,), dtype=real_t)
                weights = np.empty((4,) * dim + (n_lag,), dtype=real_t)
                comm.local_eulerian_grid_support_of_lagrangian_grid_kernel(support, idx, pos)
                support_before = support.copy()
                comm.interpolation_weights_kernel(weights, support)
                eul_vec = rng.standard_normal((dim, *grid)).astype(real_t)
                eul_sca = rng.standard_normal(grid).astype(real_t)
                lag_vec = np.zeros((dim, n_lag), dtype=real_t)
                lag_sca = np.zeros(n_lag, dtype=real_t)
                comm.eulerian_to_lagrangian_grid_interpolation_kernel(lag_vec, eul_vec, weights, idx)
                comm_s.eulerian_to_lagrangian_grid_interpolation_kernel(lag_sca, eul_sca, weights, idx)
                force_vec = rng.standard_normal((dim, n_lag)).astype(real_t)
                force_sca = rng.standard_normal(n_lag).astype(real_t)
                spread_vec = np.zeros((dim, *grid), dtype=real_t)
                spread_sca = np.zeros(grid, dtype=real_t)
                comm.lagrangian_to_eulerian_grid_interpolation_kernel(spread_vec, force_vec, weights, idx)
                comm_s.lagrangian_to_eulerian_grid_interpolation_kernel(spread_sca, force_sca, weights, idx)
                out[f"comm{dim}d_{kernel_type}"] = {
                    "dx": np.asarray(dx), "shift": np.asarray(shift), "lag_positions": pos,
                    "nearest_idx": idx, "local_support": support_before, "local_support_after_weights": support,
                    "interp_weights": weights, "eul_vector_field": eul_vec, "eul_scalar_field": eul_sca,
                    "lag_vector_field": lag_vec, "lag_scalar_field": lag_sca,
                    "lag_force_vector": force_vec, "lag_force_scalar": force_sca,
                    "spread_vector_field": spread_vec, "spread_scalar_field": spread_sca,
                }
            # virtual boundary forcing: three coupled steps (position mismatch accumulates between them)
            rng = np.random.default_rng(SEED + 1)
            nx = grid[-1]
            dx = real_t(1.0 / nx)
            n_lag = 29
            ranges = [dx * n for n in grid[::-1]]
            pos = np.stack([rng.uniform(3 * dx, r - 3 * dx, n_lag) for r in ranges]).astype(np.float64)
            body_vel = (0.1 * rng.standard_normal((dim, n_lag))).astype(np.float64)
            vb = VirtualBoundaryForcing(
                virtual_boundary_stiffness_coeff=real_t(-5e2), virtual_boundary_damping_coeff=real_t(-1e1),
                grid_dim=dim, dx=dx, num_lag_nodes=n_lag, real_t=real_t, enable_eul_grid_forcing_reset=False)
            eul_vel = rng.standard_normal((dim, *grid)).astype(real_t)
            eul_force = np.zeros((dim, *grid), dtype=real_t)
            dt = real_t(1e-3)
            hist = {}
            for step in range(3):
                vb.time_step(dt)
                vb.compute_interaction_forcing(eul_force, eul_vel, pos, body_vel)
                hist[f"lag_forcing_step{step}"] = vb.lag_grid_forcing_field.copy()
                hist[f"eul_forcing_step{step}"] = eul_force.copy()
                hist[f"position_mismatch_step{step}"] = vb.lag_grid_position_mismatch_field.copy()
            out[f"virtual_boundary_{dim}d"] = {
                "dx": np.asarray(dx), "dt": np.asarray(dt), "stiffness": np.asarray(real_t(-5e2)),
                "damping": np.asarray(real_t(-1e1)), "lag_positions": pos, "lag_body_velocity": body_vel,
                "eul_velocity_field": eul_vel, **hist,
            }
        _save(f"ib_{precision}", out)


def fastdiag_goldens():
    """Run the reference's OWN FastDiagPoissonSolver{2,3}D (numpy / scipy.sparse only) on seeded right-hand sides:
    non-cubic grids, one zero-mean and one general rhs, scalar and vector solves."""
    from sopht.numeric.eulerian_grid_ops.poisson_solver_2d.FastDiagPoissonSolver2D import FastDiagPoissonSolver2D
    from sopht.numeric.eulerian_grid_ops.poisson_solver_3d.FastDiagPoissonSolver3D import FastDiagPoissonSolver3D

    for precision in ("single", "double"):
        real_t = np.float32 if precision == "single" else np.float64
        rng = np.random.default_rng(SEED)
        out = {}
        nz, ny, nx = 12, 10, 16
        dx = real_t(1.0 / nx)
        solver = FastDiagPoissonSolver3D(grid_size_z=nz, grid_size_y=ny, grid_size_x=nx, dx=dx, real_t=real_t)
        rhs = rng.standard_normal((3, nz, ny, nx)).astype(real_t)
        rhs[1] -= rhs[1].mean()
        sol = np.zeros_like(rhs)
        solver.vector_field_solve(solution_vector_field=sol, rhs_vector_field=rhs)
        scalar = np.zeros((nz, ny, nx), dtype=real_t)
        solver.solve(solution_field=scalar, rhs_field=rhs[2])
        out["neumann3d"] = {"dx": np.asarray(dx), "rhs": rhs, "solution": sol, "scalar_solution_of_rhs2": scalar}
        ny, nx = 14, 24
        dx = real_t(1.0 / nx)
        solver2 = FastDiagPoissonSolver2D(grid_size_y=ny, grid_size_x=nx, dx=dx, real_t=real_t)
        rhs2 = rng.standard_normal((ny, nx)).astype(real_t)
        sol2 = np.zeros_like(rhs2)
        solver2.solve(solution_field=sol2, rhs_field=rhs2)
        out["neumann2d"] = {"dx": np.asarray(dx), "rhs": rhs2, "solution": sol2}
        _save(f"fastdiag_{precision}", out)


if __name__ == "__main__":
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
    _setup_imports()
    if len(sys.argv) > 2 and sys.argv[2] == "fastdiag":  # regenerate this fixture only
        fastdiag_goldens()
        sys.exit(0)
    stencil_goldens(SOLUTION_CASES_3D, "stencils3d")
    stencil_goldens(SOLUTION_CASES_2D, "stencils2d")
    poisson_goldens()
    ib_goldens()
    fastdiag_goldens()
