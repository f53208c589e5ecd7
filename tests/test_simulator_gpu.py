"""GPU: simulator-level flow steps (the step sequencers) against the oracle's
restatement of the reference steps on identical seeded state — the composition-identity strategy of the
reference's tests/test_simulator/test_flow/test_navier_stokes_flow_simulators.py:12-338."""

import numpy as np
import pytest
from conftest import REL_L2_TOL, real_t_of, rel_l2

pytestmark = pytest.mark.gpu


def _seed_state(sim, ref, rng, real_t, names):
    import torch

    for name in names:
        a = rng.standard_normal(getattr(ref, name).shape).astype(real_t)
        getattr(ref, name)[...] = a
        getattr(sim, name)[...] = torch.from_numpy(a).cuda()


@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("with_forcing", [False, True])
@pytest.mark.parametrize("with_free_stream", [False, True])
@pytest.mark.parametrize("filter_vorticity", [False, True])
@pytest.mark.parametrize("step_mode", ["unfused", "auto"])
# (10, 11, 15): x extent not a multiple of the vector width -> shared-memory fused kernels; (40, 12, 132) and
# (6, 20, 260): several z chunks / several (partially filled) x tiles of the register-marching kernels
@pytest.mark.parametrize("grid", [(16, 16, 16), (8, 12, 20), (32, 16, 64), (10, 11, 15), (40, 12, 132),
                                  (6, 20, 260)])
def test_navier_stokes_3d_step(rng, precision, with_forcing, with_free_stream, filter_vorticity, step_mode, grid):
    from oracle import flow as oflow
    from sopht_b200.simulator import UnboundedNavierStokesFlowSimulator3D

    real_t = real_t_of(precision)
    kw = dict(grid_size=grid, x_range=1.0, kinematic_viscosity=1e-2, real_t=real_t,
              with_forcing=with_forcing, with_free_stream_flow=with_free_stream,
              filter_vorticity=filter_vorticity,
              filter_setting_dict={"order": 2, "type": "convolution"}, flow_density=0.5)
    sim = UnboundedNavierStokesFlowSimulator3D(step_mode=step_mode, **kw)
    ref = oflow.UnboundedNavierStokesFlowSimulator3D(**kw)
    names = ["vorticity_field", "velocity_field"] + (["eul_grid_forcing_field"] if with_forcing else [])
    _seed_state(sim, ref, rng, real_t, names)
    dt_ref = ref.compute_stable_timestep(dt_prefac=0.5)
    dt_sim = sim.compute_stable_timestep(dt_prefac=0.5)
    rel = 1e-6 if precision == "single" else 1e-13
    assert dt_sim == pytest.approx(dt_ref, rel=rel)
    fsv = np.array([1.0, 2.0, 3.0]) if with_free_stream else np.zeros(3)
    for _ in range(2):
        sim.time_step(dt=dt_ref, free_stream_velocity=fsv)
        ref.time_step(dt_ref, free_stream_velocity=fsv)
        # stable dt after a step (the fused path serves it from the device-side reduction)
        assert sim.compute_stable_timestep() == pytest.approx(ref.compute_stable_timestep(), rel=10 * rel)
    assert sim.time == pytest.approx(ref.time)
    for name in ["vorticity_field", "velocity_field", "stream_func_field"]:
        err = rel_l2(getattr(sim, name).cpu().numpy(), getattr(ref, name))
        assert err < REL_L2_TOL[precision], (name, err)
    if with_forcing:
        assert float(sim.eul_grid_forcing_field.abs().max()) == 0.0
    d_sim, d_ref = sim.get_vorticity_divergence_l2_norm(), None
    assert np.isfinite(d_sim)


@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("with_forcing", [False, True])
@pytest.mark.parametrize("grid", [(16, 16), (24, 40)])
def test_navier_stokes_2d_step(rng, precision, with_forcing, grid):
    from oracle import flow as oflow
    from sopht_b200.simulator import UnboundedNavierStokesFlowSimulator2D

    real_t = real_t_of(precision)
    kw = dict(grid_size=grid, x_range=1.0, kinematic_viscosity=1e-2, real_t=real_t,
              with_forcing=with_forcing, with_free_stream_flow=True, flow_density=2.0)
    sim = UnboundedNavierStokesFlowSimulator2D(**kw)
    ref = oflow.UnboundedNavierStokesFlowSimulator2D(**kw)
    names = ["vorticity_field", "velocity_field"] + (["eul_grid_forcing_field"] if with_forcing else [])
    _seed_state(sim, ref, rng, real_t, names)
    dt = ref.compute_stable_timestep()
    assert sim.compute_stable_timestep() == pytest.approx(dt, rel=1e-6)
    for _ in range(2):
        sim.time_step(dt=dt, free_stream_velocity=[1.0, -1.0])
        ref.time_step(dt, free_stream_velocity=[1.0, -1.0])
    for name in ["vorticity_field", "velocity_field", "stream_func_field"]:
        err = rel_l2(getattr(sim, name).cpu().numpy(), getattr(ref, name))
        assert err < REL_L2_TOL[precision], (name, err)


# ---- BASELINE.json sizes: size-independent properties (the oracle takes minutes there) ------------------------------
@pytest.mark.parametrize("grid", [(128, 128, 256), (256, 256, 256), (512, 512, 512)])
def test_poisson_full_size_pow2_vs_cufft_path_and_linearity(grid):
    """C2 / 256^3 / 512^3 (the bench default: the warp-quartet z pass at full size): the hand-written pruned FFT pipeline against the cuFFT-based generic path of the same library
    (independent code: padded doubled domain, cuFFT transforms), plus linearity of the solve."""
    import torch

    from sopht_b200.numeric.eulerian_grid_ops import UnboundedPoissonSolverPYFFTW3D
    from sopht_b200.numeric.eulerian_grid_ops.poisson_solvers import POISSON_FORCE_GENERIC

    gen = torch.Generator(device="cuda").manual_seed(3)
    rhs = torch.randn(3, *grid, device="cuda", generator=gen)
    fast = UnboundedPoissonSolverPYFFTW3D(*grid, x_range=1.0, real_t=np.float32)
    slow = UnboundedPoissonSolverPYFFTW3D(*grid, x_range=1.0, real_t=np.float32, flags=POISSON_FORCE_GENERIC)
    assert fast.path == "pow2" and slow.path != "pow2"
    a, b = torch.zeros_like(rhs), torch.zeros_like(rhs)
    fast.vector_field_solve(solution_vector_field=a, rhs_vector_field=rhs)
    slow.vector_field_solve(solution_vector_field=b, rhs_vector_field=rhs)
    assert float((a - b).norm() / b.norm()) < 1e-5
    # linearity: solve(2 f0 - 3 f1) = 2 solve(f0) - 3 solve(f1)
    mix = torch.zeros(*grid, device="cuda")
    fast.solve(solution_field=mix, rhs_field=2 * rhs[0] - 3 * rhs[1])
    want = 2 * a[0] - 3 * a[1]
    assert float((mix - want).norm() / want.norm()) < 1e-5


def test_c2_step_fused_and_unfused_vs_oracle():
    """BASELINE configs[1] size (128x128x256, forcing + free stream): the fused step (register-marching kernels, pruned
    FFT Poisson, side-stream Nyquist plane) AND the one-kernel-per-reference-call step of the same library, each
    against the CPU oracle's restatement of the reference step, two coupled steps from the same seeded random state
    (the forcing field is re-filled between the steps, so the f = 0 reset of the forced step is exercised too)."""
    import torch

    from oracle import cstencils
    from oracle import flow as oflow
    from sopht_b200.simulator import UnboundedNavierStokesFlowSimulator3D

    grid = (128, 128, 256)
    kw = dict(grid_size=grid, x_range=1.0, kinematic_viscosity=1e-3, real_t=np.float32, with_forcing=True,
              with_free_stream_flow=True)
    sims = [UnboundedNavierStokesFlowSimulator3D(step_mode=m, **kw) for m in ("fused", "unfused")]
    ref = oflow.UnboundedNavierStokesFlowSimulator3D(workers=8, kernels=cstencils, **kw)
    rng = np.random.default_rng(5)
    state = {n: rng.standard_normal((3, *grid)).astype(np.float32)
             for n in ("vorticity_field", "velocity_field", "eul_grid_forcing_field")}
    for n, v in state.items():
        getattr(ref, n)[...] = v
        for s in sims:
            getattr(s, n)[...] = torch.from_numpy(v).cuda()
    dt = ref.compute_stable_timestep(dt_prefac=0.5)
    for s in sims:
        assert s.compute_stable_timestep(dt_prefac=0.5) == pytest.approx(dt, rel=1e-6)
    for _ in range(2):
        ref.time_step(dt, free_stream_velocity=[1.0, 0.0, 0.0])
        assert float(np.abs(ref.eul_grid_forcing_field).max()) == 0.0
        ref.eul_grid_forcing_field[...] = state["eul_grid_forcing_field"]
        for s in sims:
            s.time_step(dt=dt, free_stream_velocity=[1.0, 0.0, 0.0])
            assert float(s.eul_grid_forcing_field.abs().max()) == 0.0
            s.eul_grid_forcing_field[...] = torch.from_numpy(state["eul_grid_forcing_field"]).cuda()
    for n in ("vorticity_field", "velocity_field", "stream_func_field"):
        want = getattr(ref, n).astype(np.float64)
        for s in sims:
            got = getattr(s, n).cpu().numpy().astype(np.float64)
            err = np.linalg.norm(got - want) / np.linalg.norm(want)
            assert err < 1e-5, (s.step_mode, n, err)
    want_dt = ref.compute_stable_timestep()
    for s in sims:
        assert s.compute_stable_timestep() == pytest.approx(want_dt, rel=1e-5)


# ---- passive transport and the backward-compatibility factories ----------------------------------
@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("case", [(2, (24, 40), "scalar"), (3, (12, 20, 36), "scalar"), (3, (12, 20, 36), "vector")])
def test_passive_transport_step(rng, precision, case):
    """PassiveTransportFlowSimulator (passive_transport_flow_simulators.py:14-136): ENO3 advection + diffusion of the
    primary field against the oracle's restatement of the same two composite kernels, three steps."""
    import torch

    from oracle import flow as oflow
    from oracle import stencils as ost
    from sopht_b200.simulator import PassiveTransportFlowSimulator

    grid_dim, grid, field_type = case
    real_t = real_t_of(precision)
    sim = PassiveTransportFlowSimulator(kinematic_viscosity=1e-2, grid_dim=grid_dim, grid_size=grid, x_range=1.0,
                                        real_t=real_t, field_type=field_type)
    primary = rng.standard_normal(tuple(sim.primary_field.shape)).astype(real_t)
    velocity = rng.standard_normal(tuple(sim.velocity_field.shape)).astype(real_t)
    sim.primary_field[...] = torch.from_numpy(primary).cuda()
    sim.velocity_field[...] = torch.from_numpy(velocity).cuda()
    buf = np.zeros(grid, dtype=real_t)
    dx = real_t(1.0 / grid[-1])
    dt_ref = oflow.compute_advection_diffusion_stable_timestep(
        velocity_field=velocity, velocity_magnitude_field=buf, grid_dim=grid_dim, dx=dx, cfl=0.1,
        kinematic_viscosity=1e-2, real_t=real_t) * 0.5
    assert sim.compute_stable_timestep(dt_prefac=0.5) == pytest.approx(dt_ref, rel=1e-6 if precision == "single" else 1e-13)
    adv = (ost.advection_timestep_euler_forward_conservative_eno3_vector if field_type == "vector"
           else ost.advection_timestep_euler_forward_conservative_eno3)
    dif = ost.diffusion_timestep_euler_forward_vector if field_type == "vector" else ost.diffusion_timestep_euler_forward
    for _ in range(3):
        sim.time_step(dt=dt_ref)
        adv(primary, buf, velocity, real_t(dt_ref / dx))
        dif(primary, buf, real_t(1e-2 * dt_ref / dx / dx))
    assert sim.time == pytest.approx(3 * dt_ref)
    assert rel_l2(sim.primary_field.cpu().numpy(), primary) < REL_L2_TOL[precision]
    with pytest.raises(ValueError, match="Invalid field type"):
        PassiveTransportFlowSimulator(1e-2, 3, (8, 8, 8), 1.0, field_type="tensor")
    with pytest.raises(ValueError, match="vector 2D fields not supported"):
        PassiveTransportFlowSimulator(1e-2, 2, (8, 8), 1.0, field_type="vector")


def test_create_unbounded_flow_simulator_factories():
    from sopht_b200.simulator import (
        UnboundedNavierStokesFlowSimulator2D,
        UnboundedNavierStokesFlowSimulator3D,
        create_unbounded_flow_simulator_2d,
        create_unbounded_flow_simulator_3d,
    )

    s2 = create_unbounded_flow_simulator_2d(grid_size=(16, 32), x_range=1.0, kinematic_viscosity=1e-2,
                                            flow_type="navier_stokes_with_forcing")
    assert isinstance(s2, UnboundedNavierStokesFlowSimulator2D) and s2.with_forcing
    s3 = create_unbounded_flow_simulator_3d(grid_size=(8, 16, 32), x_range=1.0, kinematic_viscosity=1e-2,
                                            flow_type="navier_stokes", with_free_stream_flow=True)
    assert isinstance(s3, UnboundedNavierStokesFlowSimulator3D) and not s3.with_forcing and s3.with_free_stream_flow
    for make in (create_unbounded_flow_simulator_2d, create_unbounded_flow_simulator_3d):
        with pytest.raises(ValueError, match="Invalid flow type given"):
            make(grid_size=(8, 8, 8)[: 2 if make is create_unbounded_flow_simulator_2d else 3], x_range=1.0,
                 kinematic_viscosity=1e-2, flow_type="passive_scalar")
