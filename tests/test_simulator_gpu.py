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
