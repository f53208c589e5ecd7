"""Rigid-body forcing grids: the closed-form checks of the reference's
tests/test_simulator/test_immersed_body/rigid_body/test_rigid_body_forcing_grids.py, applied to the CPU restatement
(always) and to the CUDA classes (GPU), plus CUDA vs restatement on random body states with rotated directors."""

import numpy as np
import pytest


def _body(seed=None, radius=0.1, length=1.0):
    from sopht_b200.simulator.immersed_body.rigid_body_forcing_grids import RigidBodyState

    b = RigidBodyState(radius=radius, length=length)
    if seed is None:  # mock_2d_cylinder of the reference test: axis along z, v = 3, omega_z = 4
        b.position_collection[:, 0] = (1.0, 2.0, 0.5)
        b.velocity_collection[...] = 3.0
        b.omega_collection[2] = 4.0
        return b
    rng = np.random.default_rng(seed)
    b.position_collection[...] = rng.standard_normal((3, 1))
    b.velocity_collection[...] = rng.standard_normal((3, 1))
    b.omega_collection[...] = rng.standard_normal((3, 1))
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    b.director_collection[:, :, 0] = q
    return b


def _circle_local(radius, n):
    dtheta = 2.0 * np.pi / n
    theta = np.linspace(dtheta / 2.0, 2.0 * np.pi - dtheta / 2.0, n)
    return np.stack([radius * np.cos(theta), radius * np.sin(theta)]), theta


@pytest.mark.parametrize("n", [8, 16])
def test_oracle_circular_cylinder_closed_forms(n):
    from oracle import forcing_grids as ofg

    b = _body()
    local, theta = _circle_local(b.radius, n)
    pos, vel, g = ofg.cylinder_2d_kinematics(b, local)
    arm = pos - b.position_collection[:2]
    np.testing.assert_allclose(np.linalg.norm(arm, axis=0), b.radius)
    np.testing.assert_allclose((np.arctan2(arm[1], arm[0]) + 2 * np.pi) % (2 * np.pi), theta)
    np.testing.assert_allclose(vel[0], 3.0 - arm[1] * 4.0)
    np.testing.assert_allclose(vel[1], 3.0 + arm[0] * 4.0)
    f = np.zeros((2, n))
    f[0], f[1] = 2.0, 3.0
    forces, torques = ofg.cylinder_2d_transfer(b, g, f)
    np.testing.assert_allclose(forces[:, 0], [-2.0 * n, -3.0 * n, 0.0])
    np.testing.assert_allclose(torques[2, 0], -np.sum(arm[0] * 3.0 - arm[1] * 2.0), atol=1e-11)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [8, 16])
def test_cuda_circular_cylinder_grid(n):
    """test_circular_cylinder_grid_kinematics / _force_transfer / _spacing / _invalid_dim of the reference."""
    import torch

    from sopht_b200.simulator.immersed_body import CircularCylinderForcingGrid

    b = _body()
    with pytest.raises(ValueError, match="2D cylinder forcing grid is only defined for grid_dim=2"):
        CircularCylinderForcingGrid(grid_dim=3, rigid_body=b, num_forcing_points=n)
    grid = CircularCylinderForcingGrid(grid_dim=2, rigid_body=b, num_forcing_points=n)
    assert grid.cylinder is b
    assert tuple(grid.position_field.shape) == (2, n) and tuple(grid.velocity_field.shape) == (2, n)
    pos, vel = grid.position_field.cpu().numpy(), grid.velocity_field.cpu().numpy()
    _, theta = _circle_local(b.radius, n)
    arm = pos - b.position_collection[:2]
    np.testing.assert_allclose(np.linalg.norm(arm, axis=0), b.radius)
    np.testing.assert_allclose((np.arctan2(arm[1], arm[0]) + 2 * np.pi) % (2 * np.pi), theta)
    np.testing.assert_allclose(vel[0], 3.0 - arm[1] * 4.0)
    np.testing.assert_allclose(vel[1], 3.0 + arm[0] * 4.0)
    forces, torques = np.zeros((3, 1)), np.zeros((3, 1))
    f = torch.zeros(2, n, device="cuda", dtype=torch.float64)
    f[0], f[1] = 2.0, 3.0
    grid.transfer_forcing_from_grid_to_body(body_flow_forces=forces, body_flow_torques=torques,
                                            lag_grid_forcing_field=f)
    np.testing.assert_allclose(forces[:, 0], [-2.0 * n, -3.0 * n, 0.0])
    np.testing.assert_allclose(torques[2, 0], -np.sum(arm[0] * 3.0 - arm[1] * 2.0), atol=1e-11)
    assert grid.get_maximum_lagrangian_grid_spacing() == pytest.approx(2 * np.pi * b.radius / n)


@pytest.mark.gpu
@pytest.mark.parametrize("forcing_dtype", ["float32", "float64"])
def test_cuda_rigid_grids_match_restatement_on_random_bodies(forcing_dtype):
    import torch

    from oracle import forcing_grids as ofg
    from sopht_b200.simulator.immersed_body import (
        CircularCylinderForcingGrid,
        OpenEndCircularCylinderForcingGrid,
        SphereForcingGrid,
    )

    tdt = getattr(torch, forcing_dtype)
    rng = np.random.default_rng(3)
    # 2-D cylinder, rotated about z
    b = _body()
    c, s = np.cos(0.7), np.sin(0.7)
    b.director_collection[:, :, 0] = [[c, s, 0.0], [-s, c, 0.0], [0.0, 0.0, 1.0]]
    g2 = CircularCylinderForcingGrid(grid_dim=2, rigid_body=b, num_forcing_points=37)
    pos, vel, rel = ofg.cylinder_2d_kinematics(b, g2.local_frame_relative_position_field.cpu().numpy())
    np.testing.assert_allclose(g2.position_field.cpu().numpy(), pos, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(g2.velocity_field.cpu().numpy(), vel, rtol=1e-13, atol=1e-13)
    f = rng.standard_normal((2, 37)).astype(forcing_dtype)
    forces, torques = np.zeros((3, 1)), np.zeros((3, 1))
    g2.transfer_forcing_from_grid_to_body(forces, torques, torch.from_numpy(f).cuda())
    rf, rt = ofg.cylinder_2d_transfer(b, rel, f.astype(np.float64))
    np.testing.assert_allclose(forces, rf, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(torques, rt, rtol=1e-12, atol=1e-12)
    # 3-D: open-ended cylinder (local frame rotated) and sphere (global offsets), moving bodies
    for cls, kw, uses_local in ((OpenEndCircularCylinderForcingGrid, dict(num_forcing_points_along_length=9), True),
                                (SphereForcingGrid, dict(num_forcing_points_along_equator=24), False)):
        body = _body(seed=11, radius=0.3, length=1.3)
        grid = cls(grid_dim=3, rigid_body=body, **kw)
        n = grid.num_lag_nodes
        body.position_collection[...] += 0.25  # the body moves: the grid follows on the next update
        body.omega_collection[...] *= 1.5
        grid.compute_lag_grid_position_field()
        grid.compute_lag_grid_velocity_field()
        if uses_local:
            pos, vel, rel = ofg.rigid_3d_kinematics(body, local=grid.local_frame_relative_position_field.cpu().numpy())
        else:
            pos, vel, rel = ofg.rigid_3d_kinematics(
                body, global_rel=grid.global_frame_relative_position_field.cpu().numpy())
            np.testing.assert_allclose(np.linalg.norm(rel, axis=0), body.radius)
        np.testing.assert_allclose(grid.position_field.cpu().numpy(), pos, rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(grid.velocity_field.cpu().numpy(), vel, rtol=1e-13, atol=1e-13)
        f = rng.standard_normal((3, n)).astype(forcing_dtype)
        forces, torques = np.zeros((3, 1)), np.zeros((3, 1))
        grid.transfer_forcing_from_grid_to_body(forces, torques, torch.from_numpy(f).cuda().to(tdt))
        rf, rt = ofg.rigid_3d_transfer(body, rel, f.astype(np.float64))
        np.testing.assert_allclose(forces, rf, rtol=1e-11, atol=1e-11)
        np.testing.assert_allclose(torques, rt, rtol=1e-11, atol=1e-11)
    with pytest.raises(ValueError, match="3D Rigid body forcing grid is only defined for grid_dim=3"):
        SphereForcingGrid(grid_dim=2, rigid_body=_body(), num_forcing_points_along_equator=8)


@pytest.mark.gpu
def test_cuda_sphere_grid_drives_virtual_boundary_forcing():
    """The grid's device fields go straight into VirtualBoundaryForcing (no host hop), and the body receives the
    opposite of the summed Lagrangian forcing (flow_past_sphere_case.py:191-212)."""
    import torch

    from sopht_b200.numeric.immersed_boundary_ops import VirtualBoundaryForcing
    from sopht_b200.simulator import UnboundedNavierStokesFlowSimulator3D
    from sopht_b200.simulator.immersed_body import SphereForcingGrid

    sim = UnboundedNavierStokesFlowSimulator3D(grid_size=(32, 32, 64), x_range=1.0, kinematic_viscosity=2e-3,
                                               real_t=np.float32, with_forcing=True, with_free_stream_flow=True)
    body = _body(radius=0.1)
    body.position_collection[:, 0] = (0.3, 0.25, 0.25)
    body.velocity_collection[...] = 0.0
    body.omega_collection[...] = 0.0
    grid = SphereForcingGrid(grid_dim=3, rigid_body=body, num_forcing_points_along_equator=32)
    ds = grid.get_maximum_lagrangian_grid_spacing()
    vb = VirtualBoundaryForcing(virtual_boundary_stiffness_coeff=-5e4 * ds * ds, virtual_boundary_damping_coeff=-20 * ds * ds,
                                grid_dim=3, dx=sim.dx, num_lag_nodes=grid.num_lag_nodes, real_t=np.float32)
    sim.velocity_field[0] = 1.0
    forces, torques = np.zeros((3, 1)), np.zeros((3, 1))
    for _ in range(5):
        dt = sim.compute_stable_timestep(dt_prefac=0.5)
        grid.compute_lag_grid_position_field()
        grid.compute_lag_grid_velocity_field()
        vb.time_step(dt)
        vb.compute_interaction_force_on_eul_and_lag_grid(sim.eul_grid_forcing_field, sim.velocity_field,
                                                         grid.position_field, grid.velocity_field)
        grid.transfer_forcing_from_grid_to_body(forces, torques, vb.lag_grid_forcing_field)
        sim.time_step(dt=dt, free_stream_velocity=[1.0, 0.0, 0.0])
    total = vb.lag_grid_forcing_field.double().sum(dim=1).cpu().numpy()
    np.testing.assert_allclose(forces[:, 0], -total, rtol=1e-6, atol=1e-9)
    assert forces[0, 0] > 0  # the flow pushes the sphere downstream
