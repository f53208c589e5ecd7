"""Cosserat-rod forcing grids and the body <-> flow interaction objects.

The closed-form checks of the reference's tests/test_simulator/test_immersed_body/cosserat_rod/
test_cosserat_rod_forcing_grids.py (and test_immersed_body_flow_interaction.py, test_flow_forces.py,
test_cosserat_rod_flow_interaction.py), applied to the CPU restatement `oracle/forcing_grids.py` (always) and to the
CUDA classes (GPU), plus CUDA vs restatement on random bent, tapered, moving rods. pyelastica is absent from this
image: `CosseratRodState.straight_rod` lays the mock rod out like ea.CosseratRod.straight_rod."""

import logging

import numpy as np
import pytest

from sopht_b200.simulator.immersed_body.cosserat_rod_forcing_grids import CosseratRodState


def mock_straight_rod(n_elems, base_radius=0.05):
    """test_cosserat_rod_forcing_grids.py:9-32: along (1, 1, 1), node velocities 1..n+1, element omegas 1..n."""
    rod = CosseratRodState.straight_rod(n_elems, start=np.zeros(3), direction=np.array([1.0, 1.0, 1.0]),
                                        normal=np.array([0.0, -1.0, 1.0]), base_length=1.0, base_radius=base_radius)
    rod.velocity_collection[...] = np.linspace(1, n_elems + 1, n_elems + 1)
    rod.omega_collection[...] = np.linspace(1, n_elems, n_elems)
    return rod


def random_rod(n_elems, seed, planar=False):
    """A bent, tapered rod with random node velocities, element spins and material frames."""
    rng = np.random.default_rng(seed)
    steps = 0.05 + 0.02 * rng.random((3, n_elems))
    if planar:
        steps[2] = 0.0
    position = np.concatenate([rng.standard_normal((3, 1)), np.zeros((3, n_elems))], axis=1)
    position[:, 1:] = position[:, :1] + np.cumsum(steps, axis=1)
    directors = np.zeros((3, 3, n_elems))
    for e in range(n_elems):
        q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
        directors[:, :, e] = q
    return CosseratRodState(
        n_elems=n_elems, position_collection=position, velocity_collection=rng.standard_normal((3, n_elems + 1)),
        omega_collection=rng.standard_normal((3, n_elems)), director_collection=directors,
        mass=0.5 + rng.random(n_elems + 1), radius=np.linspace(0.05, 0.01, n_elems) * (1 + 0.1 * rng.random(n_elems)))


def _end_corrected_linspace(v, n_elems):
    """Element velocity of the mock rod: linear in between, (v0 + 2 v1) / 3 at the half-mass ends."""
    out = np.linspace(0.5 * (v[0] + v[1]), 0.5 * (v[-1] + v[-2]), n_elems)
    out[0] = (v[0] + 2 * v[1]) / 3
    out[-1] = (v[-1] + 2 * v[-2]) / 3
    return out


def _surface_reference_layout(n_elems, density, radius, with_cap):
    """The independent construction of the reference test's MockSurfaceForcingGrid (:381-460): explicit formulas for
    the ring counts and cap rings instead of the linspace / insert bookkeeping of the class under test."""
    counts, angles, ratios = [], [], []
    for i in range(n_elems):
        k = round(radius[i] / np.max(radius) * density)
        if k < 3:
            counts.append(1), angles.append(np.array([])), ratios.append(np.ones(1))
            continue
        a = [np.array([2 * np.pi / k * j for j in range(k)])]
        r = [np.ones(k)]
        if with_cap and i in (0, n_elems - 1):
            rings = max(int(radius[i] // (radius[i] * 2.0 * np.pi / k)), 1)
            for j in range(rings):
                size = (k // rings - 1) * j + 1
                a.append(np.linspace(0, 2 * np.pi, size, endpoint=False))
                r.append(np.full(size, j / rings))
        angles.append(np.concatenate(a)), ratios.append(np.concatenate(r)), counts.append(angles[-1].size)
    return np.array(counts), angles, ratios


# --------------------------------------------------------------------------------------------------------------
# CPU: the restatement against the reference tests' closed forms
# --------------------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("n_elems", [8, 16])
def test_oracle_pyelastica_helpers(n_elems):
    """test_pyelastica__node_to_element_velocity_func_validity / __elements_to_nodes_inplace (:35-71)."""
    from oracle import forcing_grids as ofg

    rod = mock_straight_rod(n_elems)
    v = rod.velocity_collection
    correct = 0.5 * (v[:, 1:] + v[:, :-1])
    correct[:, 0] = (v[:, 0] + 2 * v[:, 1]) / 3
    correct[:, -1] = (v[:, -1] + 2 * v[:, -2]) / 3
    np.testing.assert_allclose(ofg.node_to_element_velocity(rod.mass, v), correct)
    vec = np.random.default_rng(42).random((3, n_elems))
    out, ref = np.zeros((3, n_elems + 1)), np.zeros((3, n_elems + 1))
    ref[:, 1:] += 0.5 * vec
    ref[:, :-1] += 0.5 * vec
    ofg.elements_to_nodes_inplace(vec, out)
    np.testing.assert_allclose(out, ref)


@pytest.mark.parametrize("grid_dim", [2, 3])
@pytest.mark.parametrize("n_elems", [8, 16])
def test_oracle_nodal_and_element_centric_grids(grid_dim, n_elems):
    from oracle import forcing_grids as ofg

    rod = mock_straight_rod(n_elems)
    uniform = np.linspace(1.0, grid_dim, grid_dim).reshape(-1, 1)
    # nodal (:74-140)
    pos, vel = ofg.rod_nodal_kinematics(rod, grid_dim)
    np.testing.assert_allclose(pos, rod.position_collection[:grid_dim])
    np.testing.assert_allclose(vel, rod.velocity_collection[:grid_dim])
    forces, torques, arm = ofg.rod_nodal_transfer(rod, grid_dim, np.tile(uniform, (1, n_elems + 1)))
    _check_nodal_transfer(rod, grid_dim, forces, torques)
    # element centric (:155-231)
    pos, vel = ofg.rod_element_centric_kinematics(rod, grid_dim)
    _check_element_centric_kinematics(rod, grid_dim, pos, vel)
    forces, torques = ofg.rod_element_centric_transfer(rod, grid_dim, np.tile(uniform, (1, n_elems)))
    _check_half_weighted_forces(forces, grid_dim, uniform, per_element=1)
    np.testing.assert_allclose(torques, 0.0)


def _check_nodal_transfer(rod, grid_dim, forces, torques):
    n = rod.n_elems
    uniform = np.linspace(1.0, grid_dim, grid_dim).reshape(-1, 1)
    correct_forces = np.zeros((3, n + 1))
    correct_forces[:grid_dim] = -uniform
    np.testing.assert_allclose(forces, correct_forces)
    arm = (rod.position_collection[..., 1:] - rod.position_collection[..., :-1]) / 2.0
    correct = np.zeros((3, n))  # uniform loading: only the end corrections survive (:121-140)
    correct[..., -1] += rod.director_collection[..., -1] @ (np.cross(arm[..., -1], correct_forces[..., -1]) / 2.0)
    correct[..., 0] -= rod.director_collection[..., 0] @ (np.cross(arm[..., 0], correct_forces[..., 0]) / 2.0)
    np.testing.assert_allclose(torques, correct, atol=1e-14)


def _check_element_centric_kinematics(rod, grid_dim, pos, vel):
    n = rod.n_elems
    start = np.mean(rod.position_collection[..., :2], axis=1)
    end = np.mean(rod.position_collection[..., -2:], axis=1)
    for axis in range(grid_dim):
        np.testing.assert_allclose(pos[axis], np.linspace(start[axis], end[axis], n), atol=1e-15)
        np.testing.assert_allclose(vel[axis], _end_corrected_linspace(rod.velocity_collection[axis], n))


def _check_half_weighted_forces(forces, grid_dim, uniform, per_element):
    correct = np.zeros_like(forces)
    correct[:grid_dim] = -per_element * uniform
    correct[:grid_dim, (0, -1)] *= 0.5
    np.testing.assert_allclose(forces, correct)


def _check_edge_kinematics(rod, pos, vel, arm):
    """test_rod_edge_grid_grid_kinematics (:300-395), all three node groups."""
    n = rod.n_elems
    tangent = np.ones(3) / np.sqrt(3.0)
    normal = np.cross(np.array([0, 0, 1.0]), tangent)
    correct_arm = rod.radius * normal.reshape(3, 1)
    np.testing.assert_allclose(arm, correct_arm, atol=1e-16)
    start = np.mean(rod.position_collection[..., :2], axis=1)
    end = np.mean(rod.position_collection[..., -2:], axis=1)
    omega_cross_arm = np.zeros((3, n))
    for i in range(n):
        omega_lab = rod.director_collection[:, :, 0].T @ rod.omega_collection[:, i]
        omega_cross_arm[:, i] = np.cross(omega_lab, correct_arm[:, i])
    for axis in range(2):
        centre = np.linspace(start[axis], end[axis], n)
        np.testing.assert_allclose(pos[axis, :n], centre, atol=1e-15)
        np.testing.assert_allclose(pos[axis, n : 2 * n], centre + correct_arm[axis], atol=1e-15)
        np.testing.assert_allclose(pos[axis, 2 * n :], centre - correct_arm[axis], atol=1e-15)
        elem_vel = _end_corrected_linspace(rod.velocity_collection[axis], n)
        np.testing.assert_allclose(vel[axis, :n], elem_vel)
        np.testing.assert_allclose(vel[axis, n : 2 * n], elem_vel + omega_cross_arm[axis])
        np.testing.assert_allclose(vel[axis, 2 * n :], elem_vel - omega_cross_arm[axis])


@pytest.mark.parametrize("n_elems", [8, 16])
def test_oracle_edge_grid(n_elems):
    from oracle import forcing_grids as ofg

    rod = mock_straight_rod(n_elems)
    pos, vel, arm = ofg.rod_edge_kinematics(rod)
    _check_edge_kinematics(rod, pos, vel, arm)
    uniform = np.array([[1.0], [2.0]])
    forces, torques = ofg.rod_edge_transfer(rod, arm, np.tile(uniform, (1, 3 * n_elems)))
    _check_half_weighted_forces(forces, 2, uniform, per_element=3)  # centre + left + right (:398-424)
    np.testing.assert_allclose(torques, 0.0, atol=1e-15)


def _check_surface_kinematics(rod, counts, ratios, local, pos, vel, arm):
    """test_rod_surface_grid_grid_kinematics (:561-634): node by node."""
    start = np.cumsum(counts) - counts
    correct_arm, correct_pos, correct_vel = np.zeros_like(pos), np.zeros_like(pos), np.zeros_like(pos)
    for i in range(rod.n_elems):
        centre = 0.5 * (rod.position_collection[:, i] + rod.position_collection[:, i + 1])
        qt = rod.director_collection[:, :, i].T
        elem_vel = (rod.velocity_collection[:, i] * rod.mass[i] + rod.velocity_collection[:, i + 1] * rod.mass[i + 1]
                    ) / (rod.mass[i] + rod.mass[i + 1])
        omega_lab = qt @ rod.omega_collection[:, i]
        for j in range(counts[i]):
            g = start[i] + j
            correct_arm[:, g] = rod.radius[i] * ratios[i][j] * qt @ local[:, g]
            correct_pos[:, g] = centre + correct_arm[:, g]
            correct_vel[:, g] = elem_vel + np.cross(omega_lab, correct_arm[:, g])
    np.testing.assert_allclose(arm, correct_arm, atol=1e-14)
    np.testing.assert_allclose(pos, correct_pos, atol=1e-14)
    np.testing.assert_allclose(vel, correct_vel, atol=1e-12)


SURFACE_CASES = [(n, d, t, c) for n in (8, 16) for d in (16, 12, 8, 4) for t in (1, 2, 5, 10) for c in (True, False)]


@pytest.mark.parametrize(("n_elems", "density", "taper", "with_cap"), SURFACE_CASES)
def test_oracle_surface_grid(n_elems, density, taper, with_cap):
    """test_rod_surface_grid_setup / _grid_kinematics / _force_transfer (:480-700)."""
    from oracle import forcing_grids as ofg

    radius = np.linspace(1, 1 / taper, n_elems)
    rod = mock_straight_rod(n_elems, base_radius=radius)
    points, ratio, angles = ofg.rod_surface_layout(rod, density, with_cap)
    counts, ref_angles, ref_ratios = _surface_reference_layout(n_elems, density, radius, with_cap)
    np.testing.assert_array_equal(points, counts)
    for i in range(n_elems):
        np.testing.assert_allclose(angles[i], ref_angles[i], atol=1e-11)
    np.testing.assert_allclose(ratio, np.concatenate(ref_ratios), atol=1e-14)
    start, end, local = ofg.rod_surface_tables(points, angles)
    np.testing.assert_array_equal(end, np.cumsum(counts))
    np.testing.assert_array_equal(start, np.cumsum(counts) - counts)
    pos, vel, arm = ofg.rod_surface_kinematics(rod, points, ratio, local)
    _check_surface_kinematics(rod, counts, ref_ratios, local, pos, vel, arm)
    uniform = np.array([[1.0], [2.0], [3.0]])
    forces, torques = ofg.rod_surface_transfer(rod, points, arm, np.tile(uniform, (1, int(points.sum()))))
    _check_surface_uniform_transfer(counts, forces, torques)


def _check_surface_uniform_transfer(counts, forces, torques):
    uniform = np.array([1.0, 2.0, 3.0])
    correct = np.zeros_like(forces)
    for i, k in enumerate(counts):
        correct[:, i] -= 0.5 * uniform * k
        correct[:, i + 1] -= 0.5 * uniform * k
    np.testing.assert_allclose(forces, correct)
    np.testing.assert_allclose(torques, 0.0, atol=1e-11)  # the rings are symmetric about the centre line


def test_flow_forces():
    """test_flow_forces.py:20-32."""
    from sopht_b200.simulator import FlowForces

    class Interactor:
        body_flow_forces = 0.0
        body_flow_torques = 0.0

        def compute_flow_forces_and_torques(self):
            self.body_flow_forces, self.body_flow_torques = 1.0, 2.0

    class Rod:
        external_forces = 3.0
        external_torques = 4.0

    interactor, rod = Interactor(), Rod()
    forcing = FlowForces(body_flow_interactor=interactor)
    assert forcing.body_flow_interactor is interactor
    forcing.apply_forces(system=rod)
    assert rod.external_forces == 4.0 and rod.external_torques == 6.0


# --------------------------------------------------------------------------------------------------------------
# GPU: the CUDA classes against the same closed forms and against the restatement
# --------------------------------------------------------------------------------------------------------------


def _np(t):
    return t.cpu().numpy()


def _transfer(grid, n_elems, forcing, dtype="float64"):
    import torch

    forces, torques = np.zeros((3, n_elems + 1)), np.zeros((3, n_elems))
    grid.transfer_forcing_from_grid_to_body(
        body_flow_forces=forces, body_flow_torques=torques,
        lag_grid_forcing_field=torch.from_numpy(np.ascontiguousarray(forcing.astype(dtype))).cuda())
    return forces, torques


@pytest.mark.gpu
@pytest.mark.parametrize("grid_dim", [2, 3])
@pytest.mark.parametrize("n_elems", [8, 16])
def test_cuda_nodal_and_element_centric_grids(grid_dim, n_elems):
    from sopht_b200.simulator import CosseratRodElementCentricForcingGrid, CosseratRodNodalForcingGrid

    rod = mock_straight_rod(n_elems)
    uniform = np.linspace(1.0, grid_dim, grid_dim).reshape(-1, 1)
    grid = CosseratRodNodalForcingGrid(grid_dim=grid_dim, cosserat_rod=rod)
    assert grid.cosserat_rod is rod and grid.num_lag_nodes == n_elems + 1
    assert tuple(grid.position_field.shape) == (grid_dim, n_elems + 1) == tuple(grid.velocity_field.shape)
    np.testing.assert_allclose(_np(grid.position_field), rod.position_collection[:grid_dim])
    np.testing.assert_allclose(_np(grid.velocity_field), rod.velocity_collection[:grid_dim])
    forces, torques = _transfer(grid, n_elems, np.tile(uniform, (1, n_elems + 1)))
    _check_nodal_transfer(rod, grid_dim, forces, torques)
    assert grid.get_maximum_lagrangian_grid_spacing() == np.amax(rod.lengths)

    grid = CosseratRodElementCentricForcingGrid(grid_dim=grid_dim, cosserat_rod=rod)
    assert grid.cosserat_rod is rod and grid.num_lag_nodes == n_elems
    assert tuple(grid.position_field.shape) == (grid_dim, n_elems) == tuple(grid.velocity_field.shape)
    _check_element_centric_kinematics(rod, grid_dim, _np(grid.position_field), _np(grid.velocity_field))
    forces, torques = _transfer(grid, n_elems, np.tile(uniform, (1, n_elems)))
    _check_half_weighted_forces(forces, grid_dim, uniform, per_element=1)
    np.testing.assert_allclose(torques, 0.0)
    assert grid.get_maximum_lagrangian_grid_spacing() == np.amax(rod.lengths)


@pytest.mark.gpu
@pytest.mark.parametrize("n_elems", [8, 16])
def test_cuda_edge_grid(n_elems):
    from sopht_b200.simulator import CosseratRodEdgeForcingGrid

    rod = mock_straight_rod(n_elems)
    for bad_dim in (0, 1, 3, 4):
        with pytest.raises(ValueError, match="Cosserat rod edge forcing grid is only defined for grid_dim=2"):
            CosseratRodEdgeForcingGrid(grid_dim=bad_dim, cosserat_rod=rod)
    grid = CosseratRodEdgeForcingGrid(grid_dim=2, cosserat_rod=rod)
    assert grid.cosserat_rod is rod and grid.num_lag_nodes == 3 * n_elems
    assert tuple(grid.position_field.shape) == (2, 3 * n_elems) == tuple(grid.velocity_field.shape)
    assert tuple(grid.moment_arm.shape) == (3, n_elems)
    assert (grid.start_idx_elems, grid.end_idx_elems) == (0, n_elems)
    assert (grid.start_idx_left_edge_nodes, grid.end_idx_left_edge_nodes) == (n_elems, 2 * n_elems)
    assert (grid.start_idx_right_edge_nodes, grid.end_idx_right_edge_nodes) == (2 * n_elems, 3 * n_elems)
    _check_edge_kinematics(rod, _np(grid.position_field), _np(grid.velocity_field), _np(grid.moment_arm))
    uniform = np.array([[1.0], [2.0]])
    forces, torques = _transfer(grid, n_elems, np.tile(uniform, (1, 3 * n_elems)))
    _check_half_weighted_forces(forces, 2, uniform, per_element=3)
    np.testing.assert_allclose(torques, 0.0, atol=1e-15)
    assert grid.get_maximum_lagrangian_grid_spacing() == np.amax(rod.lengths)


@pytest.mark.gpu
@pytest.mark.parametrize(("n_elems", "density", "taper", "with_cap"), SURFACE_CASES)
def test_cuda_surface_grid(n_elems, density, taper, with_cap):
    from sopht_b200.simulator import CosseratRodSurfaceForcingGrid

    radius = np.linspace(1, 1 / taper, n_elems)
    rod = mock_straight_rod(n_elems, base_radius=radius)
    grid = CosseratRodSurfaceForcingGrid(grid_dim=3, cosserat_rod=rod,
                                         surface_grid_density_for_largest_element=density, with_cap=with_cap)
    counts, ref_angles, ref_ratios = _surface_reference_layout(n_elems, density, radius, with_cap)
    assert grid.cosserat_rod is rod and grid.num_lag_nodes == counts.sum()
    assert tuple(grid.position_field.shape) == (3, counts.sum()) == tuple(grid.moment_arm.shape)
    np.testing.assert_array_equal(grid.surface_grid_points, counts)
    for i in range(n_elems):
        np.testing.assert_allclose(grid.surface_point_rotation_angle_list[i], ref_angles[i], atol=1e-11)
    np.testing.assert_array_equal(grid.end_idx, np.cumsum(counts))
    np.testing.assert_array_equal(grid.start_idx, np.cumsum(counts) - counts)
    _check_surface_kinematics(rod, counts, ref_ratios, grid.local_frame_surface_points, _np(grid.position_field),
                              _np(grid.velocity_field), _np(grid.moment_arm))
    uniform = np.array([[1.0], [2.0], [3.0]])
    forces, torques = _transfer(grid, n_elems, np.tile(uniform, (1, grid.num_lag_nodes)))
    _check_surface_uniform_transfer(counts, forces, torques)
    spacing = max(np.amax(rod.lengths), np.max(radius) * (2 * np.pi / density))
    np.testing.assert_allclose(grid.get_maximum_lagrangian_grid_spacing(), spacing)


@pytest.mark.gpu
def test_cuda_surface_grid_dimension():
    from sopht_b200.simulator import CosseratRodSurfaceForcingGrid

    for bad_dim in (0, 1, 2, 4):
        with pytest.raises(ValueError, match="Cosserat rod surface forcing grid is only defined for grid_dim=3"):
            CosseratRodSurfaceForcingGrid(grid_dim=bad_dim, cosserat_rod=mock_straight_rod(8),
                                          surface_grid_density_for_largest_element=1)


@pytest.mark.gpu
@pytest.mark.parametrize("forcing_dtype", ["float32", "float64"])
def test_cuda_rod_grids_match_restatement_on_random_rods(forcing_dtype):
    """Bent, tapered rods with random frames, velocities and forcing; the rod then moves and the grids follow."""
    from oracle import forcing_grids as ofg
    from sopht_b200.simulator import (
        CosseratRodEdgeForcingGrid,
        CosseratRodElementCentricForcingGrid,
        CosseratRodNodalForcingGrid,
        CosseratRodSurfaceForcingGrid,
    )

    rng = np.random.default_rng(5)
    tol = dict(rtol=1e-12, atol=1e-12)
    for n_elems, seed in ((1, 1), (7, 2), (40, 3), (133, 4)):
        for dim in (2, 3):
            rod = random_rod(n_elems, seed, planar=dim == 2)
            nodal = CosseratRodNodalForcingGrid(grid_dim=dim, cosserat_rod=rod)
            centric = CosseratRodElementCentricForcingGrid(grid_dim=dim, cosserat_rod=rod)
            edge = CosseratRodEdgeForcingGrid(grid_dim=2, cosserat_rod=rod) if dim == 2 else None
            surface = (CosseratRodSurfaceForcingGrid(grid_dim=3, cosserat_rod=rod, with_cap=True,
                                                     surface_grid_density_for_largest_element=12)
                       if dim == 3 else None)
            # the rod moves: every grid follows on its next update
            rod.position_collection += 0.1 * rng.standard_normal(rod.position_collection.shape) * (
                np.array([1.0, 1.0, 0.0 if dim == 2 else 1.0]).reshape(3, 1))
            rod.velocity_collection *= 1.5
            rod.update_geometry()
            for grid in (nodal, centric, edge, surface):
                if grid is not None:
                    grid.compute_lag_grid_position_field()
                    grid.compute_lag_grid_velocity_field()

            pos, vel = ofg.rod_nodal_kinematics(rod, dim)
            np.testing.assert_allclose(_np(nodal.position_field), pos, **tol)
            np.testing.assert_allclose(_np(nodal.velocity_field), vel, **tol)
            f = rng.standard_normal((dim, n_elems + 1)).astype(forcing_dtype)
            forces, torques = _transfer(nodal, n_elems, f, forcing_dtype)
            rf, rt, arm = ofg.rod_nodal_transfer(rod, dim, f.astype(np.float64))
            np.testing.assert_allclose(forces, rf, **tol)
            np.testing.assert_allclose(torques, rt, **tol)
            np.testing.assert_allclose(_np(nodal.moment_arm), arm, **tol)

            pos, vel = ofg.rod_element_centric_kinematics(rod, dim)
            np.testing.assert_allclose(_np(centric.position_field), pos, **tol)
            np.testing.assert_allclose(_np(centric.velocity_field), vel, **tol)
            f = rng.standard_normal((dim, n_elems)).astype(forcing_dtype)
            forces, torques = _transfer(centric, n_elems, f, forcing_dtype)
            rf, _ = ofg.rod_element_centric_transfer(rod, dim, f.astype(np.float64))
            np.testing.assert_allclose(forces, rf, **tol)
            np.testing.assert_allclose(torques, 0.0)

            if edge is not None:
                pos, vel, arm = ofg.rod_edge_kinematics(rod)
                np.testing.assert_allclose(_np(edge.position_field), pos, **tol)
                np.testing.assert_allclose(_np(edge.velocity_field), vel, **tol)
                np.testing.assert_allclose(_np(edge.moment_arm), arm, **tol)
                f = rng.standard_normal((2, 3 * n_elems)).astype(forcing_dtype)
                forces, torques = _transfer(edge, n_elems, f, forcing_dtype)
                rf, rt = ofg.rod_edge_transfer(rod, arm, f.astype(np.float64))
                np.testing.assert_allclose(forces, rf, **tol)
                np.testing.assert_allclose(torques, rt, **tol)

            if surface is not None:
                points, ratio, angles = ofg.rod_surface_layout(rod, 12, with_cap=True)
                np.testing.assert_array_equal(surface.surface_grid_points, points)
                np.testing.assert_allclose(surface.grid_point_radius_ratio, ratio)
                _, _, local = ofg.rod_surface_tables(points, angles)
                np.testing.assert_allclose(surface.local_frame_surface_points, local, atol=1e-15)
                pos, vel, arm = ofg.rod_surface_kinematics(rod, points, ratio, local)
                np.testing.assert_allclose(_np(surface.position_field), pos, **tol)
                np.testing.assert_allclose(_np(surface.velocity_field), vel, **tol)
                np.testing.assert_allclose(_np(surface.moment_arm), arm, **tol)
                f = rng.standard_normal((3, surface.num_lag_nodes)).astype(forcing_dtype)
                forces, torques = _transfer(surface, n_elems, f, forcing_dtype)
                rf, rt = ofg.rod_surface_transfer(rod, points, arm, f.astype(np.float64))
                np.testing.assert_allclose(forces, rf, rtol=1e-11, atol=1e-11)
                np.testing.assert_allclose(torques, rt, rtol=1e-11, atol=1e-11)
                # Newton's third law on the whole rod
                np.testing.assert_allclose(forces.sum(axis=1), -f.astype(np.float64).sum(axis=1), rtol=1e-10,
                                           atol=1e-10)


def _cylinder_interactor(num_forcing_points=16):
    """mock_2d_cylinder_flow_interactor of test_immersed_body_flow_interaction.py:13-36, fields on the device."""
    import torch

    from sopht_b200.simulator import CircularCylinderForcingGrid, RigidBodyFlowInteraction, RigidBodyState

    body = RigidBodyState(radius=0.5, length=2.0)
    body.position_collection[:, 0] = (4.0, 4.0, 0.0)
    body.velocity_collection[...] = 3.0
    body.omega_collection[2] = 4.0
    velocity = torch.from_numpy(np.random.default_rng(seed=0).random((2, 16, 16))).cuda()
    forcing = torch.zeros_like(velocity)
    dx = body.length / 4.0
    interactor = RigidBodyFlowInteraction(
        rigid_body=body, eul_grid_forcing_field=forcing, eul_grid_velocity_field=velocity,
        virtual_boundary_stiffness_coeff=1.0, virtual_boundary_damping_coeff=1.0, dx=dx, grid_dim=2,
        real_t=np.float64, forcing_grid_cls=CircularCylinderForcingGrid, num_forcing_points=num_forcing_points)
    return interactor, forcing, velocity, dx


@pytest.mark.gpu
@pytest.mark.parametrize("num_forcing_points", [1, 4, 64])
def test_cuda_interactor_resolution_messages(num_forcing_points, caplog):
    """test_immersed_body_interactor_warnings (:39-85): level and text of the resolution message."""
    with caplog.at_level(logging.INFO):
        interactor, _, _, dx = _cylinder_interactor(num_forcing_points)
    spacing = interactor.forcing_grid.get_maximum_lagrangian_grid_spacing()
    bar = "\n" + "=" * 50
    head = f"{bar}\nFor CircularCylinderForcingGrid:\nEulerian grid spacing (dx): {dx}"
    if spacing > 2 * dx:
        body = (f"\nMax Lagrangian grid spacing: {spacing} > 2 * dx\nThe Lagrangian grid of the body is too coarse "
                "relative to\nthe Eulerian grid of the flow, which can lead to unexpected\nconvergence. Please make "
                "the Lagrangian grid finer.")
        level = logging.WARNING
    elif spacing < 0.5 * dx:
        body = (f"\nMax Lagrangian grid spacing: {spacing} < 0.5 * dx\nThe Lagrangian grid of the body is too fine "
                "relative to\nthe Eulerian grid of the flow, which corresponds to redundant\nforcing points. Please "
                "make the Lagrangian grid coarser.")
        level = logging.WARNING
    else:
        body = "\nLagrangian grid is resolved almost the same\nas the Eulerian grid of the flow."
        level = logging.INFO
    name = "sopht_b200.simulator.immersed_body.immersed_body_flow_interaction"
    assert (name, level, head + body + bar) in caplog.record_tuples
    # the penalty coefficients are rescaled by the Lagrangian spacing (immersed_body_flow_interaction.py:84-87)
    assert interactor.virtual_boundary_stiffness_coeff == pytest.approx(spacing)


@pytest.mark.gpu
def test_cuda_interactor_call_and_flow_forces():
    """test_immersed_body_interactor_call_method / _compute_flow_forces_and_torques /
    _get_grid_deviation_error_l2_norm (:88-135)."""
    import torch

    interactor, forcing, velocity, _ = _cylinder_interactor()
    interactor()
    ref_forcing = torch.zeros_like(forcing)
    grid = interactor.forcing_grid
    grid.compute_lag_grid_position_field()
    grid.compute_lag_grid_velocity_field()
    interactor.compute_interaction_forcing(
        eul_grid_forcing_field=ref_forcing, eul_grid_velocity_field=velocity,
        lag_grid_position_field=grid.position_field, lag_grid_velocity_field=grid.velocity_field)
    assert float(forcing.abs().max()) > 0
    np.testing.assert_allclose(_np(ref_forcing), _np(forcing), rtol=1e-12, atol=1e-13)

    interactor.compute_flow_forces_and_torques()
    ref_forces, ref_torques = np.zeros((3, 1)), np.zeros((3, 1))
    grid.compute_lag_grid_position_field()
    grid.compute_lag_grid_velocity_field()
    interactor.compute_interaction_force_on_lag_grid(
        eul_grid_velocity_field=velocity, lag_grid_position_field=grid.position_field,
        lag_grid_velocity_field=grid.velocity_field)
    grid.transfer_forcing_from_grid_to_body(body_flow_forces=ref_forces, body_flow_torques=ref_torques,
                                            lag_grid_forcing_field=interactor.lag_grid_forcing_field)
    assert np.abs(ref_forces).max() > 0
    np.testing.assert_allclose(ref_forces, interactor.body_flow_forces)
    np.testing.assert_allclose(ref_torques, interactor.body_flow_torques)

    interactor.lag_grid_position_mismatch_field[...] = 2.0
    np.testing.assert_allclose(interactor.get_grid_deviation_error_l2_norm(), 2.0 * np.sqrt(2))


@pytest.mark.gpu
@pytest.mark.parametrize("n_elems", [8, 16])
def test_cuda_cosserat_rod_flow_interaction(n_elems):
    """test_cosserat_rod_flow_interaction.py:12-34, then one coupled evaluation against the restatement."""
    import torch

    from oracle import forcing_grids as ofg
    from sopht_b200.simulator import CosseratRodElementCentricForcingGrid, CosseratRodFlowInteraction, FlowForces

    rod = mock_straight_rod(n_elems)
    rod.position_collection[...] = rod.position_collection * 4.0 + 5.0  # inside the 16 x 16 grid of unit cells
    rod.update_geometry()
    velocity = torch.from_numpy(np.random.default_rng(1).random((2, 16, 16))).cuda()
    forcing = torch.zeros_like(velocity)
    interactor = CosseratRodFlowInteraction(
        cosserat_rod=rod, eul_grid_forcing_field=forcing, eul_grid_velocity_field=velocity,
        virtual_boundary_stiffness_coeff=1.0, virtual_boundary_damping_coeff=1.0, dx=1.0, grid_dim=2,
        real_t=np.float64, forcing_grid_cls=CosseratRodElementCentricForcingGrid)
    np.testing.assert_allclose(interactor.body_flow_forces, np.zeros((3, n_elems + 1)))
    np.testing.assert_allclose(interactor.body_flow_torques, np.zeros((3, n_elems)))
    assert isinstance(interactor.forcing_grid, CosseratRodElementCentricForcingGrid)

    class Body:
        external_forces = np.zeros((3, n_elems + 1))
        external_torques = np.zeros((3, n_elems))

    FlowForces(interactor).apply_forces(Body)
    lag_forcing = _np(interactor.lag_grid_forcing_field)
    assert np.abs(lag_forcing).max() > 0
    ref_forces, _ = ofg.rod_element_centric_transfer(rod, 2, lag_forcing)
    np.testing.assert_allclose(Body.external_forces, ref_forces, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(Body.external_torques, 0.0)
