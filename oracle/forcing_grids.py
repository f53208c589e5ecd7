"""CPU restatement (numpy) of the reference's rigid-body and Cosserat-rod forcing grids - TEST INFRASTRUCTURE ONLY.

Follows sopht/simulator/immersed_body/rigid_body/rigid_body_forcing_grids.py line by line (cited below); the reference
module itself cannot be imported here (it imports pyelastica, absent from this image), so parity is pinned on the
closed-form answers of the reference's own tests (tests/test_simulator/test_immersed_body/rigid_body/
test_rigid_body_forcing_grids.py and cosserat_rod/test_cosserat_rod_forcing_grids.py), which tests/test_forcing_grids.py
and tests/test_rod_forcing_grids.py mirror for both this file and the CUDA path.
"""

import numpy as np


def cylinder_2d_kinematics(body, local):
    """:28-55 -> (position, velocity, global_frame_relative_position), all (2, N)."""
    q = body.director_collection[:, :, 0]
    g = np.dot(q[:2, :2].T, local)
    pos = body.position_collection[:2] + g
    omega_z = q[2, 2] * body.omega_collection[2, 0]
    vel = np.stack([body.velocity_collection[0] - omega_z * g[1], body.velocity_collection[1] + omega_z * g[0]])
    return pos, vel, g


def cylinder_2d_transfer(body, g, forcing):
    """:57-78 -> (forces (3, 1), torques (3, 1))."""
    forces, torques = np.zeros((3, 1)), np.zeros((3, 1))
    forces[:2] = -np.sum(forcing, axis=1).reshape(-1, 1)
    torques[2] = body.director_collection[2, 2, 0] * np.sum(-g[0] * forcing[1] + g[1] * forcing[0])
    return forces, torques


def rigid_3d_kinematics(body, local=None, global_rel=None):
    """:128-149 (local given) or :291-300 (sphere: global offsets stored) -> (position, velocity, g), (3, N)."""
    q = body.director_collection[:, :, 0]
    g = np.dot(q.T, local) if local is not None else global_rel
    pos = body.position_collection + g
    omega_g = np.dot(q.T, body.omega_collection)
    vel = body.velocity_collection + np.cross(np.broadcast_to(omega_g, g.shape), g, axis=0)
    return pos, vel, g


def rigid_3d_transfer(body, g, forcing):
    """:151-169."""
    forces = -np.sum(forcing, axis=1).reshape(-1, 1)
    torques = -np.dot(body.director_collection[:, :, 0], np.sum(np.cross(g, forcing, axis=0), axis=1).reshape(-1, 1))
    return forces, torques


# --------------------------------------------------------------------------------------------------------------
# Cosserat-rod forcing grids: sopht/simulator/immersed_body/cosserat_rod/cosserat_rod_forcing_grids.py. `rod` is any
# object with pyelastica's rod attribute names (n_elems, position_collection (3, n+1), velocity_collection (3, n+1),
# omega_collection (3, n), director_collection (3, 3, n), mass (n+1), radius (n), lengths (n), tangents (3, n)).
# The four pyelastica helpers the reference imports (elastica 0.3.x, absent from this image) are restated from their
# published definitions; the reference's own tests pin two of them (test_cosserat_rod_forcing_grids.py:35-71).
# --------------------------------------------------------------------------------------------------------------


def node_to_element_velocity(mass, node_velocity):
    """elastica.interaction._node_to_element_velocity: mass-weighted mean of the two end nodes."""
    return (mass[1:] * node_velocity[:, 1:] + mass[:-1] * node_velocity[:, :-1]) / (mass[1:] + mass[:-1])


def elements_to_nodes_inplace(vector_in_element_frame, vector_in_node_frame):
    """elastica.interaction._elements_to_nodes_inplace: half of every element's vector to each end node."""
    vector_in_node_frame[:, :-1] += 0.5 * vector_in_element_frame
    vector_in_node_frame[:, 1:] += 0.5 * vector_in_element_frame


def batch_matvec(mats, vecs):
    return np.einsum("ijk,jk->ik", mats, vecs)


def batch_cross(a, b):
    return np.cross(a, b, axis=0)


def rod_nodal_kinematics(rod, dim):
    """:25-33."""
    return rod.position_collection[:dim].copy(), rod.velocity_collection[:dim].copy()


def rod_nodal_transfer(rod, dim, forcing):
    """:35-73 -> (forces (3, n+1), torques (3, n), moment_arm (3, n))."""
    n = rod.n_elems
    forces, torques = np.zeros((3, n + 1)), np.zeros((3, n))
    forces[:dim] = -forcing
    arm = (rod.position_collection[..., 1:] - rod.position_collection[..., :-1]) / 2.0
    torques[...] = batch_cross(arm, (forces[..., 1:] - forces[..., :-1]) / 2.0)
    torques[..., -1] += np.cross(arm[..., -1], forces[..., -1]) / 2.0
    torques[..., 0] -= np.cross(arm[..., 0], forces[..., 0]) / 2.0
    torques[...] = batch_matvec(rod.director_collection, torques)
    return forces, torques, arm


def rod_element_centric_kinematics(rod, dim):
    """:96-109."""
    pos = (rod.position_collection[:dim, 1:] + rod.position_collection[:dim, :-1]) / 2.0
    return pos, node_to_element_velocity(rod.mass, rod.velocity_collection)[:dim]


def rod_element_centric_transfer(rod, dim, forcing):
    """:111-124 (torques are not touched by the reference: returned as zeros)."""
    n = rod.n_elems
    forces = np.zeros((3, n + 1))
    forces[:dim, 1:] -= 0.5 * forcing
    forces[:dim, :-1] -= 0.5 * forcing
    return forces, np.zeros((3, n))


def rod_edge_kinematics(rod):
    """:179-237 -> (position (2, 3n), velocity (2, 3n), moment_arm (3, n)); columns: centres, left, right edges."""
    n = rod.n_elems
    z_vector = np.repeat(np.array([0, 0, 1.0]).reshape(3, 1), n, axis=-1)
    elem_pos = 0.5 * (rod.position_collection[..., 1:] + rod.position_collection[..., :-1])
    arm = batch_cross(z_vector, rod.tangents) * rod.radius
    pos = np.concatenate([elem_pos[:2], (elem_pos + arm)[:2], (elem_pos - arm)[:2]], axis=1)
    elem_vel = node_to_element_velocity(rod.mass, rod.velocity_collection)
    omega = batch_matvec(np.transpose(rod.director_collection, (1, 0, 2)), rod.omega_collection)
    vel = np.concatenate([elem_vel[:2], (elem_vel + batch_cross(omega, arm))[:2],
                          (elem_vel + batch_cross(omega, -arm))[:2]], axis=1)
    return pos, vel, arm


def rod_edge_transfer(rod, arm, forcing):
    """:239-284."""
    n = rod.n_elems
    forces, torques = np.zeros((3, n + 1)), np.zeros((3, n))
    forces[:2, 1:] -= 0.5 * forcing[:, :n]
    forces[:2, :-1] -= 0.5 * forcing[:, :n]
    left, right = np.zeros((3, n)), np.zeros((3, n))
    left[:2] = -forcing[:, n : 2 * n]
    torques += batch_cross(arm, left)
    right[:2] = -forcing[:, 2 * n : 3 * n]
    torques += batch_cross(-arm, right)
    elements_to_nodes_inplace(left + right, forces)
    return forces, batch_matvec(rod.director_collection, torques)


def rod_surface_layout(rod, surface_grid_density_for_largest_element, with_cap=False):
    """:318-372 and _update_surface_grid_point_for_caps (:505-589) ->
    (surface_grid_points (n), grid_point_radius_ratio (N), surface_point_rotation_angle_list)."""
    n = rod.n_elems
    radius = np.asarray(rod.radius)
    points = np.rint(radius / np.max(radius) * surface_grid_density_for_largest_element).astype(int)
    points[np.where(points < 3)[0]] = 1
    ratio = np.ones(points.sum())
    angles = [np.linspace(0, 2 * np.pi, points[i], endpoint=False) if points[i] > 1 else np.array([])
              for i in range(n)]
    if with_cap:
        for end in [0, -1]:
            end_radius = radius[end]
            if points[end] > 1:
                radial_spacing = end_radius * (2.0 * np.pi / points[end])
                radial_density = max(int(end_radius // radial_spacing), 1)
            else:
                radial_density = 0
            inner_points = np.linspace(1, points[end], radial_density, endpoint=False).astype(int)
            points[end] += inner_points.sum()
            idx = 0 if end == 0 else ratio.shape[0]
            ratio = np.insert(ratio, idx, np.ones(inner_points.sum()))
            inner_ratio = np.linspace(0, end_radius, radial_density, endpoint=False) / end_radius
            start = points.cumsum()[end] - inner_points.sum()
            for ring, count in enumerate(inner_points):
                ratio[start : start + count] = inner_ratio[ring]
                start += count
            if points[end] > 1:
                ring_angles = list(np.linspace(0, 2 * np.pi, points[end] - inner_points.sum(), endpoint=False))
                for count in inner_points:
                    ring_angles.extend(np.linspace(0, 2 * np.pi, count, endpoint=False).tolist())
                angles[end] = np.array(ring_angles)
    return points, ratio, angles


def rod_surface_tables(points, angles):
    """:358-385 -> (start_idx (n), end_idx (n), local_frame_surface_points (3, N))."""
    end_idx = np.cumsum(points)
    start_idx = end_idx - points
    local = np.zeros((3, int(points.sum())))
    for i, a in enumerate(angles):
        if a.size:
            local[0, start_idx[i] : end_idx[i]] = np.cos(a)
            local[1, start_idx[i] : end_idx[i]] = np.sin(a)
    return start_idx, end_idx, local


def rod_surface_kinematics(rod, points, ratio, local):
    """:408-467 -> (position (3, N), velocity (3, N), moment_arm (3, N))."""
    elem_of = np.repeat(np.arange(rod.n_elems), points)
    elem_pos = 0.5 * (rod.position_collection[..., 1:] + rod.position_collection[..., :-1])
    qt = np.transpose(rod.director_collection, (1, 0, 2))
    arm = (np.asarray(rod.radius)[elem_of] * ratio) * batch_matvec(qt[:, :, elem_of], local)
    pos = elem_pos[:, elem_of] + arm
    elem_vel = node_to_element_velocity(rod.mass, rod.velocity_collection)
    omega = batch_matvec(qt, rod.omega_collection)
    vel = elem_vel[:, elem_of] + batch_cross(omega[:, elem_of], arm)
    return pos, vel, arm


def rod_surface_transfer(rod, points, arm, forcing):
    """:469-497."""
    n = rod.n_elems
    end_idx = np.cumsum(points)
    start_idx = end_idx - points
    forces, torques = np.zeros((3, n + 1)), np.zeros((3, n))
    for i in range(n):
        on_elem = np.sum(forcing[:, start_idx[i] : end_idx[i]], axis=1)
        forces[:, i] -= 0.5 * on_elem
        forces[:, i + 1] -= 0.5 * on_elem
    lag_torque = batch_cross(arm, -forcing)
    for i in range(n):
        torques[:, i] = rod.director_collection[:, :, i] @ np.sum(lag_torque[:, start_idx[i] : end_idx[i]], axis=1)
    return forces, torques
