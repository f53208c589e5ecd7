"""CPU restatement (numpy) of the reference's rigid-body forcing grids - TEST INFRASTRUCTURE ONLY.

Follows sopht/simulator/immersed_body/rigid_body/rigid_body_forcing_grids.py line by line (cited below); the reference
module itself cannot be imported here (it imports pyelastica, absent from this image), so parity is pinned on the
closed-form answers of the reference's own tests (tests/test_simulator/test_immersed_body/rigid_body/
test_rigid_body_forcing_grids.py), which tests/test_forcing_grids_gpu.py mirrors for both this file and the CUDA path.
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
