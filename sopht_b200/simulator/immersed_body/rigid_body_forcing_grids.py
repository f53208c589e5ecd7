"""Rigid-body forcing grids with their (dim, N) fields on the device.

Same class names, constructor arguments, attributes and method names as
sopht/simulator/immersed_body/rigid_body/rigid_body_forcing_grids.py:11-300 and
immersed_body_forcing_grid.py:9-56. `rigid_body` is any object with pyelastica's rigid-body attribute names
(`position_collection` (3, 1), `velocity_collection` (3, 1), `omega_collection` (3, 1), `director_collection`
(3, 3, 1), `radius`, `length`); `RigidBodyState` is a stand-in with exactly those. `position_field` and
`velocity_field` are float64 CUDA tensors that go straight into VirtualBoundaryForcing /
EulerianLagrangianGridCommunicator; per coupled step only the body's 21 scalars travel host -> device (as kernel
arguments) and 6 sums device -> host (csrc/forcing_grid.cu).
"""

from __future__ import annotations

import ctypes
import logging
from dataclasses import dataclass, field
from typing import Any

import numpy as np
import torch

from sopht_b200 import _lib

logger = logging.getLogger(__name__)


@dataclass
class RigidBodyState:
    """The attributes of ea.Cylinder / ea.Sphere the forcing grids read (rigid_body_forcing_grids.py:28-55, 291-300)."""

    position_collection: np.ndarray = field(default_factory=lambda: np.zeros((3, 1)))
    velocity_collection: np.ndarray = field(default_factory=lambda: np.zeros((3, 1)))
    omega_collection: np.ndarray = field(default_factory=lambda: np.zeros((3, 1)))
    director_collection: np.ndarray = field(default_factory=lambda: np.eye(3).reshape(3, 3, 1).copy())
    radius: float = 1.0
    length: float = 1.0


class ImmersedBodyForcingGrid:
    """immersed_body_forcing_grid.py:9-56."""

    def __init__(self, grid_dim: int, num_lag_nodes: int) -> None:
        if not torch.cuda.is_available():
            msg = "sopht_b200 forcing grids need a CUDA device (no CPU fallback)"
            raise _lib.SophtLibraryError(msg)
        self.grid_dim = grid_dim
        self.num_lag_nodes = num_lag_nodes
        dev = torch.device("cuda", torch.cuda.current_device())
        self.position_field = torch.zeros((grid_dim, num_lag_nodes), dtype=torch.float64, device=dev)
        self.velocity_field = torch.zeros_like(self.position_field)
        if grid_dim == 2:
            logger.warning("2D body forcing grid generated, this assumes the body moves in XY plane!")

    def compute_lag_grid_position_field(self) -> None:
        raise NotImplementedError

    def compute_lag_grid_velocity_field(self) -> None:
        raise NotImplementedError

    def transfer_forcing_from_grid_to_body(self, body_flow_forces, body_flow_torques, lag_grid_forcing_field) -> None:
        raise NotImplementedError

    def get_maximum_lagrangian_grid_spacing(self) -> float:
        raise NotImplementedError


class _RigidGrid(ImmersedBodyForcingGrid):
    """Kinematics and force transfer shared by the 2-D and 3-D rigid grids (one kernel each)."""

    _uses_local_frame = True

    def _init_rigid(self, rigid_body: Any) -> None:
        self.local_frame_relative_position_field = torch.zeros_like(self.position_field)
        self.global_frame_relative_position_field = torch.zeros_like(self.position_field)
        self._body = rigid_body
        self._sums = torch.zeros(6, dtype=torch.float64, device=self.position_field.device)

    def _rotation_and_omega(self) -> tuple[np.ndarray, np.ndarray]:
        raise NotImplementedError

    def _update_kinematics(self) -> None:
        b, d = self._body, self.grid_dim
        rot, omega_g = self._rotation_and_omega()
        raw = _lib.raw_desc
        fl = raw(self.local_frame_relative_position_field) if self._uses_local_frame else None
        fp, fv, fg = raw(self.position_field), raw(self.velocity_field), raw(self.global_frame_relative_position_field)
        com = np.zeros(3)
        vel = np.zeros(3)
        com[:d] = np.asarray(b.position_collection, dtype=np.float64)[:d, 0]
        vel[:d] = np.asarray(b.velocity_collection, dtype=np.float64)[:d, 0]
        _lib.check(_lib.load().sopht_rigid_forcing_grid_kinematics(
            d, ctypes.byref(fp), ctypes.byref(fv), ctypes.byref(fg), ctypes.byref(fl) if fl is not None else None,
            _lib.double_array(rot.reshape(-1)), _lib.double_array(com), _lib.double_array(vel),
            _lib.double_array(omega_g), _lib.current_stream()))

    # the reference updates positions and velocities in two calls; both come out of one kernel here, so either call
    # refreshes both (calling them back to back, as ImmersedBodyFlowInteraction does, launches twice: 2 x 5 us)
    def compute_lag_grid_position_field(self) -> None:
        self._update_kinematics()

    def compute_lag_grid_velocity_field(self) -> None:
        self._update_kinematics()

    def _force_sums(self, lag_grid_forcing_field: torch.Tensor) -> np.ndarray:
        dt = _lib.SOPHT_F32 if lag_grid_forcing_field.dtype == torch.float32 else _lib.SOPHT_F64
        fg, ff = _lib.raw_desc(self.global_frame_relative_position_field), _lib.raw_desc(lag_grid_forcing_field)
        _lib.check(_lib.load().sopht_rigid_forcing_grid_force_sums(
            dt, self.grid_dim, ctypes.byref(fg), ctypes.byref(ff), ctypes.c_void_p(self._sums.data_ptr()),
            _lib.current_stream()))
        return self._sums.cpu().numpy()  # the one device -> host read of the transfer: 6 doubles


class TwoDimensionalCylinderForcingGrid(_RigidGrid):
    """rigid_body_forcing_grids.py:11-82 (cross-section in the XY plane)."""

    def __init__(self, grid_dim: int, num_lag_nodes: int, rigid_body: Any) -> None:
        if grid_dim != 2:
            msg = "Invalid grid dimensions. 2D cylinder forcing grid is only defined for grid_dim=2"
            raise ValueError(msg)
        self.cylinder = rigid_body
        super().__init__(grid_dim, num_lag_nodes)
        self._init_rigid(rigid_body)

    def _rotation_and_omega(self):
        q = np.asarray(self.cylinder.director_collection, dtype=np.float64)[:, :, 0]
        rot = np.eye(3)
        rot[:2, :2] = q[:2, :2].T  # :28-33
        omega_z = q[2, 2] * float(np.asarray(self.cylinder.omega_collection)[2, 0])  # :43-46
        return rot, np.array([0.0, 0.0, omega_z])

    def transfer_forcing_from_grid_to_body(self, body_flow_forces, body_flow_torques, lag_grid_forcing_field) -> None:
        s = self._force_sums(lag_grid_forcing_field)
        body_flow_forces[:2] = -s[:2].reshape(-1, 1)  # :65-66, Newton's third law
        q22 = float(np.asarray(self.cylinder.director_collection)[2, 2, 0])
        # :70-76: Q[2, 2] * sum(-r_x f_y + r_y f_x) = -Q[2, 2] * (sum r x f)_z
        body_flow_torques[2] = q22 * (-s[5])


class CircularCylinderForcingGrid(TwoDimensionalCylinderForcingGrid):
    """rigid_body_forcing_grids.py:84-110."""

    def __init__(self, grid_dim: int, rigid_body: Any, num_forcing_points: int) -> None:
        super().__init__(grid_dim=grid_dim, num_lag_nodes=num_forcing_points, rigid_body=rigid_body)
        dtheta = 2.0 * np.pi / self.num_lag_nodes
        theta = np.linspace(0 + dtheta / 2.0, 2.0 * np.pi - dtheta / 2.0, self.num_lag_nodes)
        local = np.stack([self.cylinder.radius * np.cos(theta), self.cylinder.radius * np.sin(theta)])
        self.local_frame_relative_position_field[...] = torch.from_numpy(local).to(self.position_field.device)
        self.compute_lag_grid_position_field()
        self.compute_lag_grid_velocity_field()

    def get_maximum_lagrangian_grid_spacing(self) -> float:
        return self.cylinder.radius * (2.0 * np.pi / self.num_lag_nodes)


class ThreeDimensionalRigidBodyForcingGrid(_RigidGrid):
    """rigid_body_forcing_grids.py:113-173."""

    def __init__(self, grid_dim: int, num_lag_nodes: int, rigid_body: Any) -> None:
        if grid_dim != 3:
            msg = "Invalid grid dimensions. 3D Rigid body forcing grid is only defined for grid_dim=3"
            raise ValueError(msg)
        self.rigid_body = rigid_body
        super().__init__(grid_dim, num_lag_nodes)
        self._init_rigid(rigid_body)

    def _rotation_and_omega(self):
        q = np.asarray(self.rigid_body.director_collection, dtype=np.float64)[:, :, 0]
        omega_g = np.dot(q.T, np.asarray(self.rigid_body.omega_collection, dtype=np.float64))[:, 0]  # :141-144
        return q.T.copy(), omega_g

    def transfer_forcing_from_grid_to_body(self, body_flow_forces, body_flow_torques, lag_grid_forcing_field) -> None:
        s = self._force_sums(lag_grid_forcing_field)
        body_flow_forces[...] = -s[:3].reshape(-1, 1)  # :159
        q = np.asarray(self.rigid_body.director_collection, dtype=np.float64)[:, :, 0]
        body_flow_torques[...] = -np.dot(q, s[3:].reshape(-1, 1))  # :162-168


class OpenEndCircularCylinderForcingGrid(ThreeDimensionalRigidBodyForcingGrid):
    """rigid_body_forcing_grids.py:176-233 (no forcing at the base and the top)."""

    def __init__(self, grid_dim: int, rigid_body: Any, num_forcing_points_along_length: int) -> None:
        self.num_forcing_points_along_length = num_forcing_points_along_length
        circumference = 2 * np.pi * rigid_body.radius
        self.num_forcing_points_along_circumference = int(
            np.ceil(num_forcing_points_along_length * circumference / rigid_body.length))
        n = num_forcing_points_along_length * self.num_forcing_points_along_circumference
        super().__init__(grid_dim=grid_dim, num_lag_nodes=n, rigid_body=rigid_body)
        nc = self.num_forcing_points_along_circumference
        dtheta = 2.0 * np.pi / nc
        theta = np.linspace(0 + dtheta / 2.0, 2.0 * np.pi - dtheta / 2.0, nc)
        length_grid = np.linspace(-0.5 * rigid_body.length, 0.5 * rigid_body.length, num_forcing_points_along_length)
        local = np.zeros((3, n))
        for idx in range(0, n, nc):
            local[0, idx : idx + nc] = rigid_body.radius * np.cos(theta)
            local[1, idx : idx + nc] = rigid_body.radius * np.sin(theta)
            local[2, idx : idx + nc] = length_grid[idx // nc]
        self.local_frame_relative_position_field[...] = torch.from_numpy(local).to(self.position_field.device)
        self.compute_lag_grid_position_field()
        self.compute_lag_grid_velocity_field()

    def get_maximum_lagrangian_grid_spacing(self) -> float:
        return max(self.rigid_body.radius * (2.0 * np.pi / self.num_forcing_points_along_circumference),
                   self.rigid_body.length / self.num_forcing_points_along_length)


class RectangularPlane:
    """derived_rigid_bodies.py:12-69: a rigid rectangle with imposed motion (its dynamics are not tracked): directors
    (tangent along the length, normal x tangent, normal), centre at `origin`, zero velocity and angular velocity."""

    def __init__(self, origin: np.ndarray, plane_normal: np.ndarray, plane_tangent_along_length: np.ndarray,
                 plane_length: float, plane_breadth: float) -> None:
        logger.warning("Initialising rectangular plane object: tracking its dynamics in pyelastica is not supported, "
                       "do not add the plane to the pyelastica simulator!")
        self.n_elems = 1
        self.length = plane_length
        self.breadth = plane_breadth
        normal = np.asarray(plane_normal, dtype=np.float64).reshape(3, 1)
        tangent = np.asarray(plane_tangent_along_length, dtype=np.float64).reshape(3, 1)
        binormal = np.cross(normal, tangent, axis=0)
        self.director_collection = np.zeros((3, 3, 1))
        self.director_collection[0, ...] = tangent / np.linalg.norm(tangent)
        self.director_collection[1, ...] = binormal / np.linalg.norm(binormal)
        self.director_collection[2, ...] = normal / np.linalg.norm(normal)
        self.position_collection = np.asarray(origin, dtype=np.float64).reshape(3, 1)
        self.velocity_collection = np.zeros((3, 1))
        self.omega_collection = np.zeros((3, 1))


class RectangularPlaneForcingGrid(ThreeDimensionalRigidBodyForcingGrid):
    """rigid_body_forcing_grids.py:303-351: a length x breadth lattice in the plane's d1-d2 frame."""

    def __init__(self, grid_dim: int, rigid_body: Any, num_forcing_points_along_length: int) -> None:
        self.num_forcing_points_along_length = num_forcing_points_along_length
        self.num_forcing_points_along_breadth = int(
            num_forcing_points_along_length * rigid_body.breadth / rigid_body.length)
        self.grid_spacing = rigid_body.length / self.num_forcing_points_along_length
        n = self.num_forcing_points_along_length * self.num_forcing_points_along_breadth
        super().__init__(grid_dim=grid_dim, num_lag_nodes=n, rigid_body=rigid_body)
        along_length = np.linspace(-0.5 * rigid_body.length, 0.5 * rigid_body.length,
                                   self.num_forcing_points_along_length)
        along_breadth = np.linspace(-0.5 * rigid_body.breadth, 0.5 * rigid_body.breadth,
                                    self.num_forcing_points_along_breadth)
        length_grid, breadth_grid = np.meshgrid(along_length, along_breadth)
        local = np.stack([length_grid.reshape(-1), breadth_grid.reshape(-1), np.zeros(n)])
        self.local_frame_relative_position_field[...] = torch.from_numpy(local).to(self.position_field.device)
        self.compute_lag_grid_position_field()
        self.compute_lag_grid_velocity_field()

    def get_maximum_lagrangian_grid_spacing(self) -> float:
        return self.grid_spacing


class SphereForcingGrid(ThreeDimensionalRigidBodyForcingGrid):
    """rigid_body_forcing_grids.py:236-300: latitude rings with equal point density; the local frame is redundant
    for a sphere, positions are centre + the stored global-frame offsets (:291-300)."""

    _uses_local_frame = False

    def __init__(self, grid_dim: int, rigid_body: Any, num_forcing_points_along_equator: int) -> None:
        self.num_forcing_points_along_equator = num_forcing_points_along_equator
        polar = np.linspace(0, np.pi, num_forcing_points_along_equator // 2)
        per_latitude = np.rint(num_forcing_points_along_equator * np.sin(polar)).astype(int) + 1
        super().__init__(grid_dim=grid_dim, num_lag_nodes=int(sum(per_latitude)), rigid_body=rigid_body)
        xs, ys, zs = [], [], []
        for count, angle in zip(per_latitude, polar):
            az = np.linspace(0.0, 2 * np.pi, count, endpoint=False)
            xs.append(rigid_body.radius * np.sin(angle) * np.cos(az))
            ys.append(rigid_body.radius * np.sin(angle) * np.sin(az))
            zs.append(rigid_body.radius * np.cos(angle) * np.ones(count))
        rel = np.stack([np.concatenate(xs), np.concatenate(ys), np.concatenate(zs)])
        self.global_frame_relative_position_field[...] = torch.from_numpy(rel).to(self.position_field.device)
        self.compute_lag_grid_position_field()
        self.compute_lag_grid_velocity_field()

    def get_maximum_lagrangian_grid_spacing(self) -> float:
        return self.rigid_body.radius * (2 * np.pi / self.num_forcing_points_along_equator)
