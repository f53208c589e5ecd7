"""Cosserat-rod forcing grids with their (dim, N) fields on the device.

Same class names, constructor arguments, attributes and method names as
sopht/simulator/immersed_body/cosserat_rod/cosserat_rod_forcing_grids.py:10-589. `cosserat_rod` is any object with
pyelastica's rod attribute names (`n_elems`, `position_collection` (3, n+1), `velocity_collection` (3, n+1),
`omega_collection` (3, n), `director_collection` (3, 3, n), `mass` (n+1), `radius` (n), `lengths` (n), `tangents`
(3, n)); `CosseratRodState` is a stand-in with exactly those (pyelastica is not in this image).

Per call the rod's state (7 (n+1) + 16 n doubles: 7 KB for the 40-element rod of flow_past_rod_case.py) goes host ->
device as one pinned block; `position_field`, `velocity_field` and `moment_arm` are float64 CUDA tensors that feed
VirtualBoundaryForcing directly, and the transfer brings back the rod's (3, n+1) forces and (3, n) torques as one
block (csrc/rod_forcing_grid.cu). The (dim, N_lag) arrays never cross the bus.
"""

from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import Any

import numpy as np
import torch

from sopht_b200 import _lib

from .rigid_body_forcing_grids import ImmersedBodyForcingGrid

_NODAL, _ELEMENT, _EDGE, _SURFACE = 0, 1, 2, 3


@dataclass
class CosseratRodState:
    """The attributes of ea.CosseratRod the forcing grids read."""

    n_elems: int
    position_collection: np.ndarray
    velocity_collection: np.ndarray
    omega_collection: np.ndarray
    director_collection: np.ndarray
    mass: np.ndarray
    radius: np.ndarray
    lengths: np.ndarray = field(default=None)  # type: ignore[assignment]
    tangents: np.ndarray = field(default=None)  # type: ignore[assignment]

    def __post_init__(self) -> None:
        if self.lengths is None or self.tangents is None:
            self.update_geometry()

    def update_geometry(self) -> None:
        """lengths and tangents from the node positions (pyelastica keeps them current every rod step)."""
        edges = self.position_collection[:, 1:] - self.position_collection[:, :-1]
        self.lengths = np.linalg.norm(edges, axis=0)
        self.tangents = edges / self.lengths

    @classmethod
    def straight_rod(cls, n_elements: int, start: Any, direction: Any, normal: Any, base_length: float,
                     base_radius: Any, density: float = 1e3) -> CosseratRodState:
        """A straight rod at rest laid out like ea.CosseratRod.straight_rod: uniform elements along `direction`,
        directors (d1, d2, d3) = (normal, direction x normal, direction), element mass = density * pi r^2 l split
        half and half between its two nodes."""
        start = np.asarray(start, dtype=np.float64)
        d3 = np.asarray(direction, dtype=np.float64)
        d3 = d3 / np.linalg.norm(d3)
        d1 = np.asarray(normal, dtype=np.float64)
        d1 = d1 / np.linalg.norm(d1)
        s = np.linspace(0.0, base_length, n_elements + 1)
        position = start.reshape(3, 1) + d3.reshape(3, 1) * s
        directors = np.zeros((3, 3, n_elements))
        directors[0], directors[1], directors[2] = d1.reshape(3, 1), np.cross(d3, d1).reshape(3, 1), d3.reshape(3, 1)
        radius = np.broadcast_to(np.asarray(base_radius, dtype=np.float64), (n_elements,)).copy()
        lengths = np.full(n_elements, base_length / n_elements)
        elem_mass = density * np.pi * radius**2 * lengths
        mass = np.zeros(n_elements + 1)
        mass[:-1] += 0.5 * elem_mass
        mass[1:] += 0.5 * elem_mass
        return cls(n_elems=n_elements, position_collection=position, velocity_collection=np.zeros((3, n_elements + 1)),
                   omega_collection=np.zeros((3, n_elements)), director_collection=directors, mass=mass, radius=radius)


class _RodGrid(ImmersedBodyForcingGrid):
    """State upload, kinematics launch and force / torque read-back shared by the four rod grids."""

    _kind = -1

    def _init_rod(self, cosserat_rod: Any, moment_arm_columns: int | None) -> None:
        self.cosserat_rod = cosserat_rod
        n = int(cosserat_rod.n_elems)
        self._n = n
        dev = self.position_field.device
        count = int(_lib.load().sopht_rod_state_doubles(n))
        self._state_host = torch.empty(count, dtype=torch.float64).pin_memory()
        self._state_np = self._state_host.numpy()
        self._state_dev = torch.empty(count, dtype=torch.float64, device=dev)
        self._uploaded = torch.cuda.Event()
        self._upload_pending = False
        self._out = torch.zeros(6 * n + 3, dtype=torch.float64, device=dev)
        self._out_host = torch.empty(6 * n + 3, dtype=torch.float64).pin_memory()
        self.moment_arm = (torch.zeros((3, moment_arm_columns), dtype=torch.float64, device=dev)
                           if moment_arm_columns is not None else None)
        self._surface_tables: tuple[Any, Any, Any, Any] = (None, None, None, None)

    def _upload_rod_state(self) -> None:
        rod, n, buf = self.cosserat_rod, self._n, self._state_np
        if self._upload_pending:  # the previous async copy must have read the pinned block before it is rewritten
            self._uploaded.synchronize()
        parts = (
            (rod.position_collection, 3 * (n + 1)),
            (rod.velocity_collection, 3 * (n + 1)),
            (rod.mass, n + 1),
            (rod.director_collection, 9 * n),
            (rod.omega_collection, 3 * n),
            (rod.radius, n),
            (rod.tangents, 3 * n),
        )
        at = 0
        for values, count in parts:
            arr = np.asarray(values, dtype=np.float64)
            if arr.size != count:
                msg = f"rod attribute of {arr.size} values where {count} were expected (n_elems = {n})"
                raise ValueError(msg)
            buf[at : at + count] = arr.reshape(-1)
            at += count
        self._state_dev.copy_(self._state_host, non_blocking=True)
        self._uploaded.record()
        self._upload_pending = True

    def _update_kinematics(self) -> None:
        self._upload_rod_state()
        raw = _lib.raw_desc
        fp, fv = raw(self.position_field), raw(self.velocity_field)
        arm = ctypes.byref(raw(self.moment_arm)) if self._kind in (_EDGE, _SURFACE) else None
        node_element, local_points, radius_ratio, _ = self._surface_tables
        ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
        _lib.check(_lib.load().sopht_rod_forcing_grid_kinematics(
            self._kind, self.grid_dim, self._n, ctypes.c_void_p(self._state_dev.data_ptr()), ctypes.byref(fp),
            ctypes.byref(fv), arm, ptr(node_element), ptr(local_points), ptr(radius_ratio), _lib.current_stream()))

    # positions and velocities come out of one kernel, so either call refreshes both (as in the rigid grids)
    def compute_lag_grid_position_field(self) -> None:
        self._update_kinematics()

    def compute_lag_grid_velocity_field(self) -> None:
        self._update_kinematics()

    def _transfer(self, lag_grid_forcing_field: torch.Tensor) -> tuple[np.ndarray, np.ndarray]:
        if not isinstance(lag_grid_forcing_field, torch.Tensor):
            lag_grid_forcing_field = torch.as_tensor(np.ascontiguousarray(lag_grid_forcing_field),
                                                     device=self.position_field.device)
        self._upload_rod_state()  # the reference reads the rod's current positions / directors here as well
        dt = _lib.SOPHT_F32 if lag_grid_forcing_field.dtype == torch.float32 else _lib.SOPHT_F64
        ff = _lib.raw_desc(lag_grid_forcing_field)
        arm = ctypes.byref(_lib.raw_desc(self.moment_arm)) if self.moment_arm is not None else None
        start = self._surface_tables[3]
        _lib.check(_lib.load().sopht_rod_forcing_grid_transfer(
            dt, self._kind, self.grid_dim, self._n, ctypes.c_void_p(self._state_dev.data_ptr()), ctypes.byref(ff), arm,
            ctypes.c_void_p(start.data_ptr()) if start is not None else None,
            ctypes.c_void_p(self._out.data_ptr()), _lib.current_stream()))
        self._out_host.copy_(self._out, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the one device -> host read: 3 (n+1) + 3 n doubles
        out = self._out_host.numpy()
        n = self._n
        return out[: 3 * (n + 1)].reshape(3, n + 1), out[3 * (n + 1) :].reshape(3, n)

    def get_maximum_lagrangian_grid_spacing(self) -> float:
        return np.amax(self.cosserat_rod.lengths)


class CosseratRodNodalForcingGrid(_RodGrid):
    """cosserat_rod_forcing_grids.py:10-79: one forcing point on every rod node."""

    _kind = _NODAL

    def __init__(self, grid_dim: int, cosserat_rod: Any) -> None:
        super().__init__(grid_dim, cosserat_rod.n_elems + 1)
        self._init_rod(cosserat_rod, cosserat_rod.n_elems)
        self.compute_lag_grid_position_field()
        self.compute_lag_grid_velocity_field()

    def transfer_forcing_from_grid_to_body(self, body_flow_forces, body_flow_torques, lag_grid_forcing_field) -> None:
        forces, torques = self._transfer(lag_grid_forcing_field)
        # :42 writes the first grid_dim rows only; the torque arithmetic (:45-73) sees zeros in the others, which is
        # what pyelastica's external-force buffers hold there
        body_flow_forces[: self.grid_dim] = forces[: self.grid_dim]
        body_flow_torques[...] = torques


class CosseratRodElementCentricForcingGrid(_RodGrid):
    """cosserat_rod_forcing_grids.py:82-131: one forcing point on every element centre."""

    _kind = _ELEMENT

    def __init__(self, grid_dim: int, cosserat_rod: Any) -> None:
        super().__init__(grid_dim, cosserat_rod.n_elems)
        self._init_rod(cosserat_rod, None)
        self.compute_lag_grid_position_field()
        self.compute_lag_grid_velocity_field()

    def transfer_forcing_from_grid_to_body(self, body_flow_forces, body_flow_torques, lag_grid_forcing_field) -> None:
        forces, _ = self._transfer(lag_grid_forcing_field)
        body_flow_forces[...] = forces  # :118-120; the torques are not touched (:122-123)


class CosseratRodEdgeForcingGrid(_RodGrid):
    """cosserat_rod_forcing_grids.py:134-289: element centres plus the two edges at +- r (z x t), 2-D only; for
    tapered / thick rods."""

    _kind = _EDGE

    def __init__(self, grid_dim: int, cosserat_rod: Any) -> None:
        if grid_dim != 2:
            msg = "Invalid grid dimensions. Cosserat rod edge forcing grid is only defined for grid_dim=2"
            raise ValueError(msg)
        n = cosserat_rod.n_elems
        super().__init__(grid_dim, n + 2 * n)
        self._init_rod(cosserat_rod, n)
        self.z_vector = np.repeat(np.array([0, 0, 1.0]).reshape(3, 1), n, axis=-1)
        self.start_idx_elems = 0
        self.end_idx_elems = self.start_idx_elems + n
        self.start_idx_left_edge_nodes = self.end_idx_elems
        self.end_idx_left_edge_nodes = self.start_idx_left_edge_nodes + n
        self.start_idx_right_edge_nodes = self.end_idx_left_edge_nodes
        self.end_idx_right_edge_nodes = self.start_idx_right_edge_nodes + n
        self.compute_lag_grid_position_field()
        self.compute_lag_grid_velocity_field()

    def transfer_forcing_from_grid_to_body(self, body_flow_forces, body_flow_torques, lag_grid_forcing_field) -> None:
        forces, torques = self._transfer(lag_grid_forcing_field)
        body_flow_forces[...] = forces
        body_flow_torques[...] = torques


class CosseratRodSurfaceForcingGrid(_RodGrid):
    """cosserat_rod_forcing_grids.py:292-589: rings of forcing points on the surface of every element (3-D only), the
    ring density scaled with the element radius, optional concentric cap rings on the two end elements."""

    _kind = _SURFACE

    def __init__(self, grid_dim: int, cosserat_rod: Any, surface_grid_density_for_largest_element: int,
                 with_cap: bool = False) -> None:
        if grid_dim != 3:
            msg = "Invalid grid dimensions. Cosserat rod surface forcing grid is only defined for grid_dim=3"
            raise ValueError(msg)
        self.cosserat_rod = cosserat_rod
        self.n_elems = cosserat_rod.n_elems
        self.surface_grid_density_for_largest_element = surface_grid_density_for_largest_element
        self.with_cap = with_cap
        radius = np.asarray(cosserat_rod.radius, dtype=np.float64)
        # ring sizes follow the local radius; fewer than 3 points collapse to one point on the centre line (:318-327)
        self.surface_grid_points = np.rint(
            radius / np.max(radius) * surface_grid_density_for_largest_element).astype(int)
        self.surface_grid_points[self.surface_grid_points < 3] = 1
        self.grid_point_radius_ratio = np.ones(self.surface_grid_points.sum())
        self.surface_point_rotation_angle_list = [
            np.linspace(0, 2 * np.pi, count, endpoint=False) if count > 1 else np.array([])
            for count in self.surface_grid_points]
        if self.with_cap:
            self._update_surface_grid_point_for_caps()

        num_lag_nodes = int(self.surface_grid_points.sum())
        super().__init__(grid_dim, num_lag_nodes)
        self._init_rod(cosserat_rod, num_lag_nodes)
        self.end_idx = np.cumsum(self.surface_grid_points).astype(int)
        self.start_idx = (self.end_idx - self.surface_grid_points).astype(int)
        local = np.zeros((3, num_lag_nodes))
        for i, angles in enumerate(self.surface_point_rotation_angle_list):
            if angles.size:
                local[0, self.start_idx[i] : self.end_idx[i]] = np.cos(angles)
                local[1, self.start_idx[i] : self.end_idx[i]] = np.sin(angles)
        self.local_frame_surface_points = local
        dev = self.position_field.device
        node_element = np.repeat(np.arange(self.n_elems, dtype=np.int32), self.surface_grid_points)
        windows = np.concatenate([[0], self.end_idx]).astype(np.int32)
        self._surface_tables = (
            torch.from_numpy(node_element).to(dev),
            torch.from_numpy(np.ascontiguousarray(local[:2])).to(dev),
            torch.from_numpy(np.ascontiguousarray(self.grid_point_radius_ratio, dtype=np.float64)).to(dev),
            torch.from_numpy(windows).to(dev),
        )
        self.compute_lag_grid_position_field()
        self.compute_lag_grid_velocity_field()

    def transfer_forcing_from_grid_to_body(self, body_flow_forces, body_flow_torques, lag_grid_forcing_field) -> None:
        forces, torques = self._transfer(lag_grid_forcing_field)
        body_flow_forces[...] = forces
        body_flow_torques[...] = torques

    def get_maximum_lagrangian_grid_spacing(self) -> float:
        grid_angular_spacing = 2 * np.pi / self.surface_grid_density_for_largest_element
        return max(np.amax(self.cosserat_rod.lengths), np.amax(self.cosserat_rod.radius) * grid_angular_spacing)

    def _update_surface_grid_point_for_caps(self) -> None:
        """Concentric rings on the two end faces (:505-589): the ring count follows from how many angular spacings
        fit in the end radius, the ring sizes grow linearly outwards, and the inner points are stored after the
        lateral ring inside the end element's window."""
        for end in (0, -1):
            end_radius = self.cosserat_rod.radius[end]
            lateral = self.surface_grid_points[end]
            if lateral > 1:
                radial_spacing = end_radius * (2.0 * np.pi / lateral)
                rings = max(int(end_radius // radial_spacing), 1)
            else:
                rings = 0  # a single centre point already: nothing to add
            ring_sizes = np.linspace(1, lateral, rings, endpoint=False).astype(int)
            extra = int(ring_sizes.sum())
            self.surface_grid_points[end] += extra
            at = 0 if end == 0 else self.grid_point_radius_ratio.shape[0]
            self.grid_point_radius_ratio = np.insert(self.grid_point_radius_ratio, at, np.ones(extra))
            ring_ratio = np.linspace(0, end_radius, rings, endpoint=False) / end_radius
            first = self.surface_grid_points.cumsum()[end] - extra
            for ring, size in enumerate(ring_sizes):
                self.grid_point_radius_ratio[first : first + size] = ring_ratio[ring]
                first += size
            if self.surface_grid_points[end] > 1:
                angles = list(np.linspace(0, 2 * np.pi, lateral, endpoint=False))
                for size in ring_sizes:
                    angles.extend(np.linspace(0, 2 * np.pi, size, endpoint=False).tolist())
                self.surface_point_rotation_angle_list[end] = np.array(angles)
