"""Base flow simulator: grid, coordinates, time (same interface as
sopht/simulator/flow/flow_simulators.py:9-133). Fields are torch CUDA tensors."""

from __future__ import annotations

import logging
from abc import abstractmethod

import numpy as np
import torch

from sopht_b200 import _lib

logger = logging.getLogger(__name__)


class FlowSimulator:
    """Base class: all flow simulators share this interface."""

    def __init__(
        self,
        grid_dim: int,
        grid_size,
        x_range: float,
        real_t: type = np.float32,
        num_threads: int = 1,
        time: float = 0.0,
    ) -> None:
        if grid_dim not in [2, 3]:
            msg = "Invalid grid dimensions. Supported values include 2 and 3."
            raise ValueError(msg)
        self.grid_dim = grid_dim
        self.grid_size = tuple(grid_size)
        self.x_range = x_range
        self.real_t = real_t
        self.num_threads = num_threads
        self.time = time
        self._torch_t = _lib.torch_dtype(real_t)
        if not torch.cuda.is_available():
            msg = "sopht_b200 flow simulators need a CUDA device (no CPU fallback)"
            raise _lib.SophtLibraryError(msg)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._init_domain()
        self._init_fields()
        self._compile_kernels()
        self._finalise_flow_time_step()

    def _init_domain(self) -> None:
        """Cell-centred coordinates dx/2 ... range - dx/2 per axis (flow_simulators.py:50-95)."""
        grid_size_x = self.grid_size[-1]
        self.dx = self.real_t(self.x_range / grid_size_x)
        shift = self.dx / 2.0
        self._axis_coords = []  # numpy, array-axis order (z, y, x) / (y, x)
        ranges = []
        for n in self.grid_size:
            rng = self.x_range * n / grid_size_x
            ranges.append(rng)
            self._axis_coords.append(np.linspace(shift, rng - shift, n).astype(self.real_t))
        if self.grid_dim == 2:
            self.y_range = ranges[0]
        else:
            self.z_range, self.y_range = ranges[0], ranges[1]
        # component order x, y[, z] (reverse of the meshgrid order)
        mesh = np.flipud(np.array(np.meshgrid(*self._axis_coords, indexing="ij")))
        self.position_field = torch.from_numpy(np.ascontiguousarray(mesh)).to(self.device)
        logger.info(
            "\n==================================================\n%dD flow domain initialized with "
            "x_range %s, grid %s\nPlease initialize bodies within these bounds!"
            "\n==================================================",
            self.grid_dim, self.x_range, self.grid_size,
        )

    def _zeros(self, *shape) -> torch.Tensor:
        return torch.zeros(*shape, dtype=self._torch_t, device=self.device)

    @abstractmethod
    def _init_fields(self) -> None: ...

    @abstractmethod
    def _compile_kernels(self) -> None: ...

    @abstractmethod
    def _finalise_flow_time_step(self) -> None: ...

    @abstractmethod
    def compute_stable_timestep(self) -> float: ...

    def _update_simulator_time(self, dt: float) -> None:
        self.time += dt

    def time_step(self, dt: float, **kwargs) -> None:
        """Final simulator time step."""
        self._flow_time_step(dt=dt, **kwargs)
        self._update_simulator_time(dt=dt)

    def graph_time_step(self, dt: float, before=None, extra_state=(), **kwargs):
        """The time step with FIXED dt (and free stream ...) captured into a CUDA graph: returns a callable that
        replays it with one launch and advances `time`. `before` (optional callable, e.g. the immersed-body
        interaction of the coupled step) is captured in front of the flow step; `extra_state` lists tensors `before`
        updates in place. For the small grids of BASELINE configs[0-2], where the 15-40 launches of a step are issued
        more slowly than the device executes them. Not part of the reference API (it has no device)."""
        from sopht_b200 import _lib

        def body() -> None:
            if before is not None:
                before()
            self._flow_time_step(dt=dt, **kwargs)

        state = [getattr(self, n) for n in ("vorticity_field", "velocity_field", "eul_grid_forcing_field",
                                            "primary_field", "stream_func_field", "buffer_vector_field")
                 if isinstance(getattr(self, n, None), torch.Tensor)]
        absmax = getattr(self, "_vel_absmax", None)
        if isinstance(absmax, torch.Tensor):
            state.append(absmax)
        graph = _lib.StepGraph(body, list(state) + list(extra_state))

        def replay() -> None:
            graph()
            if hasattr(self, "_vel_absmax_version"):
                self._vel_absmax_version = (self.velocity_field.data_ptr(), self.velocity_field._version)
            self._update_simulator_time(dt=dt)

        replay.graph = graph
        return replay
