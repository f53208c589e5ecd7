"""Unbounded Poisson solve on a z-slab decomposed grid.

    x forward (local z-slab)  ->  all-to-all per component  ->  y fwd, z fwd x G_hat x z inv, y inv
    (local kx-slab)           ->  all-to-all back           ->  x inverse (local z-slab)

`SlabTransposePlan` owns the exchange buffers and the choreography (torch.distributed collectives; NCCL on
the GPUs, gloo in the CPU tests of this host logic); the three compute phases are injected. The product
binding `SlabUnboundedPoissonSolver3D` plugs in the CUDA phases of libsopht_b200
(sopht_poisson_slab_*, include/sopht_b200.h). The reference has no distributed solver; what is computed is
UnboundedPoissonSolverPYFFTW3D.py:111-172 on the global grid.
"""

from __future__ import annotations

import ctypes
import os
from typing import Any, Callable

import numpy as np
import torch
import torch.distributed as dist

from sopht_b200 import _lib
from sopht_b200.numeric.eulerian_grid_ops.poisson_solvers import _reflected_axis

from .slab import SlabPartition


class SlabTransposePlan:
    """Buffers + collectives of the slab-transposed spectral solve.

    Layouts (complex, stored as trailing (re, im) pairs of `dtype`):
      send / recv   (C, P, nz/P, ny, nx/P): chunk q of a component's send goes to rank q; after the
                    exchange a component's recv is (nz, ny, nx/P), z-major, this rank's kx range.
      nyq_local     (C, nz/P, ny): the kx = nx bins of this rank's rows; nyq_all (C, nz, ny) after all-gather.
    """

    def __init__(self, part: SlabPartition, ncomp: int, dtype: torch.dtype, device, group=None) -> None:
        nz, ny, nx = part.grid_size
        p = part.world_size
        if nx % p:
            msg = f"nx = {nx} is not divisible by the number of ranks ({p})"
            raise ValueError(msg)
        self.part, self.ncomp, self.group = part, ncomp, group
        self.nzl, self.nxl = nz // p, nx // p
        shape = (ncomp, p, self.nzl, ny, self.nxl, 2)
        self.send = torch.zeros(shape, dtype=dtype, device=device)
        self.recv = torch.zeros(shape, dtype=dtype, device=device)
        self.nyq_local = torch.zeros((ncomp, self.nzl, ny, 2), dtype=dtype, device=device)
        self.nyq_all = torch.zeros((ncomp, nz, ny, 2), dtype=dtype, device=device)

    def to_kx_slabs(self) -> None:
        """send (z-slab rows, all kx) -> recv (all z, this rank's kx); Nyquist bins gathered on every rank."""
        p = self.part.world_size
        for c in range(self.ncomp):
            if p == 1:
                self.recv[c].copy_(self.send[c])
                self.nyq_all[c].copy_(self.nyq_local[c])
            else:
                dist.all_to_all_single(self.recv[c], self.send[c], group=self.group)
                dist.all_gather_into_tensor(self.nyq_all[c], self.nyq_local[c], group=self.group)

    def to_z_slabs(self) -> None:
        """recv (all z, this rank's kx) -> send (this rank's z rows, all kx as P chunks); Nyquist slice."""
        p, r = self.part.world_size, self.part.rank
        for c in range(self.ncomp):
            if p == 1:
                self.send[c].copy_(self.recv[c])
            else:
                dist.all_to_all_single(self.send[c], self.recv[c], group=self.group)
            self.nyq_local[c].copy_(self.nyq_all[c, r * self.nzl : (r + 1) * self.nzl])

    def solve(self, forward_x: Callable[[], None], middle: Callable[[], None], inverse_x: Callable[[], None]) -> None:
        forward_x()
        with _lib.profile_range("comm.transpose_to_kx_slabs"):
            self.to_kx_slabs()
        middle()
        with _lib.profile_range("comm.transpose_to_z_slabs"):
            self.to_z_slabs()
        inverse_x()


class SlabUnboundedPoissonSolver3D:
    """Distributed counterpart of UnboundedPoissonSolverPYFFTW3D for fp32 power-of-two grids.

    Constructor takes the GLOBAL grid size like the reference class; `vector_field_solve` takes this
    rank's (3, nz/P, ny, nx) slabs (strided views of halo-padded arrays are fine)."""

    def __init__(
        self,
        grid_size_z: int,
        grid_size_y: int,
        grid_size_x: int,
        x_range: float = 1.0,
        num_threads: int = 1,
        real_t: type = np.float32,
        n_components: int = 3,
        group: Any = None,
        peer_exchange: bool | None = None,
        peer_arena: Any = None,
    ) -> None:
        if _lib.dtype_code(real_t) != _lib.SOPHT_F32:
            msg = "the slab-decomposed Poisson solver is implemented for fp32 (power-of-two grids)"
            raise ValueError(msg)
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.grid_size_z, self.grid_size_y, self.grid_size_x = grid_size_z, grid_size_y, grid_size_x
        self.x_range = x_range
        self.y_range = x_range * (grid_size_y / grid_size_x)
        self.z_range = x_range * (grid_size_z / grid_size_x)
        self.dx = real_t(x_range / grid_size_x)
        self.real_t, self.num_threads = real_t, num_threads
        self.part = SlabPartition((grid_size_z, grid_size_y, grid_size_x), world, rank)
        if not torch.cuda.is_available():
            msg = "sopht_b200 Poisson solver needs a CUDA device (no CPU fallback)"
            raise _lib.SophtLibraryError(msg)
        device = torch.device("cuda", torch.cuda.current_device())
        lib = _lib.load()
        nz, ny = grid_size_z, grid_size_y
        self._handle = self._create_handle(lib, n_components, world, rank)
        self._make_buffers(n_components, device, group)
        # transposes fused into the kernels over NVLink peer memory (default on >1 rank; SOPHT_SLAB_PEER=0 or
        # peer_exchange=False keeps the NCCL all-to-all path)
        if peer_exchange is None:
            peer_exchange = world > 1 and os.environ.get("SOPHT_SLAB_PEER", "1") != "0"
        self.peer_exchange = bool(peer_exchange) and world > 1
        self._peer_arena = peer_arena  # its device-side barrier replaces the all-reduce between y inverse and x inverse
        if self.peer_exchange:
            self._open_peer_exchange(lib, world, device, group)
            self.path += "-peer"
            nzl = self.plan.nzl
            self._nyq_gather = torch.zeros((world, n_components, nzl, ny, 2), dtype=torch.float32, device=device)
            self._barrier_flag = torch.zeros(1, dtype=torch.float32, device=device)
            self.plan.send = self.plan.recv = None  # exchange buffers live in the library
        # component-pipelined solve with copy-engine transposes (opt-in: SOPHT_SLAB_PIPELINE=1; measured slower than
        # the transposes fused into the x kernels on 2 GPUs - the DMA traffic competes with the HBM-bound y / z passes)
        self.pipelined = (self.peer_exchange and peer_arena is not None
                          and os.environ.get("SOPHT_SLAB_PIPELINE", "0") != "0")
        if self.pipelined:
            self._init_pipeline(n_components, device)
            self.path += "-pipelined"

    def _create_handle(self, lib, n_components: int, world: int, rank: int) -> ctypes.c_void_p:
        real_t = self.real_t
        mx = _reflected_axis(self.x_range, self.dx, self.grid_size_x, real_t)
        my = _reflected_axis(self.y_range, self.dx, self.grid_size_y, real_t)
        mz = _reflected_axis(self.z_range, self.dx, self.grid_size_z, real_t)
        origin = real_t(1 / (4 * np.pi * self.dx))  # UnboundedPoissonSolverPYFFTW3D.py:79
        handle = ctypes.c_void_p()
        _lib.check(lib.sopht_poisson_slab_create(
            ctypes.byref(handle), n_components, self.grid_size_z, self.grid_size_y, self.grid_size_x, world, rank,
            float(self.dx), _lib.double_array(mz), _lib.double_array(my), _lib.double_array(mx),
            float(origin), _lib.current_stream()))
        self.path = "pow2-slab"
        return handle

    def _make_buffers(self, n_components: int, device, group) -> None:
        nz, ny = self.grid_size_z, self.grid_size_y
        self.plan = SlabTransposePlan(self.part, n_components, torch.float32, device, group)
        # two halves: the x-major spectrum the y forward pass writes, the kx-tile-major one the z pass writes
        self._work = torch.zeros((2, n_components, nz, 2 * ny, self.plan.nxl, 2), dtype=torch.float32, device=device)
        self._nyq_work = torch.zeros((n_components, nz, 2 * ny, 2), dtype=torch.float32, device=device)

    def _open_peer_exchange(self, lib, world: int, device, group) -> None:
        mine = (ctypes.c_ubyte * 128)()
        _lib.check(lib.sopht_poisson_slab_enable_peer_exchange(self._handle, ctypes.cast(mine, ctypes.c_void_p)))
        local = torch.tensor(list(mine), dtype=torch.uint8, device=device)
        gathered = torch.zeros(world * 128, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(gathered, local, group=group)
        blob = bytes(gathered.cpu().tolist())
        buf = (ctypes.c_ubyte * len(blob)).from_buffer_copy(blob)
        _lib.check(lib.sopht_poisson_slab_open_peers(self._handle, ctypes.cast(buf, ctypes.c_void_p)))
        dist.barrier(group=group)

    # -- pipelined solve ---------------------------------------------------------------------------------------
    def _init_pipeline(self, n_components: int, device) -> None:
        nz, ny = self.grid_size_z, self.grid_size_y
        if self._work is not None:  # one component at a time: a third of the workspace
            self._work = torch.zeros((2, 1, nz, 2 * ny, self.plan.nxl, 2), dtype=torch.float32, device=device)
            self._nyq_work = torch.zeros((1, nz, 2 * ny, 2), dtype=torch.float32, device=device)
        self._copy_streams = [torch.cuda.Stream(device=device) for _ in range(4)]
        # high priority: its one-warp barrier kernels must slip in between the persistent compute kernels
        self._ctrl_stream = torch.cuda.Stream(device=device, priority=-1)
        self._stream_array = (ctypes.c_void_p * len(self._copy_streams))(
            *[ctypes.c_void_p(s.cuda_stream) for s in self._copy_streams])
        self._n_components = n_components

    def _transpose(self, lib, c: int, backward: bool, after: torch.cuda.Event) -> torch.cuda.Event:
        """Enqueue the copy-engine transposes of component c once `after` has fired; returns the event that fires when
        EVERY rank's copies of this component have landed (copies -> all-ranks barrier on the control stream)."""
        plan = self.plan
        for s in self._copy_streams:
            s.wait_event(after)
        nyq = None if backward else ctypes.c_void_p(plan.nyq_local.data_ptr())
        _lib.check(lib.sopht_poisson_slab_pipe_transpose(
            self._handle, c, int(backward), self._stream_array, len(self._copy_streams), nyq))
        for s in self._copy_streams:
            ev = torch.cuda.Event()
            ev.record(s)
            self._ctrl_stream.wait_event(ev)
        with torch.cuda.stream(self._ctrl_stream):
            self._peer_arena.barrier()
        done = torch.cuda.Event()
        done.record(self._ctrl_stream)
        return done

    def _solve_pipelined(self, fr, fs) -> None:
        """Per component: x forward -> DMA transpose -> y, z, y inverse -> DMA transpose back -> x inverse, with the
        transposes of one component running under the compute of the others (three compute phases on the caller's
        stream, copies on four side streams, the all-ranks barriers on a control stream)."""
        lib, plan = _lib.load(), self.plan
        p = ctypes.c_void_p
        main = torch.cuda.current_stream()
        st = _lib.current_stream()
        ncomp = self._n_components
        nyq_local = p(plan.nyq_local.data_ptr())
        work = p(self._work.data_ptr()) if self._work is not None else None
        nyq_work = p(self._nyq_work.data_ptr()) if self._nyq_work is not None else None
        arrived = []
        for c in range(ncomp):
            _lib.check(lib.sopht_poisson_slab_pipe_forward_x(self._handle, ctypes.byref(fr), c, nyq_local, st))
            ev = torch.cuda.Event()
            ev.record(main)
            arrived.append(self._transpose(lib, c, False, ev))
        returned = []
        for c in range(ncomp):
            main.wait_event(arrived[c])
            _lib.check(lib.sopht_poisson_slab_pipe_yz(self._handle, c, work, nyq_work, st))
            ev = torch.cuda.Event()
            ev.record(main)
            returned.append(self._transpose(lib, c, True, ev))
            if c >= 1:
                main.wait_event(returned[c - 1])
                _lib.check(lib.sopht_poisson_slab_pipe_inverse_x(
                    self._handle, ctypes.byref(fs), c - 1, nyq_local, st))
        main.wait_event(returned[ncomp - 1])
        _lib.check(lib.sopht_poisson_slab_pipe_inverse_x(self._handle, ctypes.byref(fs), ncomp - 1, nyq_local, st))

    def __del__(self) -> None:
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                _lib.load().sopht_poisson_slab_destroy(h)
            except Exception:  # noqa: BLE001 - interpreter shutdown
                pass
            self._handle = None

    def vector_field_solve(self, solution_vector_field: torch.Tensor, rhs_vector_field: torch.Tensor) -> None:
        """-del^2(solution) = rhs on the unbounded global domain; arguments are this rank's z-slabs."""
        lib, plan, st = _lib.load(), self.plan, _lib.current_stream()
        p = ctypes.c_void_p
        fr = _lib.field_desc(rhs_vector_field, _lib.SOPHT_F32)
        fs = _lib.field_desc(solution_vector_field, _lib.SOPHT_F32)

        def forward_x() -> None:
            _lib.check(lib.sopht_poisson_slab_forward_x(
                self._handle, ctypes.byref(fr), p(plan.send.data_ptr()), p(plan.nyq_local.data_ptr()), st))

        def middle() -> None:
            _lib.check(lib.sopht_poisson_slab_yz(
                self._handle, p(plan.recv.data_ptr()), p(plan.nyq_all.data_ptr()),
                p(self._work.data_ptr()) if self._work is not None else None,
                p(self._nyq_work.data_ptr()) if self._nyq_work is not None else None, st))

        def inverse_x() -> None:
            _lib.check(lib.sopht_poisson_slab_inverse_x(
                self._handle, ctypes.byref(fs), p(plan.send.data_ptr()), p(plan.nyq_local.data_ptr()), st))

        if not self.peer_exchange:
            plan.solve(forward_x, middle, inverse_x)
            return
        if self.pipelined:
            self._solve_pipelined(fr, fs)
            return
        # peer-memory path: the kernels themselves move the spectrum over NVLink (x forward pushes its chunks into the
        # owners' buffers, x inverse pulls its chunks from the y-inverse outputs); collectives only order the phases
        part, nzl = self.part, plan.nzl
        _lib.check(lib.sopht_poisson_slab_forward_x(
            self._handle, ctypes.byref(fr), None, p(plan.nyq_local.data_ptr()), st))
        with _lib.profile_range("comm.nyquist_allgather_barrier"):
            # all ranks have finished writing into each other's buffers once this all-gather completes
            dist.all_gather_into_tensor(self._nyq_gather, plan.nyq_local, group=plan.group)
            plan.nyq_all.view(plan.ncomp, part.world_size, nzl, -1, 2).copy_(self._nyq_gather.transpose(0, 1))
        _lib.check(lib.sopht_poisson_slab_yz(
            self._handle, None, p(plan.nyq_all.data_ptr()),
            p(self._work.data_ptr()) if self._work is not None else None,
            p(self._nyq_work.data_ptr()) if self._nyq_work is not None else None, st))
        if self._peer_arena is not None:
            self._peer_arena.barrier()
        else:
            with _lib.profile_range("comm.barrier"):
                dist.all_reduce(self._barrier_flag, group=plan.group)
        plan.nyq_local.copy_(plan.nyq_all[:, part.rank * nzl : (part.rank + 1) * nzl])
        _lib.check(lib.sopht_poisson_slab_inverse_x(
            self._handle, ctypes.byref(fs), None, p(plan.nyq_local.data_ptr()), st))


class SlabPeriodicPoissonSolver3D(SlabUnboundedPoissonSolver3D):
    """-del^2(solution) = rhs on the PERIODIC global box (mean mode dropped), z-slab decomposed: the distributed
    counterpart of PeriodicPoissonSolver3D (an extension for BASELINE config 4 - the reference has nothing periodic,
    parity is against the numpy restatement and analytic modes). Same choreography as the unbounded solver; rows carry
    nx / 2 complex bins (+ the Nyquist plane), the y and z passes run in place on the received kx slab."""

    def __init__(self, grid_size_z: int, grid_size_y: int, grid_size_x: int, x_range: float = 1.0,
                 num_threads: int = 1, real_t: type = np.float32, n_components: int = 3, group: Any = None,
                 peer_exchange: bool | None = None, peer_arena: Any = None, symbol: str = "spectral") -> None:
        if symbol not in ("spectral", "three_point"):
            msg = "symbol must be 'spectral' or 'three_point'"
            raise ValueError(msg)
        self.symbol = symbol
        super().__init__(grid_size_z, grid_size_y, grid_size_x, x_range=x_range, num_threads=num_threads,
                         real_t=real_t, n_components=n_components, group=group, peer_exchange=peer_exchange,
                         peer_arena=peer_arena)

    def _create_handle(self, lib, n_components: int, world: int, rank: int) -> ctypes.c_void_p:
        handle = ctypes.c_void_p()
        _lib.check(lib.sopht_poisson_slab_create_periodic(
            ctypes.byref(handle), n_components, self.grid_size_z, self.grid_size_y, self.grid_size_x, world, rank,
            float(self.dx), int(self.symbol == "three_point"), _lib.current_stream()))
        self.path = "periodic-pow2-slab"
        return handle

    def _make_buffers(self, n_components: int, device, group) -> None:
        # the spectrum has nx / 2 bins per row: the transposes move (C, P, nz/P, ny, nx/2/P)
        spec_part = SlabPartition((self.grid_size_z, self.grid_size_y, self.grid_size_x // 2), self.part.world_size,
                                  self.part.rank)
        self.plan = SlabTransposePlan(spec_part, n_components, torch.float32, device, group)
        self._work = self._nyq_work = None  # y / z passes run in place
