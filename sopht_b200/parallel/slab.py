"""z-slab partition of a (nz, ny, nx) grid and the nearest-neighbour halo exchange.

Host-side logic only (works on CPU tensors with the gloo backend too, which is how it is tested without
GPUs): rank r owns planes [r nz/P, (r+1) nz/P); local arrays carry `halo` extra planes on each z side.
"""

from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class SlabPartition:
    """Ownership of z planes by rank."""

    grid_size: tuple[int, int, int]  # global (nz, ny, nx)
    world_size: int
    rank: int
    halo: int = 1

    def __post_init__(self) -> None:
        nz = self.grid_size[0]
        if self.world_size < 1 or not 0 <= self.rank < self.world_size:
            msg = "invalid rank / world size"
            raise ValueError(msg)
        if nz % self.world_size:
            msg = f"nz = {nz} is not divisible by the number of ranks ({self.world_size})"
            raise ValueError(msg)
        if self.world_size > 1 and nz // self.world_size < max(self.halo, 1):
            msg = "fewer planes per rank than halo planes"
            raise ValueError(msg)

    @property
    def nz_local(self) -> int:
        return self.grid_size[0] // self.world_size

    @property
    def z_start(self) -> int:
        return self.rank * self.nz_local

    @property
    def is_first(self) -> bool:
        return self.rank == 0

    @property
    def is_last(self) -> bool:
        return self.rank == self.world_size - 1

    @property
    def local_shape(self) -> tuple[int, int, int]:
        """Shape of a local scalar array including halo planes."""
        return (self.nz_local + 2 * self.halo, self.grid_size[1], self.grid_size[2])

    def owned(self, field: torch.Tensor) -> torch.Tensor:
        """View of the owned planes of a local (..., nz_local + 2 halo, ny, nx) array."""
        h = self.halo
        return field[..., h : h + self.nz_local, :, :]

    def stencil_view(self, field: torch.Tensor) -> torch.Tensor:
        """View handed to the stencil kernels: owned planes plus the halo planes that hold real neighbours.

        The kernels treat the first / last plane of the array they are given as the global ghost ring
        (not updated). On an interior z side that plane is a halo plane (its value comes from the
        neighbour and is refreshed by the next exchange); on a global boundary the halo plane is cut off
        so that the ring rule lands on the true boundary plane."""
        h = self.halo
        lo = h if self.is_first else h - 1
        hi = h + self.nz_local if self.is_last else h + self.nz_local + 1
        return field[..., lo:hi, :, :]

    @property
    def z_faces(self) -> int:
        """bit 0 / bit 1: the low / high z face of the owned slab is a global boundary."""
        return (1 if self.is_first else 0) | (2 if self.is_last else 0)


class HaloExchanger:
    """Nearest-neighbour exchange of the z halo planes with persistent pack / receive buffers (no allocation
    per step: temporaries handed to the NCCL stream would otherwise churn the caching allocator)."""

    def __init__(self, part: SlabPartition, group=None) -> None:
        self.part, self.group = part, group
        self._bufs: dict = {}

    def _buffers(self, key, like: torch.Tensor):
        b = self._bufs.get(key)
        if b is None:
            b = tuple(torch.empty_like(like) for _ in range(4))  # send-low, recv-low, send-high, recv-high
            self._bufs[key] = b
        return b

    def __call__(self, fields) -> None:
        part = self.part
        if part.world_size == 1:
            return
        h, n = part.halo, part.nz_local
        ops, unpack = [], []
        for slot, f in enumerate(fields):
            lo_src, hi_src = f[..., h : 2 * h, :, :], f[..., n : n + h, :, :]
            sl, rl, sh, rh = self._buffers((slot, tuple(f.shape), f.dtype, f.device), lo_src)
            if not part.is_first:  # low neighbour: send my first owned planes, receive its last owned planes
                sl.copy_(lo_src)
                ops.append(dist.P2POp(dist.isend, sl, part.rank - 1, group=self.group))
                ops.append(dist.P2POp(dist.irecv, rl, part.rank - 1, group=self.group))
                unpack.append((f, slice(0, h), rl))
            if not part.is_last:
                sh.copy_(hi_src)
                ops.append(dist.P2POp(dist.isend, sh, part.rank + 1, group=self.group))
                ops.append(dist.P2POp(dist.irecv, rh, part.rank + 1, group=self.group))
                unpack.append((f, slice(n + h, n + 2 * h), rh))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        for f, sl_, recv in unpack:
            f[..., sl_, :, :].copy_(recv)


def exchange_halos(part: SlabPartition, fields, group=None) -> None:
    """Fill the halo planes of every local array in `fields` from the z neighbours (one batched
    send/recv group; the outermost halos on the global boundary are left untouched). One-off form of
    HaloExchanger (allocates its buffers per call)."""
    HaloExchanger(part, group)(fields)
