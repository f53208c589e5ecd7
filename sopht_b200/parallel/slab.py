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


def exchange_halos(part: SlabPartition, fields, group=None) -> None:
    """Fill the halo planes of every local array in `fields` from the z neighbours (one batched
    send/recv group; the outermost halos on the global boundary are left untouched)."""
    if part.world_size == 1:
        return
    h = part.halo
    n = part.nz_local
    ops = []
    keep = []
    for f in fields:
        if not part.is_first:  # low neighbour: send my first owned planes, receive its last owned planes
            send = f[..., h : 2 * h, :, :].contiguous()
            recv = torch.empty_like(send)
            ops.append(dist.P2POp(dist.isend, send, part.rank - 1, group=group))
            ops.append(dist.P2POp(dist.irecv, recv, part.rank - 1, group=group))
            keep.append((f, slice(0, h), recv))
        if not part.is_last:
            send = f[..., n : n + h, :, :].contiguous()
            recv = torch.empty_like(send)
            ops.append(dist.P2POp(dist.isend, send, part.rank + 1, group=group))
            ops.append(dist.P2POp(dist.irecv, recv, part.rank + 1, group=group))
            keep.append((f, slice(n + h, n + 2 * h), recv))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    for f, sl, recv in keep:
        f[..., sl, :, :] = recv
