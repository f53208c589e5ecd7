"""Peer-memory arena of the z-slab decomposition: field storage that the other ranks of the box address over
NVLink (CUDA IPC), the halo exchange as one kernel of direct stores into the neighbours' halo planes, and a
device-side barrier (libsopht_b200: sopht_peer_*, csrc/peer.cu). Every rank must create the arena with the same
size and allocate the same arrays in the same order, so that an offset names the same array on every rank."""

from __future__ import annotations

import ctypes
from typing import Any

import numpy as np
import torch
import torch.distributed as dist

from sopht_b200 import _lib


class _DeviceSpan:
    """Raw device memory presented to torch through __cuda_array_interface__ (zero copy)."""

    def __init__(self, owner: Any, ptr: int, count: int, typestr: str) -> None:
        self._owner = owner  # keeps the arena alive as long as a tensor views it
        self.__cuda_array_interface__ = {
            "shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 2, "strides": None}


class PeerArena:
    def __init__(self, payload_bytes: int, group=None, periodic_z: bool = False) -> None:
        if not torch.cuda.is_available():
            msg = "the peer arena needs CUDA devices (no CPU fallback)"
            raise _lib.SophtLibraryError(msg)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = torch.device("cuda", torch.cuda.current_device())
        lib = _lib.load()
        handle = ctypes.c_void_p()
        mine = (ctypes.c_ubyte * 64)()
        _lib.check(lib.sopht_peer_arena_create(
            ctypes.byref(handle), int(payload_bytes), self.world, self.rank, ctypes.cast(mine, ctypes.c_void_p)))
        self._handle = handle
        if self.world > 1:
            local = torch.tensor(list(mine), dtype=torch.uint8, device=self.device)
            gathered = torch.zeros(self.world * 64, dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(gathered, local, group=group)
            blob = bytes(gathered.cpu().tolist())
            buf = (ctypes.c_ubyte * len(blob)).from_buffer_copy(blob)
            _lib.check(lib.sopht_peer_arena_open(handle, ctypes.cast(buf, ctypes.c_void_p)))
            dist.barrier(group=group)
        if periodic_z:  # the ranks form a ring: halo exchange of a periodic box
            _lib.check(lib.sopht_peer_arena_set_periodic(handle, 1))
        self._base = int(lib.sopht_peer_arena_payload(handle))
        self._size = int(payload_bytes)
        self._used = 0

    def alloc(self, shape: tuple[int, ...], real_t: type = np.float32) -> torch.Tensor:
        """Zero-initialised array of `shape` inside the arena (256-byte aligned)."""
        dt = np.dtype(real_t)
        count = int(np.prod(shape))
        start = (self._used + 255) // 256 * 256
        if start + count * dt.itemsize > self._size:
            msg = "peer arena exhausted"
            raise MemoryError(msg)
        self._used = start + count * dt.itemsize
        span = _DeviceSpan(self, self._base + start, count, dt.str)
        return torch.as_tensor(span, device=self.device).view(*shape)

    def halo_exchange(self, fields, nz_local: int, halo: int) -> None:
        """Fill the z halo planes of local arrays (C, nz_local + 2 halo, ny, nx) allocated from this arena."""
        n = len(fields)
        offs = (ctypes.c_int64 * n)(*[f.data_ptr() - self._base for f in fields])
        cstr = (ctypes.c_int64 * n)(*[f.stride(0) * f.element_size() for f in fields])
        ncomp = (ctypes.c_int * n)(*[f.shape[0] for f in fields])
        f0 = fields[0]
        plane = f0.shape[-1] * f0.shape[-2] * f0.element_size()
        for f in fields:
            if not f[0].is_contiguous() or f.shape[1] != nz_local + 2 * halo:
                msg = "halo_exchange expects (C, nz_local + 2 halo, ny, nx) arrays with contiguous components"
                raise ValueError(msg)
        _lib.check(_lib.load().sopht_peer_halo_exchange(
            self._handle, n, offs, cstr, ncomp, nz_local, halo, plane, _lib.current_stream()))

    def barrier(self) -> None:
        _lib.check(_lib.load().sopht_peer_barrier(self._handle, _lib.current_stream()))

    def check(self) -> None:
        """Raise if a device-side poll of this arena gave up waiting for a peer (a rank died, or the ranks issued
        different sequences of exchanges / barriers). Synchronises the device; call it where the host waits anyway."""
        stalled = ctypes.c_int(-1)
        _lib.check(_lib.load().sopht_peer_arena_status(self._handle, ctypes.byref(stalled)))

    def __del__(self) -> None:
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                _lib.load().sopht_peer_arena_destroy(h)
            except Exception:  # noqa: BLE001 - interpreter shutdown
                pass
            self._handle = None
