"""Host side of the NVLink peer-memory collectives (csrc/peer_comm.cu): allocate this rank's communication region,
exchange the cudaIpc handles over the existing ``torch.distributed`` group (plumbing only), map the peers, and hand out
zero-copy torch views of the region.  Used by ``CaptionTrainer`` for the gradient exchange of the data-parallel step
(train.py:218 in the reference: DDP's bucketed NCCL all-reduce)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import torch

from . import lib as L

_TYPESTR = {torch.bfloat16: None, torch.float32: "<f4", torch.int64: "<i8", torch.int32: "<i4", torch.uint8: "|u1", torch.int16: "<i2"}


class _Raw:
    """Minimal __cuda_array_interface__ holder so torch can wrap device memory it did not allocate."""

    def __init__(self, ptr: int, shape: Tuple[int, ...], typestr: str, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3, "strides": None}
        self._owner = owner


def peer_comm_supported(device: torch.device, world: int) -> bool:
    if os.environ.get("VCT_COMM", "peer") != "peer" or world < 2 or world > 8:
        return False
    try:
        me = device.index if device.index is not None else torch.cuda.current_device()
        return all(torch.cuda.can_device_access_peer(me, p) for p in range(torch.cuda.device_count()) if p != me)
    except Exception:
        return False


class PeerComm:
    def __init__(self, rank: int, world: int, nbytes: int, device: torch.device, group=None, ctas: Optional[int] = None):
        import torch.distributed as dist
        self.lib = L.load()
        self.rank, self.world, self.device = rank, world, device
        self.nbytes = (int(nbytes) + 255) // 256 * 256
        self.ctas = int(ctas or os.environ.get("VCT_COMM_CTAS", "32"))
        h = C.c_void_p()
        with torch.cuda.device(device):
            L.check(self.lib.vct_comm_create(rank, world, self.nbytes, self.ctas, C.byref(h)), "vct_comm_create")
            self.handle = h
            self.base = int(self.lib.vct_comm_base(h))
            if world > 1:
                mine = (C.c_ubyte * 64)()
                L.check(self.lib.vct_comm_ipc_handle(h, mine), "vct_comm_ipc_handle")
                handles = [None] * world
                dist.all_gather_object(handles, bytes(mine), group=group)
                blob = b"".join(handles)
                L.check(self.lib.vct_comm_connect(h, blob), "vct_comm_connect")
                dist.barrier(group=group)            # every rank has mapped every region before anyone uses them
        self._cursor = 0

    def reserve(self, nbytes: int) -> int:
        """Byte offset of a fresh 256-byte aligned block of the region (the same sequence of calls on every rank gives the
        same layout everywhere, which is what the kernels assume)."""
        off = self._cursor
        self._cursor = (off + int(nbytes) + 255) // 256 * 256
        if self._cursor > self.nbytes:
            raise RuntimeError(f"peer communication region exhausted ({self._cursor} > {self.nbytes} bytes)")
        return off

    def tensor(self, byte_off: int, shape, dtype: torch.dtype) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= int(s)
        if dtype == torch.bfloat16:
            t = torch.as_tensor(_Raw(self.base + byte_off, (n,), "<i2", self), device=self.device).view(torch.bfloat16)
        else:
            t = torch.as_tensor(_Raw(self.base + byte_off, (n,), _TYPESTR[dtype], self), device=self.device)
        return t.view(*shape)

    def status(self) -> int:
        with torch.cuda.device(self.device):
            return int(self.lib.vct_comm_status(self.handle))

    def close(self) -> None:
        if getattr(self, "handle", None) is not None:
            with torch.cuda.device(self.device):
                self.lib.vct_comm_destroy(self.handle)
            self.handle = None
