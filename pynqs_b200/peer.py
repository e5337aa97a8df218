"""Symmetric-memory buffers for the pull collectives of csrc/peer.cu (one process per GPU, NVLink / NVSwitch).

`PeerExchange.get(nbytes)` returns the process-wide exchange area -- a torch.distributed._symmetric_memory allocation of at
least `nbytes` mapped into every peer -- or None when peer memory cannot be set up (no NVLink peer access, a sandbox that
forbids the handle exchange, a CPU / gloo run); callers then use torch.distributed collectives (NCCL), with identical results.
Whether it works is decided once, collectively, so that all ranks take the same route.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist
from torch import Tensor

from . import _lib
from ._lib import i64, vp

_state = {"enabled": True, "area": None, "failed": False}


def set_enabled(on: bool) -> None:
    """Switch the peer-memory route off (NCCL collectives only) or back on."""
    _state["enabled"] = bool(on)


class PeerExchange:
    def __init__(self, nbytes: int, device: torch.device):
        import torch.distributed._symmetric_memory as symm_mem

        self.nbytes = int(nbytes)
        self.device = device
        self.buf = symm_mem.empty(self.nbytes, dtype=torch.uint8, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, dist.group.WORLD)
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        self._ptrs = int(self.hdl.buffer_ptrs_dev)  # device array of the peers' buffer addresses
        self._phase = 0

    # -- collective life cycle: publish (write into self.buf), barrier, pull, barrier ------------------------------------
    def barrier(self) -> None:
        self.hdl.barrier(channel=self._phase & 1)
        self._phase += 1

    def local(self, offset: int, nbytes: int) -> Tensor:
        return self.buf[offset : offset + nbytes]

    def gather(self, src_offset: int, bytes_per_rank: int, out: Tensor) -> None:
        """out (contiguous, world * bytes_per_rank bytes) <- every peer's bytes [src_offset, src_offset + bytes_per_rank)"""
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().pynqs_peer_gather(vp(self._ptrs), self.world, i64(src_offset), i64(bytes_per_rank), vp(out.data_ptr()),
                                                     vp(torch.cuda.current_stream(self.device).cuda_stream)))

    def gather_rows(self, src_offset: int, pos: Tensor, total: int, out: Tensor) -> None:
        """out[i] <- element pos[i] of the concatenation of the peers' pieces (split_length_idx sizes) at src_offset"""
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().pynqs_peer_gather_rows(vp(self._ptrs), self.world, i64(src_offset), vp(pos.data_ptr()), i64(pos.numel()),
                                                          i64(total), int(out.element_size()), vp(out.data_ptr()),
                                                          vp(torch.cuda.current_stream(self.device).cuda_stream)))

    # ---------------------------------------------------------------------------------------------------------------------
    @staticmethod
    def get(nbytes: int, device: torch.device) -> "Optional[PeerExchange]":
        """The exchange area (grown collectively when a larger one is asked for), or None.  Collective: every rank must call
        it with the same nbytes at the same point."""
        if not _state["enabled"] or _state["failed"] or device.type != "cuda":
            return None
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and dist.get_backend() == "nccl"):
            return None
        area = _state["area"]
        if area is not None and area.nbytes >= nbytes:
            return area
        ok = 1
        try:
            size = max(int(nbytes), 1 << 20)
            size = 1 << (size - 1).bit_length()  # powers of two: few re-allocations
            area = PeerExchange(size, device)
        except Exception:  # noqa: BLE001  (anything: unsupported driver, forbidden handle exchange, ...)
            area, ok = None, 0
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            _state["failed"] = True
            _state["area"] = None
            return None
        _state["area"] = area
        return area


def route() -> str:
    return "peer-memory pull kernels (NVLink)" if _state["area"] is not None else "torch.distributed collectives"


def all_gather_columns(cols: List[Tensor], n_rows: int) -> Optional[List[Tensor]]:
    """All-gather of equally sized column tensors ([n_rows, ...] each) through the exchange area; None if unavailable."""
    world = dist.get_world_size()
    dev = cols[0].device
    sizes = [((t.numel() * t.element_size() + 15) // 16) * 16 for t in cols]
    area = PeerExchange.get(sum(sizes), dev)
    if area is None:
        return None
    off = 0
    offs = []
    for t, sz in zip(cols, sizes):
        area.local(off, t.numel() * t.element_size()).copy_(t.reshape(-1).view(torch.uint8))
        offs.append(off)
        off += sz
    area.barrier()  # every rank's columns are in place
    outs = []
    for t, sz, o in zip(cols, sizes, offs):
        if sz == t.numel() * t.element_size():
            out = torch.empty((world * n_rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
            area.gather(o, sz, out)
        else:  # padded pieces: gather padded, then drop the padding
            tmp = torch.empty(world * sz, dtype=torch.uint8, device=dev)
            area.gather(o, sz, tmp)
            out = tmp.view(world, sz)[:, : t.numel() * t.element_size()].contiguous().view(t.dtype).reshape((world * n_rows,) + tuple(t.shape[1:]))
        outs.append(out)
    area.barrier()  # every rank has pulled: the area may be overwritten
    return outs
