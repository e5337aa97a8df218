"""utils/distributed/comm.py of the reference, name for name, on fewer collectives.

Same call signatures and return values (tests/test_compat_gloo.py runs both, rank for rank, under gloo); what changes
is the plumbing: no barrier after every wrapper (comm.py:56-67, 117, 125, 183, 192, 239), one header broadcast instead of
three (scatter_tensor / broadcast_tensor), all_gather_into_tensor instead of lists of per-rank tensors.  Works with the
`nccl` backend on CUDA tensors (B200, NVLink) and with `gloo` on CPU tensors (tests).
"""
from __future__ import annotations

import sys
from typing import List, Union

import torch
import torch.distributed as dist
from torch import Tensor

_MAX_DIMS = 8


def get_world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def get_rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def synchronize() -> None:
    """barrier among all processes (comm.py:25-37)"""
    if get_world_size() > 1:
        dist.barrier()


def destroy_all_rank() -> None:
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
    sys.exit(0)


def all_reduce_tensor(tensors: Union[Tensor, List[Tensor]], op=dist.ReduceOp.SUM, world_size: int = 1,
                      in_place: bool = True) -> Union[List[Tensor], None]:
    """all_reduce then divide by `world_size` (comm.py:40-73, quirk Q9 kept: world size 1 returns [tensors] undivided)."""
    if isinstance(tensors, list):
        tensor_list = tensors
    elif isinstance(tensors, Tensor):
        tensor_list = [tensors]
    else:
        raise TypeError("tensors must be Tensor or List[Tensor]")
    if get_world_size() == 1:
        return [tensors]
    out: List[Tensor] = []
    for tensor in tensor_list:
        if not in_place:
            tensor = tensor.clone()
        if tensor.is_complex():  # gloo reduces real tensors only
            dist.all_reduce(torch.view_as_real(tensor), op)
        else:
            dist.all_reduce(tensor, op)
        tensor.div_(world_size)
        out.append(tensor)
    return None if in_place else out


def _header(tensor, device, master_rank: int):
    """[dim, shape...] of the master's tensor on every rank: ONE broadcast"""
    h = torch.zeros(1 + _MAX_DIMS, device=device, dtype=torch.int64)
    if get_rank() == master_rank:
        assert tensor.dim() <= _MAX_DIMS
        h[0] = tensor.dim()
        if tensor.dim():
            h[1 : 1 + tensor.dim()] = torch.tensor(tuple(tensor.shape), dtype=torch.int64)
    dist.broadcast(h, src=master_rank)
    h = h.tolist()
    return tuple(h[1 : 1 + h[0]])


def scatter_tensor(tensor: Tensor, device: torch.device, dtype: torch.dtype, world_size: int, master_rank: int = 0) -> Tensor:
    """rank r receives rows [r k + min(r, res), ...) of the master's tensor, k, res = divmod(rows, world_size)
    (comm.py:76-158)."""
    if get_world_size() == 1:
        return tensor
    shape = _header(tensor, device, master_rank)
    k, res = divmod(shape[0], world_size)
    sizes = [k + (1 if r < res else 0) for r in range(world_size)]
    rank = get_rank()
    data = torch.zeros((sizes[0],) + shape[1:], dtype=dtype, device=device)
    scatter_data = None
    if rank == master_rank:
        parts = tensor.to(dtype).split(sizes, dim=0)
        scatter_data = [p if p.size(0) == sizes[0] else torch.cat((p, p.new_zeros((sizes[0] - p.size(0),) + shape[1:]))) for p in parts]
    dist.scatter(data, scatter_data, src=master_rank)
    return data[: sizes[rank]]


def broadcast_tensor(tensor: Union[Tensor, None], device: torch.device, dtype: torch.dtype, master_rank: int = 0) -> Tensor:
    """the master's tensor, converted to `dtype`, on every rank (comm.py:161-207); complex supported."""
    if get_world_size() == 1:
        return tensor
    shape = _header(tensor, device, master_rank)
    if get_rank() == master_rank:
        tensor = tensor.to(dtype=dtype).contiguous()
    else:
        assert tensor is None
        tensor = torch.empty(shape, dtype=dtype, device=device)
    if dtype.is_complex:
        dist.broadcast(torch.view_as_real(tensor), src=master_rank)
    else:
        dist.broadcast(tensor, src=master_rank)
    return tensor


def _gather_sizes(tensor: Tensor, device, world_size: int) -> List[int]:
    mine = torch.tensor([tensor.size(0)], device=device, dtype=torch.int64)
    sizes = torch.empty(world_size, device=device, dtype=torch.int64)
    dist.all_gather_into_tensor(sizes, mine)
    return sizes.tolist()


def all_gather_tensor(tensor: Tensor, device: torch.device, world_size: int) -> List[Tensor]:
    """every rank's tensor (first dimensions may differ) on every rank (comm.py:277-328)."""
    if get_world_size() == 1:
        return [tensor]
    if tensor.dim() == 0:  # dist_stats gathers 0-d counts (dist_stats.py:73-76); the reference returns 0-d tensors
        out = torch.empty(world_size, device=device, dtype=tensor.dtype)
        dist.all_gather_into_tensor(out, tensor.reshape(1))
        return [out[r] for r in range(world_size)]
    sizes = _gather_sizes(tensor, device, world_size)
    cplx = tensor.is_complex()
    t = torch.view_as_real(tensor) if cplx else tensor
    n_max = max(sizes)
    if t.size(0) < n_max:
        t = torch.cat((t, t.new_zeros((n_max - t.size(0),) + tuple(t.shape[1:]))))
    out = torch.empty((world_size * n_max,) + tuple(t.shape[1:]), device=device, dtype=t.dtype)
    dist.all_gather_into_tensor(out, t.contiguous())
    parts = [out[r * n_max : r * n_max + sizes[r]] for r in range(world_size)]
    return [torch.view_as_complex(p) for p in parts] if cplx else parts


def gather_tensor(tensor: Tensor, device: torch.device, world_size: int, master_rank: int = 0) -> Union[List[Tensor], None]:
    """every rank's tensor on the master, None elsewhere (comm.py:210-274)."""
    if get_world_size() == 1:
        return [tensor]
    sizes = _gather_sizes(tensor, device, world_size)
    cplx = tensor.is_complex()
    t = torch.view_as_real(tensor) if cplx else tensor
    n_max = max(sizes)
    if t.size(0) < n_max:
        t = torch.cat((t, t.new_zeros((n_max - t.size(0),) + tuple(t.shape[1:]))))
    t = t.contiguous()
    if get_rank() == master_rank:
        padded = [torch.zeros_like(t) for _ in range(world_size)]
        dist.gather(t, gather_list=padded, dst=master_rank)
        parts = [p[:s] for p, s in zip(padded, sizes)]
        return [torch.view_as_complex(p) for p in parts] if cplx else parts
    dist.gather(t, gather_list=None, dst=master_rank)
    return None
