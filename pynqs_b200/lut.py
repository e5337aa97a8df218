"""WavefunctionLUT mirror (utils/public_function.py:749-868) and the key sort it relies on
(torch_sort_onv / torch_lexsort, utils/public_function.py:626-689).

The reference sorts with 8L stable argsorts, one per byte column.  A stable sort by the full
little-endian multi-word integer gives the identical permutation, so here it is L stable sorts of
64-bit words, least-significant word first: on CUDA tensors the library's sort_table (radix passes
over the significant bits only + one gather of keys, values and permutation, csrc/table.cu); on CPU
tensors (host-side tests, gloo) torch.sort.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import Tensor

_SIGN = -(1 << 63)


def sort_onv(bra: Tensor) -> Tensor:
    """Permutation that sorts ONV rows ascending as little-endian multi-word integers, ties in
    input order -- same result as the reference's torch_sort_onv(bra) (public_function.py:651-689)."""
    assert bra.dim() == 2 and bra.dtype == torch.uint8 and bra.size(1) % 8 == 0
    n, w = bra.shape
    L = w // 8
    if n == 0:
        return torch.empty(0, dtype=torch.int64, device=bra.device)
    if bra.is_cuda:
        from .C_extension import sort_table

        return sort_table(bra.contiguous(), None, 0, want_perm=True)[2]
    words = bra.contiguous().view(torch.int64).view(n, L)
    idx: Optional[Tensor] = None
    for k in range(L):  # least-significant word first (LSD)
        col = words[:, k] ^ _SIGN  # unsigned order -> signed order
        if idx is None:
            idx = torch.sort(col, stable=True).indices
        else:
            idx = idx[torch.sort(col[idx], stable=True).indices]
    return idx


def split_length_idx(dim: int, length: int) -> List[int]:
    """cumulative slice ends of `dim` items over `length` ranks (public_function.py:720-746)."""
    q, r = divmod(dim, length)
    out, acc = [], 0
    for i in range(length):
        acc += q + (1 if i < r else 0)
        out.append(acc)
    return out


class WavefunctionLUT:
    """Sorted key table + values with lookup; same constructor, attributes and return values as
    the reference class.  On CUDA the table is sorted by the library (sort_table) and two device accelerators
    are built on first use: a hash index behind lookup(), string-grouped copies behind the one-pass local energy."""

    def __init__(self, bra_key: Tensor, wf_value: Tensor, sorb: int, device=None, sort: bool = True,
                 rank: Optional[int] = None, world_size: Optional[int] = None) -> None:
        assert bra_key.dim() == 2 and bra_key.dtype == torch.uint8
        assert bra_key.size(0) == wf_value.size(0)
        self.sort = sort
        if device is not None:
            bra_key = bra_key.to(device)
            wf_value = wf_value.to(device)
        if sort:
            if bra_key.is_cuda and wf_value.dim() == 1 and wf_value.element_size() in (8, 16):
                from .C_extension import sort_table

                # ONVs carry no bits at or above sorb, so only ceil(sorb / 8) radix digits are sorted
                self._bra_key, self._wf_value, idx = sort_table(bra_key.contiguous(), wf_value.contiguous(), sorb)
            else:
                idx = sort_onv(bra_key)
                self._bra_key = bra_key[idx].contiguous()
                self._wf_value = wf_value[idx].contiguous()
            self._sort_perm = idx
            self._idx_sorted = None  # inverse permutation, built on first use (index_value only)
        else:
            self._bra_key = bra_key.contiguous()
            self._wf_value = wf_value.contiguous()
        self.sorb = sorb
        if rank is None or world_size is None:
            import torch.distributed as dist

            on = dist.is_available() and dist.is_initialized()
            rank = dist.get_rank() if on else 0
            world_size = dist.get_world_size() if on else 1
        self.rank, self.world_size = rank, world_size
        self.rank_idx = [0] + split_length_idx(bra_key.size(0), world_size)
        self.rank_begin = self.rank_idx[rank]
        self.rank_end = self.rank_idx[rank + 1]
        # device accelerators, built on first use: the hash index behind lookup(), the string-grouped
        # copies behind the one-pass local energy
        self._hash = None
        self._group = None

    @property
    def idx_sorted(self) -> Tensor:
        """position of every input row in the sorted table (reference attribute, public_function.py:776)"""
        if self._idx_sorted is None:
            inv = torch.empty_like(self._sort_perm)
            inv[self._sort_perm] = torch.arange(self._sort_perm.numel(), device=self._sort_perm.device)
            self._idx_sorted = inv
        return self._idx_sorted

    @property
    def bra_key(self) -> Tensor:
        return self._bra_key

    @property
    def wf_value(self) -> Tensor:
        return self._wf_value

    @property
    def dtype(self):
        return self._wf_value.dtype

    @property
    def hash_index(self):
        if self._hash is None and self._bra_key.is_cuda:
            from .C_extension import HashIndex

            self._hash = HashIndex(self._bra_key)
        return self._hash

    @property
    def group_index(self):
        if self._group is None and self._bra_key.is_cuda:
            from .C_extension import GroupIndex

            self._group = GroupIndex(self._bra_key)
        return self._group

    @property
    def memory(self) -> float:
        extra = sum(ix.nbytes for ix in (self._hash, self._group) if ix is not None)
        return (self.bra_key.numel() + extra) / 2**20

    def lookup(self, onv: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        """(positions of onv rows found in the table, positions not found, table values of the found rows)
        -- public_function.py:817-838."""
        from .C_extension import hash_worthwhile, wavefunction_lut

        # the hash index (~200 bytes per key) is only built for query batches that pay for it; small ones take the
        # classic binary search like the reference
        use_hash = self._bra_key.is_cuda and (self._hash is not None or hash_worthwhile(onv.size(0), self._bra_key.size(0)))
        idx_array, mask = wavefunction_lut(self._bra_key, onv, self.sorb, hash_index=self.hash_index if use_hash else False)
        if self._wf_value.dim() == 1 and self._wf_value.element_size() in (8, 16):
            from .C_extension import lookup_compact

            return lookup_compact(idx_array, mask, self._wf_value)  # the reference's four index operations in two passes
        baseline = torch.arange(onv.size(0), device=onv.device, dtype=torch.int64)
        onv_idx = baseline[mask]
        onv_not_idx = baseline[torch.logical_not(mask)]
        value = self._wf_value[idx_array.masked_select(mask)]
        return onv_idx, onv_not_idx, value

    def index_value(self, begin: int, end: int) -> Tensor:
        assert self.sort, "not-sorted does not support index-value"
        begin = self.rank_begin + begin
        end = self.rank_begin + end
        assert self.rank_end >= end, "Index date must be in the same rank"
        return self.wf_value[self.idx_sorted[begin:end]]

    def clean_memory(self) -> None:
        del self._bra_key, self._wf_value
        self._hash = self._group = None

    def __repr__(self) -> str:
        return (
            f"{type(self).__name__}(\n"
            + f"    bra-key shape: {tuple(self.bra_key.size())}\n"
            + f"    wf-value shape: {self.wf_value.size(0)}\n"
            + f"    sorb: {self.sorb}\n"
            + f"    device hash index: {self._hash is not None}\n"
            + f"    Memory: {self.memory:.3f} MiB\n)"
        )
