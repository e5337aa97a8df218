"""Replacement body for Sampler.gather_scatter_sample (vmc/sample.py:627-772): same arguments, same four return values
(unique_rank uint8 ONVs, placeholder, prob_rank * world_size, WF_LUT), same merge order (torch.unique(dim=0) order, psi of
the first occurrence, counts summed; plain concatenation when use_same_tree), same side effect (self.all_sample_counts).

    from pynqs_b200.compat.sampler import gather_scatter_sample
    Sampler.gather_scatter_sample = gather_scatter_sample          # the one-line patch of INTEGRATION.md section 6

Plumbing: one all-gather of the per-rank sizes + ONE packed all-gather of (ONV, psi, counts) and an identical merge on every
rank, instead of three gathers to rank 0, a merge there, two scatters, two broadcasts and ~10 barriers.  CUDA tensors go
through the library (tensor_to_onv, sort, lookup index); CPU tensors (gloo tests) through torch / the caller's C_extension.
"""
from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor

from ..distributed import exchange_unique_samples, rank_slice


def gather_scatter_sample(self, unique: Tensor, counts: Tensor, wf_value: Tensor) -> Tuple[Tensor, Tensor, Tensor, object]:
    if unique.is_cuda:
        from ..C_extension import tensor_to_onv
        from ..lut import WavefunctionLUT
    else:  # host tensors: whatever libs.C_extension the caller runs on (the reference's CPU build in the gloo tests)
        from libs.C_extension import tensor_to_onv
        from utils.public_function import WavefunctionLUT
    onv = tensor_to_onv(unique.byte(), self.sorb)
    use_lut = bool(getattr(self, "use_LUT", True))
    psi = wf_value if use_lut else torch.zeros(onv.size(0), dtype=torch.float64, device=onv.device)
    merged, wf, cnt = exchange_unique_samples(onv, psi, counts, disjoint=bool(self.use_same_tree))
    self.all_sample_counts = cnt if self.rank == 0 else None
    b, e = rank_slice(merged.size(0), self.rank, self.world_size)
    real = self.dtype.to_real() if hasattr(self.dtype, "to_real") else torch.float64
    prob = (cnt / cnt.sum()).to(real)
    lut = WavefunctionLUT(merged, wf.to(self.dtype), self.sorb, self.device) if use_lut else None
    placeholders = torch.ones([], device=self.device, dtype=torch.int64)
    return merged[b:e].contiguous(), placeholders, prob[b:e] * self.world_size, lut
