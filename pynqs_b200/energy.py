"""Sample-space local energy (ElocMethod.SAMPLE_SPACE, vmc/energy/eloc.py:326-397) on the new ops.

Two routes with identical results:
  * `local_energy_sample_space`  -- one fused kernel (enumerate -> hash probe -> H_ij on hits ->
    reduce); nothing of size [n, M] is materialised, so 10^6 samples are a single launch;
  * `local_energy_three_call`    -- the reference's sequence get_comb_hij_fused ->
    WavefunctionLUT.lookup -> scatter/divide/multiply/sum, chunked like vmc/energy/etot.py:76-146,
    kept so the unchanged reference Python keeps working on the new operators.
"""
from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor

from . import C_extension as ops
from .lut import WavefunctionLUT


def local_energy_sample_space(x: Tensor, h1e: Tensor, h2e: Tensor, WF_LUT: WavefunctionLUT, sorb: int, nele: int,
                              noa: int, nob: int, dtype=torch.double) -> Tuple[Tensor, Tensor, Tensor]:
    """Returns (eloc, sloc, psi_x) like _only_sample_space (eloc.py:508); sloc is zero (no spin-raising)."""
    eloc, psi0 = ops.eloc_sample_space(x, h1e, h2e, sorb, nele, noa, nob, WF_LUT.bra_key, WF_LUT.wf_value, WF_LUT.group_index)
    return eloc.to(dtype), torch.zeros_like(eloc).to(dtype), psi0.to(dtype)


def local_energy_three_call(x: Tensor, h1e: Tensor, h2e: Tensor, WF_LUT: WavefunctionLUT, sorb: int, nele: int,
                            noa: int, nob: int, dtype=torch.double, batch: int = 4096) -> Tuple[Tensor, Tensor, Tensor]:
    """The reference's op sequence (eloc.py:369-397) in chunks of `batch` samples."""
    n = x.size(0)
    M = ops.get_Num_SinglesDoubles(sorb, noa, nob) + 1
    eloc = torch.empty(n, dtype=WF_LUT.dtype, device=x.device)
    psi_x = torch.empty(n, dtype=WF_LUT.dtype, device=x.device)
    for b in range(0, n, batch):
        xb = x[b : b + batch]
        comb_x, comb_hij = ops.get_comb_hij_fused(xb, h1e, h2e, sorb, nele, noa, nob)
        x1 = comb_x.reshape(-1, comb_x.size(2))
        psi_x1 = torch.zeros(xb.size(0), M, device=x.device, dtype=WF_LUT.dtype)
        idx, _, value = WF_LUT.lookup(x1)
        psi_x1.view(-1)[idx] = value
        real = torch.double if not WF_LUT.dtype.is_complex else torch.double
        comb_hij = comb_hij.to(real)
        eloc[b : b + batch] = ((psi_x1.T / psi_x1[..., 0]).T * comb_hij).sum(-1)
        psi_x[b : b + batch] = psi_x1[..., 0]
    return eloc.to(dtype), torch.zeros_like(eloc).to(dtype), psi_x.to(dtype)
