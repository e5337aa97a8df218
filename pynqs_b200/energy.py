"""Sample-space local energy (ElocMethod.SAMPLE_SPACE, vmc/energy/eloc.py:326-397) on the new ops.

`local_energy_reduced` is the REDUCE method (eloc.py:257-297) on compacted rows.

Two sample-space routes with identical results:
  * `local_energy_sample_space`  -- the one-pass kernels (scan of the string-grouped table copies -> H_ij on
    the hits -> reduce); nothing of size [n, M] is materialised, so 10^6 samples are a single call;
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


def Func(func, x: Tensor, WF_LUT: "WavefunctionLUT | None" = None, use_unique: bool = False) -> Tensor:
    """psi of every row of x: table values where the LUT has the row, `func` (the ansatz) on the rest -- on the distinct
    rows only when use_unique.  Mirror of vmc/energy/flip.py:29-63 on the library's lookup + compaction
    (WavefunctionLUT.lookup) and its sort-based unique_onv instead of torch.unique(dim=0)."""
    batch = x.size(0)
    lut_idx = lut_not_idx = lut_value = None
    if WF_LUT is not None:
        lut_idx, lut_not_idx, lut_value = WF_LUT.lookup(x)
        x = x[lut_not_idx]
    if use_unique and x.size(0) > 0:
        unique_x, inverse = ops.unique_onv(x.contiguous())
        psi0 = torch.index_select(func(unique_x), 0, inverse)
    else:
        psi0 = func(x)
    if WF_LUT is None:
        return psi0
    psi = torch.empty(batch, dtype=psi0.dtype, device=psi0.device)
    psi[lut_idx] = lut_value.to(psi0.dtype)
    psi[lut_not_idx] = psi0
    return psi


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


def local_energy_reduced(x: Tensor, h1e: Tensor, h2e: Tensor, psi_of, sorb: int, nele: int, noa: int, nob: int,
                         dtype=torch.double, eps: float = 1.0e-12, batch: int = 65536, eps_sample: int = 0, seed=None,
                         draws=None) -> Tuple[Tensor, Tensor, Tensor]:
    """ElocMethod.REDUCE (eloc.py:204-323).  eps_sample = 0: only the connected determinants with |<x|H|x'>| >= eps enter
    E_loc.  eps_sample > 0 (the setting of the shipped inputs, main.py:149-159): the sub-eps rows are importance-sampled,
    eps_sample draws per sample from p ~ |H|, and enter with (count / eps_sample) * H / p (eloc.py:257-283); `seed` / `draws`
    as in get_comb_hij_sampled (`draws` [n, eps_sample] covers all of x, it is sliced per batch).  `psi_of(x_kept uint8 [K, 8L]) -> Tensor [K]` supplies the
    amplitudes (the reference's Func(ansatz, x, WF_LUT, use_unique)); a WavefunctionLUT may be passed instead,
    absent determinants then count as 0.  Returns (eloc, sloc, psi_x) like _reduce_psi."""
    M = ops.get_Num_SinglesDoubles(sorb, noa, nob) + 1
    if isinstance(psi_of, WavefunctionLUT):
        lut = psi_of

        def psi_of(xk: Tensor) -> Tensor:  # noqa: F811
            found, _, value = lut.lookup(xk)
            out = torch.zeros(xk.size(0), dtype=lut.dtype, device=xk.device)
            out[found] = value
            return out

    elocs, psis = [], []
    for b in range(0, x.size(0), batch):
        if eps_sample > 0:
            xk, hij, idx, offsets = ops.get_comb_hij_sampled(
                x[b : b + batch], h1e, h2e, sorb, nele, noa, nob, eps, eps_sample,
                seed=None if seed is None else int(seed) + b, draws=None if draws is None else draws[b : b + batch])
        else:
            xk, hij, idx, offsets = ops.get_comb_hij_reduced(x[b : b + batch], h1e, h2e, sorb, nele, noa, nob, eps)
        psi = psi_of(xk)
        psi = psi.to(torch.complex128 if psi.is_complex() else torch.float64)
        e, p0 = ops.reduce_eloc(psi, hij, idx, offsets, M)
        elocs.append(e)
        psis.append(p0)
    eloc = torch.cat(elocs) if elocs else torch.empty(0, dtype=dtype, device=x.device)
    psi_x = torch.cat(psis) if psis else torch.empty(0, dtype=dtype, device=x.device)
    return eloc.to(dtype), torch.zeros_like(eloc).to(dtype), psi_x.to(dtype)
