"""Multi-GPU plumbing of the local-energy path: one process per GPU, torch.distributed (NCCL over
NVLink on B200, gloo on CPU for tests).

Replaces the reference's rank-0-centric exchange (vmc/sample.py:627-772: three padded gathers to
rank 0 -> merge -> two scatters -> two broadcasts with shape handshakes, every wrapper followed by
a barrier, utils/distributed/comm.py:56-67) by
  1. one all_gather of the per-rank unique counts,
  2. one padded all_gather of a packed record {ONV 8L B | psi 8/16 B | count 8 B},
  3. an identical deterministic merge on every rank (so no broadcast), local slice by
     split_length_idx (utils/public_function.py:720-746),
and the three collectives of utils/stats/dist_stats.py:18-79 by a single all_reduce of a 5-vector.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from .lut import WavefunctionLUT, split_length_idx


def _world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def exchange_unique_samples(onv: Tensor, psi: Tensor, counts: Optional[Tensor] = None, disjoint: bool = False,
                            equal_sizes: bool = False) -> Tuple[Tensor, Tensor, Tensor]:
    """All ranks contribute their locally unique ONVs (uint8 [n_r, 8L]), psi values and sample
    counts; every rank returns the same merged (unique_onv, psi, counts).

    Merge order = the reference's (sample.py:672-698): if `disjoint` (use_same_tree) plain
    concatenation in rank order, else torch.unique(dim=0) order (row-lexicographic, byte 0 most
    significant) with psi taken from the first occurrence and counts summed.
    `equal_sizes`: the caller guarantees every rank contributes the same number of rows, which saves
    the all-gather of the counts and its host synchronisation."""
    rank, world = _world()
    dev = onv.device
    n_r, w = onv.shape
    if counts is None:
        counts = torch.ones(n_r, dtype=torch.int64, device=dev)
    cplx = psi.dtype.is_complex
    pw = 16 if cplx else 8
    psi_bytes = torch.view_as_real(psi.to(torch.complex128)).contiguous().view(torch.uint8) if cplx else psi.to(torch.float64).contiguous().view(torch.uint8)
    rec_w = w + pw + 8
    if world == 1:
        all_onv, all_psi, all_cnt = onv, psi, counts
    else:
        if equal_sizes:
            n_list = [n_r] * world
        else:
            n_all = torch.empty(world, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(n_all, torch.tensor([n_r], dtype=torch.int64, device=dev))
            n_list = n_all.tolist()
        n_max = max(n_list)
        rec = torch.zeros((n_max, rec_w), dtype=torch.uint8, device=dev)
        rec[:n_r, :w] = onv
        rec[:n_r, w : w + pw] = psi_bytes.view(n_r, pw)
        rec[:n_r, w + pw :] = counts.to(torch.int64).contiguous().view(torch.uint8).view(n_r, 8)
        gathered = torch.empty((world * n_max, rec_w), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(gathered, rec)
        gathered = gathered.view(world, n_max, rec_w)
        if min(n_list) == n_max:  # equal pieces: the gathered buffer already is the concatenation
            cat = gathered.view(world * n_max, rec_w)
        else:
            cat = torch.cat([gathered[r, : n_list[r]] for r in range(world)])
        all_onv = cat[:, :w].contiguous()
        pb = cat[:, w : w + pw].contiguous()
        all_psi = torch.view_as_complex(pb.view(torch.float64).view(-1, 2)) if cplx else pb.view(torch.float64).view(-1)
        all_psi = all_psi.to(psi.dtype)
        all_cnt = cat[:, w + pw :].contiguous().view(torch.int64).view(-1)
    if disjoint:
        return all_onv, all_psi, all_cnt
    uniq, inv = torch.unique(all_onv, dim=0, return_inverse=True)
    m = uniq.size(0)
    first = torch.full((m,), all_onv.size(0), dtype=torch.int64, device=dev)
    first.scatter_reduce_(0, inv, torch.arange(all_onv.size(0), device=dev), reduce="amin")
    merged_cnt = torch.zeros(m, dtype=torch.int64, device=dev).index_add_(0, inv, all_cnt)
    return uniq, all_psi[first], merged_cnt


def rank_slice(n_total: int, rank: Optional[int] = None, world: Optional[int] = None) -> Tuple[int, int]:
    r, w = _world()
    rank = r if rank is None else rank
    world = w if world is None else world
    ends = [0] + split_length_idx(n_total, world)
    return ends[rank], ends[rank + 1]


def build_shared_lut(onv: Tensor, psi: Tensor, sorb: int, counts: Optional[Tensor] = None, disjoint: bool = False):
    """exchange + identical LUT on every rank + this rank's slice of (unique, prob).
    prob follows the reference convention prob_rank * world_size (sample.py:772)."""
    rank, world = _world()
    uniq, wf, cnt = exchange_unique_samples(onv, psi, counts, disjoint)
    lut = WavefunctionLUT(uniq, wf, sorb, uniq.device, rank=rank, world_size=world)
    b, e = rank_slice(uniq.size(0), rank, world)
    prob = cnt.to(torch.float64) / cnt.sum()
    return uniq[b:e].contiguous(), prob[b:e] * world, lut


def energy_statistics(eloc: Tensor, prob: Tensor, counts: Optional[int] = None) -> dict:
    """mean / var / sd / se of the local energy -- the quantities of utils/stats/dist_stats.py:18-79
    (mean = sum_ranks sum_i p_i E_i / W and var = sum_ranks sum_i p_i |mean - E_i|^2 / W with
    p = prob * W, sample.py:772 + comm.py:65-67) from ONE collective: an all_gather of the per-rank
    [sum p, local mean, centred second moment, n], combined with the exact identity
    sum p|E - m|^2 = sum p|E - mu|^2 + (sum p)|mu - m|^2, so there is no cancellation."""
    rank, world = _world()
    cplx = eloc.is_complex()
    e = eloc.to(torch.complex128) if cplx else eloc.to(torch.float64)
    p = prob.to(torch.float64)
    if e.numel():
        # one pass: moments of d = E - c about the shift c = E[0] (stable: |d| is of the size of the
        # spread), then M2 about the local mean mu = c + sum(p d) / w by the same exact identity
        c = e[0]
        d = e - c
        if cplx:
            rows = torch.stack([torch.ones_like(p), d.real, d.imag, d.real * d.real + d.imag * d.imag])
        else:
            rows = torch.stack([torch.ones_like(p), d, torch.zeros_like(p), d * d])
        mom = rows @ p  # [w, sum p d_re, sum p d_im, sum p |d|^2]
        w = mom[0]
        dm_re, dm_im = mom[1] / w, mom[2] / w
        m2 = mom[3] - w * (dm_re * dm_re + dm_im * dm_im)
        mu_re = (c.real if cplx else c) + dm_re
        mu_im = (c.imag + dm_im) if cplx else dm_im
    else:
        w = m2 = mu_re = mu_im = torch.zeros((), dtype=torch.float64, device=e.device)
    vec = torch.stack([w, mu_re, mu_im, m2, torch.tensor(float(e.numel()), dtype=torch.float64, device=e.device)])
    if world > 1:
        allv = torch.empty(world * 5, dtype=torch.float64, device=vec.device)
        dist.all_gather_into_tensor(allv, vec)
        allv = allv.view(world, 5)
    else:
        allv = vec.view(1, 5)
    allv = allv.cpu()
    wr, mr, mi, m2r, nr = allv[:, 0], allv[:, 1], allv[:, 2], allv[:, 3], allv[:, 4]
    mean_re = float((wr * mr).sum() / world)
    mean_im = float((wr * mi).sum() / world)
    var = float((m2r + wr * ((mr - mean_re) ** 2 + (mi - mean_im) ** 2)).sum() / world)
    n = int(nr.sum()) if counts is None else counts
    sd = var ** 0.5
    mean = complex(mean_re, mean_im) if cplx else mean_re
    return {"mean": mean, "var": var, "sd": sd, "se": sd / n ** 0.5, "n": n}
