"""Multi-GPU plumbing of the local-energy path: one process per GPU, torch.distributed (NCCL over
NVLink on B200, gloo on CPU for tests).

Replaces the reference's rank-0-centric exchange (vmc/sample.py:627-772: three padded gathers to
rank 0 -> merge -> two scatters -> two broadcasts with shape handshakes, every wrapper followed by
a barrier, utils/distributed/comm.py:56-67) by
  1. one all_gather of the per-rank unique counts (skipped when the caller knows they are equal),
  2. one all_gather per column (ONVs, psi, and the sample counts when there are any) straight into
     the merged tensors -- no pack / unpack pass,
  3. an identical deterministic merge on every rank (so no broadcast), local slice by
     split_length_idx (utils/public_function.py:720-746),
and the three collectives of utils/stats/dist_stats.py:18-79 by a single all_gather of seven doubles per rank
(shifted moments; combined exactly, and lazily -- the host read can be left outside the step).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from . import peer
from .lut import WavefunctionLUT, split_length_idx


def _world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def exchange_unique_samples(onv: Tensor, psi: Tensor, counts: Optional[Tensor] = None, disjoint: bool = False,
                            equal_sizes: bool = False, sizes: Optional[list] = None) -> Tuple[Tensor, Tensor, Tensor]:
    """All ranks contribute their locally unique ONVs (uint8 [n_r, 8L]), psi values and sample
    counts; every rank returns the same merged (unique_onv, psi, counts).

    Merge order = the reference's (sample.py:672-698): if `disjoint` (use_same_tree) plain
    concatenation in rank order, else torch.unique(dim=0) order (row-lexicographic, byte 0 most
    significant) with psi taken from the first occurrence and counts summed.
    `equal_sizes`: the caller guarantees every rank contributes the same number of rows, which saves
    the all-gather of the counts and its host synchronisation; `sizes`: the caller knows every rank's row count (a list
    of world_size ints) -- same saving for ragged pieces."""
    rank, world = _world()
    dev = onv.device
    n_r, w = onv.shape
    if world == 1:
        all_onv, all_psi = onv, psi
        all_cnt = counts if counts is not None else torch.ones(n_r, dtype=torch.int64, device=dev)
    else:
        if sizes is not None:
            n_list = [int(v) for v in sizes]
            assert len(n_list) == world and n_list[rank] == n_r
        elif equal_sizes:
            n_list = [n_r] * world
        else:
            n_all = torch.empty(world, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(n_all, torch.tensor([n_r], dtype=torch.int64, device=dev))
            n_list = n_all.tolist()
        n_max = max(n_list)
        ragged = min(n_list) != n_max

        # the columns travel as they are, one all-gather each straight into the merged tensors: packing them into one
        # buffer (one collective) was measured slower -- the pack / unpack copies of 16 MB cost more than a second launch
        # (2 GPUs: 0.17 ms packed, 0.11 ms per column)
        cols = [onv.contiguous(), torch.view_as_real(psi).contiguous() if psi.dtype.is_complex else psi.contiguous()]
        if counts is not None:  # counts travel only when the caller has them (unit counts otherwise)
            cols.append(counts.to(torch.int64).contiguous())
        if ragged:
            cols = [torch.cat([t, t.new_zeros((n_max - n_r,) + tuple(t.shape[1:]))]) if n_r < n_max else t for t in cols]
        # NVLink peer memory when it can be set up (one pull kernel per column between two barriers), else NCCL
        outs = peer.all_gather_columns(cols, n_max) if dev.type == "cuda" else None
        if outs is None:
            outs = [torch.empty((world * n_max,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev) for t in cols]
            for o, t in zip(outs, cols):
                dist.all_gather_into_tensor(o, t)
        if ragged:  # drop the padding rows
            outs = [torch.cat([o[r * n_max : r * n_max + n_list[r]] for r in range(world)]) for o in outs]
        all_onv = outs[0]
        all_psi = torch.view_as_complex(outs[1]) if psi.dtype.is_complex else outs[1]
        all_cnt = outs[2] if counts is not None else torch.ones(all_onv.size(0), dtype=torch.int64, device=dev)
    if disjoint:
        return all_onv, all_psi, all_cnt
    if all_onv.is_cuda:
        # torch.unique(dim=0) orders rows with byte 0 most significant, i.e. as the little-endian integer of the byte-reversed
        # row: the library's radix sort + unique on the reversed rows gives the same order without torch's comparison sort
        from .C_extension import unique_onv

        rev, inv = unique_onv(all_onv.flip(1).contiguous())
        uniq = rev.flip(1).contiguous()
    else:
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


def sample_space_energy_sharded(lut: WavefunctionLUT, h1e: Tensor, h2e: Tensor, sorb: int, nele: int, noa: int, nob: int) -> Tuple[Tensor, Tensor]:
    """Sample-space local energy of THIS rank's slice of the table (rows rank_begin .. rank_end of the sorted unique set,
    split_length_idx) when the samples are the table itself -- (eloc, psi0) like eloc_sample_space on lut.bra_key[b:e].

    The work is not split by table rows but by beta string: rank r evaluates the r-th W-th of the table's beta-GROUPED copy
    (the samples that share a beta string sit next to each other there, so the block kernel keeps its full tiles at any
    world size -- a contiguous row slice would cut every string's samples into W pieces), then one all-gather hands every
    rank the energies of all samples and each picks its own rows.  One collective of 8 (16) bytes per sample."""
    from . import C_extension as ops

    rank, world = _world()
    gi = lut.group_index
    N = lut.bra_key.size(0)
    b, e = lut.rank_begin, lut.rank_end
    if world == 1:
        # one rank: the whole table, evaluated in the beta-grouped order (the samples of a beta string next to each other)
        eloc_b, _ = ops.eloc_sample_space(gi.keys(0), h1e, h2e, sorb, nele, noa, nob, lut.bra_key, lut.wf_value, gi)
        return eloc_b[gi.pos()], lut.wf_value
    ends = [0] + split_length_idx(N, world)
    n_max = max(ends[k + 1] - ends[k] for k in range(world))
    mine = gi.keys(0)[ends[rank] : ends[rank + 1]]
    eloc_part, _ = ops.eloc_sample_space(mine, h1e, h2e, sorb, nele, noa, nob, lut.bra_key, lut.wf_value, gi)
    cplx = eloc_part.is_complex()
    area = peer.PeerExchange.get(n_max * eloc_part.element_size(), eloc_part.device)
    if area is not None:
        # publish this rank's energies, then pull exactly the rows this rank owns out of the peers' memory (csrc/peer.cu):
        # 1/W of the data moves and no all-gathered copy is stored
        area.local(0, eloc_part.numel() * eloc_part.element_size()).copy_(eloc_part.view(torch.float64).view(torch.uint8) if not cplx
                                                                           else torch.view_as_real(eloc_part).reshape(-1).view(torch.uint8))
        area.barrier()
        eloc = torch.empty(e - b, dtype=eloc_part.dtype, device=eloc_part.device)
        area.gather_rows(0, gi.pos()[b:e], N, eloc)
        area.barrier()
        return eloc, lut.wf_value[b:e]
    send = torch.view_as_real(eloc_part) if cplx else eloc_part
    if send.size(0) < n_max:
        send = torch.cat([send, send.new_zeros((n_max - send.size(0),) + tuple(send.shape[1:]))])
    recv = torch.empty((world * n_max,) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
    dist.all_gather_into_tensor(recv, send.contiguous())
    if n_max * world != N:  # ragged: drop the padding rows
        recv = torch.cat([recv[k * n_max : k * n_max + ends[k + 1] - ends[k]] for k in range(world)])
    # recv is in grouped order; the group index knows where that order keeps every table row: this rank's rows
    eloc = recv[gi.pos()[b:e]]
    return (torch.view_as_complex(eloc.contiguous()) if cplx else eloc), lut.wf_value[b:e]


def _local_moments(e: Tensor, weight: Tensor, amplitude: bool) -> Tensor:
    """float64[7] = [sum w, sum w Re d, sum w Im d, sum w |d|^2, Re c, Im c, n], d = E - c, c = E[0]:
    moments about a shift of the size of the values themselves, so the variance has no cancellation.
    CUDA tensors: one kernel of the library (csrc/table.cu); CPU tensors (gloo tests): torch."""
    n = e.numel()
    if e.is_cuda:
        from .C_extension import weighted_moments

        return weighted_moments(e.contiguous(), weight.contiguous(), amplitude)
    cplx = e.is_complex()
    w = (weight.abs() ** 2 if amplitude else weight).to(torch.float64)
    if n == 0:
        return torch.zeros(7, dtype=torch.float64)
    c = e[0]
    d = e - c
    if cplx:
        rows = torch.stack([torch.ones_like(w), d.real, d.imag, d.real * d.real + d.imag * d.imag])
    else:
        rows = torch.stack([torch.ones_like(w), d, torch.zeros_like(w), d * d])
    mom = rows @ w
    tail = torch.tensor([float(c.real), float(c.imag) if cplx else 0.0, float(n)], dtype=torch.float64)
    return torch.cat([mom, tail])


class PendingStatistics:
    """Energy statistics whose per-rank moments are on the device (after the collective): `result()` does the one
    host read and the exact combination.  Lets a caller keep the D2H read outside a timed / graph-captured step."""

    def __init__(self, allv: Tensor, world: int, amplitude: bool, cplx: bool, counts: Optional[int]):
        self._allv, self._world, self._amplitude, self._cplx, self._counts = allv, world, amplitude, cplx, counts
        self._out: Optional[dict] = None

    def result(self) -> dict:
        if self._out is not None:
            return self._out
        world, cplx = self._world, self._cplx
        rows = self._allv.tolist()  # the only host synchronisation of the statistics; plain floats from here on
        z = sum(r[0] for r in rows)
        # amplitudes: p_i = |psi_i|^2 / Z * W with Z over all ranks (sample.py:772 convention)
        scale = (world / z if z > 0 else 0.0) if self._amplitude else 1.0
        parts = []
        for w, s_re, s_im, s_sq, c_re, c_im, _n in rows:
            if w == 0.0:
                continue
            dm_re, dm_im = s_re / w, s_im / w
            parts.append((w * scale, c_re + dm_re, c_im + dm_im, (s_sq - w * (dm_re * dm_re + dm_im * dm_im)) * scale))
        mean_re = math.fsum(w * mr for w, mr, _, _ in parts) / world
        mean_im = math.fsum(w * mi for w, _, mi, _ in parts) / world
        var = math.fsum(m2 + w * ((mr - mean_re) ** 2 + (mi - mean_im) ** 2) for w, mr, mi, m2 in parts) / world
        n = int(sum(r[6] for r in rows)) if self._counts is None else self._counts
        sd = max(var, 0.0) ** 0.5
        mean = complex(mean_re, mean_im) if cplx else mean_re
        self._out = {"mean": mean, "var": var, "sd": sd, "se": sd / n ** 0.5 if n else 0.0, "n": n}
        return self._out

    def __getitem__(self, k):
        return self.result()[k]


def _combine_moments(eloc: Tensor, weight: Tensor, amplitude: bool, counts: Optional[int], lazy: bool = False):
    rank, world = _world()
    cplx = eloc.is_complex()
    e = eloc.to(torch.complex128) if cplx else eloc.to(torch.float64)
    vec = _local_moments(e, weight if amplitude else weight.to(torch.float64), amplitude)
    if world > 1:
        allv = torch.empty(world * 7, dtype=torch.float64, device=vec.device)
        dist.all_gather_into_tensor(allv, vec)
        allv = allv.view(world, 7)
    else:
        allv = vec.view(1, 7)
    pending = PendingStatistics(allv, world, amplitude, cplx, counts)
    return pending if lazy else pending.result()


def energy_statistics(eloc: Tensor, prob: Tensor, counts: Optional[int] = None, lazy: bool = False):
    """mean / var / sd / se of the local energy -- the quantities of utils/stats/dist_stats.py:18-79
    (mean = sum_ranks sum_i p_i E_i / W and var = sum_ranks sum_i p_i |mean - E_i|^2 / W with
    p = prob * W, sample.py:772 + comm.py:65-67) from ONE kernel and ONE collective: an all_gather of
    the per-rank shifted moments, combined on the host with the exact identity
    sum p|E - m|^2 = sum p|E - mu|^2 + (sum p)|mu - m|^2.  lazy=True returns a PendingStatistics: the collective is
    issued, the host read happens at .result()."""
    return _combine_moments(eloc, prob, False, counts, lazy)


def energy_statistics_amplitudes(eloc: Tensor, psi0: Tensor, counts: Optional[int] = None, lazy: bool = False):
    """Same statistics with p_i = |psi0_i|^2 / Z * W, Z = sum over ALL ranks' rows of |psi0|^2 -- the
    sample-space probabilities when the ranks' slices partition the unique set (build_shared_lut).
    Z comes out of the same all_gather, so the probabilities are never materialised."""
    return _combine_moments(eloc, psi0, True, counts, lazy)
