"""world_size-2 (and 3) CPU tests of the multi-rank plumbing with the gloo backend: the sample
exchange and the fused energy statistics must reproduce the serial result and the reference
conventions (prob * W then / W; merged order = torch.unique order; slices by split_length_idx)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pynqs_b200 import synthetic as S


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, overlap, cplx, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pynqs_b200.distributed import energy_statistics, energy_statistics_amplitudes, exchange_unique_samples, rank_slice

        keys = S.random_onvs(1000, 40, 15, 15, seed=100)
        psi_all = S.random_psi(1000, seed=101, complex_=cplx)
        # ragged per-rank pieces; with `overlap` ranks share 100 samples (the non-disjoint case)
        cuts = [0, 334, 667, 1000] if world == 3 else [0, 450, 1000]
        lo, hi = cuts[rank], cuts[rank + 1]
        sel = np.arange(lo, hi)
        if overlap:
            sel = np.unique(np.concatenate([sel, np.arange(400, 500)]))
        onv = torch.from_numpy(keys[sel])
        psi = torch.from_numpy(psi_all[sel])
        counts = torch.from_numpy(np.arange(1, len(sel) + 1, dtype=np.int64))
        uniq, wf, cnt = exchange_unique_samples(onv, psi, counts, disjoint=not overlap)
        b, e = rank_slice(uniq.size(0))
        # energy statistics on this rank's slice with the reference's prob * world convention
        rng = np.random.default_rng(7)
        eloc_all = rng.standard_normal(uniq.size(0)) - 116.6
        if cplx:
            eloc_all = eloc_all + 1j * rng.standard_normal(uniq.size(0)) * 1e-3
        prob_all = cnt.numpy() / cnt.numpy().sum()
        st = energy_statistics(torch.from_numpy(eloc_all[b:e]), torch.from_numpy(prob_all[b:e]) * world)
        st_amp = energy_statistics_amplitudes(torch.from_numpy(eloc_all[b:e]), wf[b:e].contiguous())
        # equal pieces without counts: no size handshake, unit counts
        u2, w2, c2 = exchange_unique_samples(onv[:300].contiguous(), psi[:300].contiguous(), None, disjoint=True, equal_sizes=True)
        assert u2.shape == (300 * world, onv.size(1)) and torch.equal(u2[300 * rank : 300 * (rank + 1)], onv[:300])
        assert torch.equal(w2[300 * rank : 300 * (rank + 1)], psi[:300]) and int(c2.sum()) == 300 * world
        q.put((rank, uniq.numpy(), wf.numpy(), cnt.numpy(), (b, e), st, st_amp))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,overlap,cplx", [(2, False, False), (2, True, True), (3, True, False)])
def test_exchange_and_statistics(world, overlap, cplx):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, overlap, cplx, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    keys = S.random_onvs(1000, 40, 15, 15, seed=100)
    psi_all = S.random_psi(1000, seed=101, complex_=cplx)
    cuts = [0, 334, 667, 1000] if world == 3 else [0, 450, 1000]
    sels = []
    for r in range(world):
        sel = np.arange(cuts[r], cuts[r + 1])
        if overlap:
            sel = np.unique(np.concatenate([sel, np.arange(400, 500)]))
        sels.append(sel)
    cat_idx = np.concatenate(sels)
    cat_cnt = np.concatenate([np.arange(1, len(s) + 1) for s in sels])
    if overlap:
        # reference merge (vmc/sample.py:672-688): torch.unique(dim=0) order, psi of the first occurrence, counts summed
        uq, inv = torch.unique(torch.from_numpy(keys[cat_idx]), dim=0, return_inverse=True)
        want_keys = uq.numpy()
        want_cnt = np.zeros(len(want_keys), dtype=np.int64)
        np.add.at(want_cnt, inv.numpy(), cat_cnt)
        first = np.full(len(want_keys), len(cat_idx))
        np.minimum.at(first, inv.numpy(), np.arange(len(cat_idx)))
        want_psi = psi_all[cat_idx][first]
    else:
        want_keys, want_cnt, want_psi = keys[cat_idx], cat_cnt, psi_all[cat_idx]

    rng = np.random.default_rng(7)
    eloc_all = rng.standard_normal(len(want_keys)) - 116.6
    if cplx:
        eloc_all = eloc_all + 1j * rng.standard_normal(len(want_keys)) * 1e-3
    prob_all = want_cnt / want_cnt.sum()
    mean = np.sum(prob_all * eloc_all)
    var = np.sum(prob_all * np.abs(eloc_all - mean) ** 2)
    covered = []
    p_amp = np.abs(want_psi) ** 2 / np.sum(np.abs(want_psi) ** 2)
    mean_amp = np.sum(p_amp * eloc_all)
    var_amp = np.sum(p_amp * np.abs(eloc_all - mean_amp) ** 2)
    for rank, uniq, wf, cnt, (b, e), st, st_amp in res:
        assert abs(st_amp["mean"] - mean_amp) < 1e-12 and abs(st_amp["var"] - var_amp) < 1e-12 * max(1.0, var_amp)
        np.testing.assert_array_equal(uniq, want_keys)      # every rank holds the identical table
        np.testing.assert_array_equal(wf, want_psi)
        np.testing.assert_array_equal(cnt, want_cnt)
        covered.append((b, e))
        assert abs(st["mean"] - mean) < 1e-12                # mean energy: well inside 1e-10 Ha
        assert abs(st["var"] - var) < 1e-12 * max(1.0, var)
        assert st["n"] == len(want_keys)
        assert abs(st["se"] - np.sqrt(var / len(want_keys))) < 1e-13
    ends = [0]
    qq, rr = divmod(len(want_keys), world)
    for i in range(world):
        ends.append(ends[-1] + qq + (1 if i < rr else 0))
    assert covered == [(ends[i], ends[i + 1]) for i in range(world)]


def _lut_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pynqs_b200.distributed import build_shared_lut

        keys = S.random_onvs(900, 40, 15, 15, seed=200)
        psi = S.random_psi(900, seed=201)
        cuts = [0, 250, 900]
        sel = np.arange(cuts[rank], cuts[rank + 1])
        counts = torch.from_numpy(np.arange(1, len(sel) + 1, dtype=np.int64))
        x, prob, lut = build_shared_lut(torch.from_numpy(keys[sel]), torch.from_numpy(psi[sel]), 40, counts, disjoint=True)
        q.put((rank, x.numpy(), prob.numpy(), lut.bra_key.numpy(), lut.wf_value.numpy(), lut.rank_begin, lut.rank_end))
    finally:
        dist.destroy_process_group()


def test_build_shared_lut_same_table_everywhere_and_slices_partition():
    """build_shared_lut (the replacement of Sampler.gather_scatter_sample, vmc/sample.py:627-772): every rank ends
    up with the same sorted table, the ranks' slices partition the merged unique set and prob carries the
    reference's prob * world_size convention."""
    from oracle import oracle as O

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_lut_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    keys = S.random_onvs(900, 40, 15, 15, seed=200)
    psi = S.random_psi(900, seed=201)
    order = O.sort_onv(keys)
    cnt = np.concatenate([np.arange(1, 251), np.arange(1, 651)]).astype(np.float64)
    xs, probs = [], []
    for rank, x, prob, tab_keys, tab_psi, rb, re_ in res:
        np.testing.assert_array_equal(tab_keys, keys[order])   # the reference's sorted order, identical on every rank
        np.testing.assert_array_equal(tab_psi, psi[order])
        assert (rb, re_) == ((0, 450) if rank == 0 else (450, 900))
        xs.append(x)
        probs.append(prob)
    np.testing.assert_array_equal(np.concatenate(xs), keys)    # disjoint pieces: concatenation in rank order
    np.testing.assert_allclose(np.concatenate(probs), cnt / cnt.sum() * world, rtol=1e-15)
