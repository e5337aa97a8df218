#!/usr/bin/env python
"""Throughput of the BASELINE.json configurations bench.py does not run (bench.py is configs[1], Fe2S2):

    c1  H6 STO-3G, 12 spin orbitals, 3a3b: 1e4 samples (with repeats) of the 400-determinant space
    c3  N2 cc-pVDZ frozen core, 52 spin orbitals, 5a5b (M = 15 436): 1e6 unique samples, sharded over the ranks
    c4  H50 STO-3G, 100 spin orbitals (two-word ONVs), 25a25b (M = 571 876): 1e5 unique samples, sharded
    c5  192 spin orbitals (three-word ONVs), 4a4b (M = 186 393), random 8-fold-symmetric integrals (h2e 1.345 GB):
        1e7 ONVs streamed through the one-pass local energy; H_ij throughput of the materialising operator

    python profiles/bench_configs.py --config c3 [--samples N] [--steps K]
    python -m torch.distributed.run --nproc-per-node 8 ... profiles/bench_configs.py --config c4

Integrals: random 8-fold-symmetric (pyscf is not installed; BASELINE.json allows them for the named orbital count).
One step = exchange -> table sort + grouped copies -> one-pass E_loc (work split by beta string) -> statistics, like
bench.py; per-step CUDA events, L2 flushed between steps, max over ranks.  On one rank it also times the materialising
operators on a chunk (rows/s, fraction of the HBM peak), with the integrals pinned in a persisting-L2 window and without,
and checks a few samples against the CPU oracle.  Prints one JSON line per run (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CONFIGS = {
    "c1": dict(name="H6 STO-3G", sorb=12, noA=3, noB=3, samples=10_000, repeats=True),
    "c3": dict(name="N2 cc-pVDZ frozen core", sorb=52, noA=5, noB=5, samples=1_000_000),
    "c4": dict(name="H50 STO-3G", sorb=100, noA=25, noB=25, samples=100_000),
    "c5": dict(name="192-sorb microbenchmark", sorb=192, noA=4, noB=4, samples=10_000_000),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=sorted(CONFIGS))
    ap.add_argument("--samples", type=int, default=0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--oracle-samples", type=int, default=4)
    ap.add_argument("--no-api", action="store_true")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from pynqs_b200 import C_extension as ops
    from pynqs_b200 import _lib
    from pynqs_b200 import synthetic as S
    from pynqs_b200.distributed import energy_statistics_amplitudes, exchange_unique_samples, sample_space_energy_sharded
    from pynqs_b200.lut import WavefunctionLUT, split_length_idx

    _lib.load()
    sorb, noA, noB = cfg["sorb"], cfg["noA"], cfg["noB"]
    nele, L = noA + noB, (sorb - 1) // 64 + 1
    n_want = args.samples or cfg["samples"]
    M = ops.get_Num_SinglesDoubles(sorb, noA, noB) + 1
    if n_want > 2_000_000:  # in pieces: the generator holds an [n, sorb] array
        parts = [S.random_onvs(1_000_000, sorb, noA, noB, seed=1234 + i) for i in range(-(-n_want // 1_000_000) + 1)]
        keys = np.concatenate(parts)
        w = keys.shape[1]
        keys = np.unique(keys.view(np.dtype((np.void, w)))).view(np.uint8).reshape(-1, w)  # rows as opaque records
        keys = np.ascontiguousarray(keys[np.random.default_rng(99).permutation(keys.shape[0])[:n_want]])
    else:
        keys = S.random_onvs(n_want, sorb, noA, noB, seed=1234)  # unique; small spaces: the whole space
    psi = S.random_psi(keys.shape[0], seed=1235)
    h1e_np, h2e_np = S.random_packed_integrals(sorb, seed=7, symmetric=True)
    h1e, h2e = torch.from_numpy(h1e_np).to(dev), torch.from_numpy(h2e_np).to(dev)
    N = keys.shape[0]
    peak = 6538.3
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    cuts = [0] + split_length_idx(N, world)
    d_keys = torch.from_numpy(keys[cuts[rank] : cuts[rank + 1]]).to(dev)
    d_psi = torch.from_numpy(psi[cuts[rank] : cuts[rank + 1]]).to(dev)
    keep = {}

    def step():
        uniq, wf, _ = exchange_unique_samples(d_keys, d_psi, None, disjoint=True, equal_sizes=N % world == 0)
        lut = WavefunctionLUT(uniq, wf, sorb, dev, rank=rank, world_size=world)
        if cfg.get("repeats"):  # c1: more samples than determinants -- every determinant evaluated several times
            reps = -(-n_want // (N * world))
            x = lut.bra_key[lut.rank_begin : lut.rank_end].repeat(reps, 1)
            eloc, psi0 = ops.eloc_sample_space(x, h1e, h2e, sorb, nele, noA, noB, lut.bra_key, lut.wf_value, lut.group_index)
        else:
            eloc, psi0 = sample_space_energy_sharded(lut, h1e, h2e, sorb, nele, noA, noB)
        st = energy_statistics_amplitudes(eloc, psi0, lazy=True)
        keep.update(eloc=eloc, lut=lut, n_eval=eloc.numel())
        return st

    for _ in range(args.warmup):
        st = step()
    torch.cuda.synchronize()
    total_ms = 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        st = step()
        b.record()
        torch.cuda.synchronize()
        total_ms += a.elapsed_time(b)
    tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    n_eval = torch.tensor([float(keep["n_eval"])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(n_eval)
    ms = float(tt.item()) / args.steps
    st = st.result()
    line = {"config": args.config, "name": cfg["name"], "sorb": sorb, "noA": noA, "noB": noB, "L": L, "M": M, "n_gpus": world,
            "table_keys": N, "samples_per_step": int(n_eval.item()), "ms_per_step": ms, "samples_per_s": n_eval.item() / ms * 1e3,
            "connected_pairs_per_s": n_eval.item() * M / ms * 1e3, "steps": args.steps, "energy_mean": st["mean"],
            "integrals": "random 8-fold symmetric, seed 7", "h2e_bytes": int(h2e.numel() * 8),
            "step": "exchange + table sort + grouped table + one-pass E_loc + statistics; L2 flushed between steps"}

    if rank == 0 and args.oracle_samples:
        from oracle import oracle as O

        lut = keep["lut"]
        k = args.oracle_samples
        x = lut.bra_key[lut.rank_begin : lut.rank_begin + k].contiguous()
        got, _ = ops.eloc_sample_space(x, h1e, h2e, sorb, nele, noA, noB, lut.bra_key, lut.wf_value, lut.group_index)
        want = O.eloc_sample_space(x.cpu().numpy(), h1e_np, h2e_np, lut.bra_key.cpu().numpy(), lut.wf_value.cpu().numpy(), sorb, nele, noA, noB)
        rel = float(np.max(np.abs(got.cpu().numpy() - want) / np.abs(want)))
        line["parity"] = {"n": k, "against": "CPU oracle (oracle/pynqs_oracle.c), same inputs", "max_rel_err": rel, "ok": bool(rel <= 1e-12)}

    if rank == 0 and world == 1 and not args.no_api:
        lut = keep["lut"]
        chunk = int(max(1, min(N, (12 << 30) // (M * (8 * L + 8)))))
        x = lut.bra_key[:chunk].contiguous()

        def t(fn, reps=5):
            for _ in range(2):
                fn()
            ts = []
            for _ in range(reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            return statistics.median(ts)

        api = {"chunk_samples": chunk, "rows": chunk * M}
        fb = (8 * L + 8) * M * chunk
        prep_bytes = ops.prepared_nbytes(sorb, torch.float64)
        prep = ops.PreparedIntegrals(h2e, sorb)
        for label, prepared, pin in (("prepared", prep, None), ("prepared_l2_window", prep, prep.workspace), ("packed_h2e", False, None),
                                     ("packed_h2e_l2_window", False, h2e)):
            granted = ops.pin_in_l2(pin) if pin is not None else (0, 0)
            f_ms = t(lambda: ops.get_comb_hij_fused(x, h1e, h2e, sorb, nele, noA, noB, prepared=prepared))
            if pin is not None:
                ops.pin_in_l2(None)
            api[label] = {"ms": f_ms, "rows_per_s": chunk * M / f_ms * 1e3, "GBps": fb / f_ms / 1e6, "frac_of_hbm_peak": fb / f_ms / 1e6 / peak,
                          "l2_set_aside_bytes": granted[0], "l2_window_bytes": granted[1]}
        api["prepared_bytes"] = prep_bytes
        api["algorithmic_bytes_per_launch"] = fb
        comb, _ = ops.get_comb_hij_fused(x, h1e, h2e, sorb, nele, noA, noB, prepared=prep)
        flat = comb.view(-1, 8 * L)
        l_ms = t(lambda: ops.wavefunction_lut(lut.bra_key, flat, sorb))
        lb = (8 * L + 9) * M * chunk
        api["wavefunction_lut"] = {"ms": l_ms, "lookups_per_s": chunk * M / l_ms * 1e3, "GBps": lb / l_ms / 1e6, "frac_of_hbm_peak": lb / l_ms / 1e6 / peak}
        line["api_kernels"] = api
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:  # (no destroy_process_group: it can block after the symmetric-memory rendezvous -- see bench.leave)
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
