"""Quick A/B of the two API kernels (get_comb_hij_fused, wavefunction_lut) under the library's tuning knobs.

    python profiles/micro/api_quick.py [knob=value ...]

Times both operators on 32 768 Fe2S2 samples (the bench's API chunk) with every knob combination given on the command line
(lut_pipeline on / off), and checks the outputs bit for bit against the knobs-off run."""
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pynqs_b200 import C_extension as ops  # noqa: E402
from pynqs_b200 import _lib  # noqa: E402
from pynqs_b200.lut import WavefunctionLUT  # noqa: E402

dev = torch.device("cuda", 0)
SORB, NOA, NOB, NELE = bench.SORB, bench.NOA, bench.NOB, bench.NELE


def t(fn, reps=9):
    for _ in range(3):
        out = fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts), out


def main():
    chunk = 32768
    keys = bench.make_table("uniform", 1_000_000)
    psi = bench.make_psi(keys.shape[0], False)
    h1e_np, h2e_np, _ = bench.load_integrals()
    h1e, h2e = torch.from_numpy(h1e_np).to(dev), torch.from_numpy(h2e_np).to(dev)
    lut = WavefunctionLUT(torch.from_numpy(keys).to(dev), torch.from_numpy(psi).to(dev), SORB, dev, rank=0, world_size=1)
    x = lut.bra_key[:chunk]
    prep = ops.PreparedIntegrals(h2e, SORB)
    M = ops.get_Num_SinglesDoubles(SORB, NOA, NOB) + 1
    peak = bench.hbm_peak_gbs()[0]
    fb, lb = (16 * M + 8) * chunk, 17 * M * chunk
    _lib.set_tuning()
    _lib.set_tuning("lut_pipeline", 0)
    f0, (comb0, hmat0) = t(lambda: ops.get_comb_hij_fused(x, h1e, h2e, SORB, NELE, NOA, NOB, prepared=prep))
    flat = comb0.view(-1, 8)
    l0, (idx0, mask0) = t(lambda: ops.wavefunction_lut(lut.bra_key, flat, SORB, hash_index=lut.hash_index))
    print(f"baseline (knobs off): fused {f0:.4f} ms = {fb / f0 / 1e6 / peak:.3f} of HBM peak, lut {l0:.4f} ms = {lb / l0 / 1e6 / peak:.3f}", flush=True)
    _lib.set_tuning("lut_pipeline", 1)
    l1, (idx1, mask1) = t(lambda: ops.wavefunction_lut(lut.bra_key, flat, SORB, hash_index=lut.hash_index))
    same = bool(torch.equal(idx0, idx1) and torch.equal(mask0, mask1))
    print(f"lut_pipeline=1: lut {l1:.4f} ms = {lb / l1 / 1e6 / peak:.3f} of HBM peak, identical {same}, hits {int(mask1.sum())}", flush=True)
    _lib.set_tuning()


if __name__ == "__main__":
    main()
