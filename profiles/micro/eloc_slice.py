"""E_loc of one rank's share of the 10^6-sample Fe2S2 set on ONE GPU: what rank r of W ranks runs inside the step
(profiles the per-rank regime of the multi-GPU runs without a multi-GPU box).

    python profiles/micro/eloc_slice.py [W ...]       # default 1 2 4 8
Prints, per W, the CUDA-event time of pynqs_eloc_sample_space on N / W samples of the beta-grouped copy against the full table,
and of the table build (sort + grouped copies)."""
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pynqs_b200 import C_extension as ops  # noqa: E402
from pynqs_b200 import _lib  # noqa: E402
from pynqs_b200.lut import WavefunctionLUT  # noqa: E402

dev = torch.device("cuda", 0)
SORB, NOA, NOB, NELE = bench.SORB, bench.NOA, bench.NOB, bench.NELE


def timed(fn, reps=7, flush=None, pre=None):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.fill_(1)
        if pre is not None:
            pre()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def main():
    worlds = [int(a) for a in sys.argv[1:] if a.isdigit()] or [1, 2, 4, 8]
    for a in sys.argv[1:]:
        if "=" in a:
            k, v = a.split("=")
            _lib.set_tuning(k, int(v))
    keys = bench.make_table("uniform", 1_000_000)
    psi = bench.make_psi(keys.shape[0], False)
    h1e_np, h2e_np, _ = bench.load_integrals()
    h1e, h2e = torch.from_numpy(h1e_np).to(dev), torch.from_numpy(h2e_np).to(dev)
    d_keys, d_psi = torch.from_numpy(keys).to(dev), torch.from_numpy(psi).to(dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def build():
        lut = WavefunctionLUT(d_keys, d_psi, SORB, dev, rank=0, world_size=1)
        lut.group_index
        return lut

    t_build = timed(build, flush=flush)
    lut = build()
    gi = lut.group_index
    n = keys.shape[0]
    # like inside the step: L2 flushed, then the table is built (which leaves its copies in L2), then E_loc is timed
    state = {"lut": lut, "gi": gi}

    def rebuild():
        state["lut"] = build()
        state["gi"] = state["lut"].group_index

    print(f"table build (sort + grouped copies), eager launches: {t_build:.3f} ms", flush=True)
    base = None
    for w in worlds:
        t = timed(lambda: ops.eloc_sample_space(state["gi"].keys(0)[: n // w], h1e, h2e, SORB, NELE, NOA, NOB, state["lut"].bra_key,
                                                state["lut"].wf_value, state["gi"]), flush=flush, pre=rebuild)
        base = base or t
        print(f"W = {w}: E_loc of {n // w} samples {t:.3f} ms  ({n // w / t / 1e3:.1f} M samples/s per rank; {base / t:.2f}x of W = 1)", flush=True)


if __name__ == "__main__":
    main()
