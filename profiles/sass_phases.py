#!/usr/bin/env python
"""Instruction / stall-sample share of each phase of a kernel (phases = code between barriers, calls and exits),
from an .ncu-rep captured with --import-source on:  python profiles/sass_phases.py rep [samples_per_launch] [lo hi] [--kernel=K]
With lo hi: per-instruction listing (executions per sample) of SASS lines lo..hi."""
import csv
import subprocess
import sys


def main(path, per=1, lo=None, hi=None, kernel=0):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
    rows = list(csv.reader(out))
    heads = [i for i, r in enumerate(rows) if "Instructions Executed" in r]  # one section per profiled launch
    hi_ = heads[kernel]
    hdr = rows[hi_]
    ie, src, smp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    body = []
    for r in rows[hi_ + 1:]:  # this launch's section only
        if len(r) != len(hdr) or r == hdr:
            break
        body.append(r)
    T = sum(int(r[ie]) for r in body)
    S = sum(int(r[smp]) for r in body)
    print(f"total warp-instructions {T} ({T / per:.0f} per unit), stall samples {S}, SASS lines {len(body)}")
    if lo is not None:
        for i in range(lo, hi + 1):
            r = body[i]
            print(f"{i:5d} {int(r[ie]) / per:8.1f} {int(r[smp]):6d}  {r[src].strip()[:90]}")
        return
    acc = accs = 0
    start = 0
    for i, r in enumerate(body):
        acc += int(r[ie])
        accs += int(r[smp])
        t = r[src]
        if "BAR.SYNC" in t or "CALL" in t or "EXIT" in t or "RET" in t or i == len(body) - 1:
            if acc > T * 0.003:
                print(f"[{start:5d},{i:5d}] {100 * acc / T:5.1f}% inst ({acc / per:7.0f}/unit) {100 * accs / max(S, 1):5.1f}% samples   ends: {t.strip()[:50]}")
            acc = accs = 0
            start = i + 1


if __name__ == "__main__":
    a = [x for x in sys.argv if not x.startswith("--kernel=")]
    k = next((int(x.split("=")[1]) for x in sys.argv if x.startswith("--kernel=")), 0)  # --kernel=K: K-th launch of the report
    main(a[1], float(a[2]) if len(a) > 2 else 1, int(a[3]) if len(a) > 4 else None, int(a[4]) if len(a) > 4 else None, k)
