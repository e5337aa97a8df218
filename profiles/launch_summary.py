#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
python profiles/launch_summary.py gpurun_out/x_launches.csv"""
import csv
import sys
from collections import OrderedDict


def main(path):
    agg = OrderedDict()
    for r in csv.reader(open(path, errors="replace")):
        if len(r) > 10 and r[0].isdigit():
            name = r[4]
            v = float(r[-1].replace(",", ""))
            unit = r[-2]
            ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
            t, c = agg.get(name, (0.0, 0))
            agg[name] = (t + ns, c + 1)
    total = sum(t for t, _ in agg.values())
    for name, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{t / 1e6:10.3f} ms {c:5d}x {t / c / 1e6:9.4f} ms/launch {100 * t / total:5.1f}%  {name[:110]}")


if __name__ == "__main__":
    main(sys.argv[1])
