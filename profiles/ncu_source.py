#!/usr/bin/env python
"""Per-instruction hot spots of a kernel from an .ncu-rep: python profiles/ncu_source.py rep regex [top]"""
import csv
import subprocess
import sys


def main(path, regex, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", f"regex:{regex}"],
                         capture_output=True, text=True).stdout.splitlines()
    # first kernel only
    rows = list(csv.reader(out))
    hdr = rows[1]
    body = []
    for r in rows[2:]:
        if len(r) != len(hdr):
            break
        body.append(dict(zip(hdr, r)))
    tot_inst = sum(int(d["Instructions Executed"]) for d in body)
    tot_samp = sum(int(d["# Samples"]) for d in body)
    print(f"kernel: {rows[0][1][:100]}\ninstructions(warp) {tot_inst}  samples {tot_samp}")
    f = lambda d, k: int(float(d.get(k, "0") or 0))
    print("memory instructions: idx  warp-inst  L1tagReqGlobal  smemWavefronts(ideal)  L2sectors(ideal)  source")
    for i, d in enumerate(body):
        if d["Address Space"] not in ("-", ""):
            print(f"{i:5d} {f(d,'Instructions Executed'):10d} {f(d,'L1 Tag Requests Global'):12d} {f(d,'L1 Wavefronts Shared'):10d}({f(d,'L1 Wavefronts Shared Ideal')}) "
                  f"{f(d,'L2 Theoretical Sectors Global'):12d}({f(d,'L2 Theoretical Sectors Global Ideal')})  {d['Source'].strip()[:70]}")
    print(f"\ntop {top} by stall samples: idx samples% inst  stall breakdown  source")
    keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    for i, d in sorted(enumerate(body), key=lambda t: -int(t[1]["# Samples"]))[:top]:
        st = sorted(((int(d[k]), k[6:]) for k in keys if int(d[k] or 0) > 0), reverse=True)[:3]
        print(f"{i:5d} {100*int(d['# Samples'])/max(tot_samp,1):5.1f}% {f(d,'Instructions Executed'):10d}  {st}  {d['Source'].strip()[:60]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
