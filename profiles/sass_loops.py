#!/usr/bin/env python
"""List the loops (backward branches) of a kernel's SASS with their static body size -- a GPU-free way
to see what a source change does to the hot loops.
    python profiles/sass_loops.py pynqs_b200/csrc/eloc.o eloc_filter_kernelILi1E [min_body] [dump_index]"""
import re
import subprocess
import sys


def main(obj, pattern, min_body=20, dump=None):
    names = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", names)
    for f in funcs[1:]:
        name = f.split("\n", 1)[0].strip()
        if pattern not in name:
            continue
        ins = re.findall(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", f)
        addr = {int(a, 16): i for i, (a, _) in enumerate(ins)}
        print(f"{name[:110]}: {len(ins)} instructions")
        loops = []
        for i, (a, txt) in enumerate(ins):
            m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", txt)
            if m and int(m.group(1), 16) in addr and addr[int(m.group(1), 16)] <= i:
                loops.append((addr[int(m.group(1), 16)], i))
        for k, (lo, hi) in enumerate(loops):
            if hi - lo + 1 >= min_body:
                body = [t for _, t in ins[lo:hi + 1]]
                ops = {}
                for t in body:
                    op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]
                    ops[op] = ops.get(op, 0) + 1
                top = ", ".join(f"{o}:{c}" for o, c in sorted(ops.items(), key=lambda kv: -kv[1])[:10])
                print(f"  loop {k}: [{lo}, {hi}] {hi - lo + 1} instr   {top}")
                if dump is not None and k == dump:
                    for j in range(lo, hi + 1):
                        print(f"      {j:5d} {ins[j][1]}")
        break


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 20, int(sys.argv[4]) if len(sys.argv) > 4 else None)
