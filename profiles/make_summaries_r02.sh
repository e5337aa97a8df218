#!/bin/bash
# Turn the reports of profiles/run_profiles_r02.sh <tag> (in gpurun_out/) into the tracked summaries under profiles/r02/.
# Runs here (no GPU needed).  Usage: bash profiles/make_summaries_r02.sh <tag>
TAG=${1:-r02}
OUT=profiles/r02
mkdir -p $OUT
cp gpurun_out/${TAG}_launches.csv $OUT/launches.csv
( echo "ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --no-cpu-baseline --no-variants --steps 2 --warmup 1"
  echo "(eager warm-up steps + the capture + graph replays of 1e6 samples, then the API-path chunk timings with the reference CUDA extension"
  echo " beside them; per-launch times are cold-cache and serialised)"; echo
  python profiles/launch_summary.py $OUT/launches.csv 2>/dev/null | head -45 ) > $OUT/launches_summary.txt
python profiles/ncu_summary.py gpurun_out/${TAG}_eloc.ncu-rep > $OUT/eloc_kernels_ncu.txt
python profiles/ncu_summary.py gpurun_out/${TAG}_api.ncu-rep > $OUT/api_kernels_ncu.txt
python profiles/ncu_summary.py gpurun_out/${TAG}_lut.ncu-rep > $OUT/lut_kernel_ncu.txt
( echo "eloc_block_kernel: warp-instructions per sample and stall samples per phase (code between calls / exits)"
  echo "python profiles/sass_phases.py gpurun_out/${TAG}_eloc.ncu-rep 1000000 --kernel=2"
  python profiles/sass_phases.py gpurun_out/${TAG}_eloc.ncu-rep 1000000 --kernel=2
  echo; echo "eloc_eval_tile_kernel:"
  python profiles/sass_phases.py gpurun_out/${TAG}_eloc.ncu-rep 1000000 --kernel=3 ) > $OUT/eloc_block_phases.txt
python - "$OUT" <<'PY'
import json, re, sys
out = sys.argv[1]
def grab(path):
    ks, cur = [], None
    for line in open(path):
        if line.startswith("Kernel Name"):
            cur = {"name": line.split(None, 2)[2].strip()[:60]}
            ks.append(cur)
        m = re.match(r"(dram__bytes_(read|write)\.sum|gpu__time_duration\.sum|launch__grid_size|smsp__inst_executed\.sum)\s+([0-9.]+)\s*(\S*)", line)
        if m and cur is not None:
            v = float(m.group(3)) * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "": 1, "us": 1, "ms": 1e3, "ns": 1e-3, "inst": 1}.get(m.group(4), 1)
            cur[m.group(1)] = v
    return ks
e = grab(f"{out}/eloc_kernels_ncu.txt")
a = grab(f"{out}/api_kernels_ncu.txt") + grab(f"{out}/lut_kernel_ncu.txt")
dram = lambda k: int(k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"])
n = 1_000_000
enum = next(k for k in a if "enumerate" in k["name"])
lut = next(k for k in a if "lut_indexed" in k["name"])
t = {"source": "ncu --set full --clock-control none, bench.py Fe2S2 workload (1e6 samples per call), profiles/r02/*_ncu.txt",
     "eloc_samples_per_launch": n,
     "eloc_kernels": [{"name": k["name"], "us": k["gpu__time_duration.sum"], "dram_bytes": dram(k), "warp_instructions": int(k["smsp__inst_executed.sum"])} for k in e],
     "eloc_dram_bytes_per_sample": round(sum(dram(k) for k in e) / n),
     "eloc_warp_instructions_per_sample": round(sum(k["smsp__inst_executed.sum"] for k in e) / n),
     "enumerate_dram_bytes_per_launch": dram(enum), "enumerate_samples_per_launch": 32768,
     "lut_dram_bytes_per_launch": dram(lut), "lut_samples_per_launch": 32768}
json.dump(t, open(f"{out}/traffic.json", "w"), indent=1)
print(t)
PY
