#!/bin/bash
# Turn the reports of profiles/run_profiles.sh <tag> (in gpurun_out/) into the tracked summaries under profiles/<tag>/.
# Runs here (no GPU needed).  Usage: bash profiles/make_summaries.sh r01
TAG=${1:-r01}
OUT=profiles/$TAG
mkdir -p $OUT
cp gpurun_out/${TAG}_launches.csv $OUT/launches.csv
( echo "ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
  echo "(6 E_loc steps (1 warm-up, 2 timed, 1 end-to-end warm-up, 2 end-to-end) of 1e6 samples = 4 batches of <= 262 144 samples each, then the API-path"
  echo " chunk timings; per-launch times are cold-cache and serialised)"; echo
  python profiles/launch_summary.py $OUT/launches.csv 2>/dev/null | head -40 ) > $OUT/launches_summary.txt
python profiles/ncu_summary.py gpurun_out/${TAG}_eloc.ncu-rep > $OUT/eloc_kernels_ncu.txt
python profiles/ncu_summary.py gpurun_out/${TAG}_api.ncu-rep > $OUT/api_kernels_ncu.txt
python profiles/ncu_summary.py gpurun_out/${TAG}_lut.ncu-rep > $OUT/lut_kernel_ncu.txt
( echo "eloc_scan_kernel<1,true,128>: warp-instructions per sample and stall samples per phase (code between barriers / calls / exits)"
  echo "python profiles/sass_phases.py gpurun_out/${TAG}_eloc.ncu-rep 262144"
  python profiles/sass_phases.py gpurun_out/${TAG}_eloc.ncu-rep 262144 ) > $OUT/eloc_scan_phases.txt
python - "$OUT" <<'PY'
import json, re, sys
out = sys.argv[1]
def grab(path):
    ks, cur = [], None
    for line in open(path):
        if line.startswith("Kernel Name"):
            cur = {"name": line.split(None, 2)[2].strip()[:60]}
            ks.append(cur)
        m = re.match(r"(dram__bytes_(read|write)\.sum|gpu__time_duration\.sum|launch__grid_size)\s+([0-9.]+)\s*(\S*)", line)
        if m and cur is not None:
            v = float(m.group(3)) * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "": 1, "us": 1, "ms": 1e3, "ns": 1e-3}.get(m.group(4), 1)
            cur[m.group(1)] = v
    return ks
e = grab(f"{out}/eloc_kernels_ncu.txt")
a = grab(f"{out}/api_kernels_ncu.txt")
l = grab(f"{out}/lut_kernel_ncu.txt")
dram = lambda k: int(k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"])
n = int(e[0]["launch__grid_size"])
t = {"source": "ncu --set full --clock-control none, bench.py Fe2S2-shaped workload, profiles/%s/*_ncu.txt (dram__bytes_read.sum + dram__bytes_write.sum per launch)" % out.split("/")[-1],
     "eloc_samples_per_launch": n, "eloc_scan_dram_bytes_per_launch": dram(e[0]), "eloc_eval_dram_bytes_per_launch": dram(e[1]),
     "eloc_dram_bytes_per_sample": round((dram(e[0]) + dram(e[1])) / n),
     "enumerate_dram_bytes_per_launch": dram(a[0]), "enumerate_samples_per_launch": 32768,
     "lut_dram_bytes_per_launch": dram(l[0]), "lut_samples_per_launch": 32768}
json.dump(t, open(f"{out}/traffic.json", "w"), indent=1)
print(t)
PY
