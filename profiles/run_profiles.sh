#!/bin/bash
# Run under gpurun (one GPU).  Usage: bash profiles/run_profiles.sh <tag>
# Produces gpurun_out/<tag>_launches.csv (per-launch device times of one bench run) and
# gpurun_out/<tag>_{eloc,api,lut}.ncu-rep (--set full captures of the dominant kernels).
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:eloc_scan_kernel|eloc_eval_kernel' -s 2 -c 2 -f -o gpurun_out/${TAG}_eloc \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-api-path > gpurun_out/${TAG}_eloc.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:enumerate_kernel' -s 4 -c 1 -f -o gpurun_out/${TAG}_api \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_api.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:lut_indexed_kernel' -s 4 -c 1 -f -o gpurun_out/${TAG}_lut \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_lut.log 2>&1
ls -la gpurun_out | tail -8
