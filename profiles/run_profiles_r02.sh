#!/bin/bash
# Run under gpurun (one GPU).  Usage: bash profiles/run_profiles_r02.sh <tag> [launches|eloc|api|all]
# gpurun_out/<tag>_launches.csv : per-launch device times of one short bench run (ncu gpu__time_duration.sum)
# gpurun_out/<tag>_eloc.ncu-rep : --set full capture of the one-pass kernels (block scan, per-sample scan, eval)
# gpurun_out/<tag>_api.ncu-rep  : --set full capture of enumerate_kernel and lut_indexed_kernel
TAG=${1:-r02}
WHAT=${2:-all}
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-variants"
if [ "$WHAT" = launches ] || [ "$WHAT" = all ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    $B --steps 2 --warmup 1 > gpurun_out/${TAG}_launches_bench.log 2>&1
fi
if [ "$WHAT" = eloc ] || [ "$WHAT" = all ]; then
ncu --set full --clock-control none --import-source on -k 'regex:eloc_block_kernel|eloc_eval_tile_kernel|diag_table_kernel|sample_alloc_kernel' -s 8 -c 4 -f -o gpurun_out/${TAG}_eloc \
    $B --steps 1 --warmup 1 --no-api-path --no-graph > gpurun_out/${TAG}_eloc.log 2>&1
fi
if [ "$WHAT" = api ] || [ "$WHAT" = all ]; then
ncu --set full --clock-control none --import-source on -k 'regex:enumerate_kernel' -s 4 -c 1 -f -o gpurun_out/${TAG}_api \
    $B --steps 1 --warmup 1 > gpurun_out/${TAG}_api.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:lut_indexed_kernel' -s 4 -c 1 -f -o gpurun_out/${TAG}_lut \
    $B --steps 1 --warmup 1 > gpurun_out/${TAG}_lut.log 2>&1
fi
ls -la gpurun_out | tail -6
