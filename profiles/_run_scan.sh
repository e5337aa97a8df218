mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
(timeout 500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-api-path 2>&1 | tail -1) > gpurun_out/bench.log 2>&1
python -c "
import json;d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['phases_ms_rank0'],d['energy'],d['gpu_launches'])"
ncu --set full --clock-control none --import-source on -k regex:eloc_scan -s 2 -c 1 -f -o gpurun_out/r02f_scan python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-api-path > gpurun_out/r02f_scan.log 2>&1
