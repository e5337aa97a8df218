mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
(PYNQS_SCAN_THREADS=64 timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k eloc 2>&1 | tail -15) > gpurun_out/pytest_gpu64.log 2>&1; tail -4 gpurun_out/pytest_gpu64.log
for t in 64 128 256; do
(PYNQS_SCAN_THREADS=$t timeout 500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-api-path 2>&1 | tail -1) > gpurun_out/bench_$t.log 2>&1
python -c "
import json,sys;d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]);print(sys.argv[1],d['value'],d['ms_per_step'],d['phases_ms_rank0']['eloc_kernels'],d['energy']['mean'])" gpurun_out/bench_$t.log
done
