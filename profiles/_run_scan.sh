mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
(timeout 500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench.log 2>&1
python -c "
import json;d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['phases_ms_rank0'],d['energy'],d['gpu_launches']);print([(k['kernel'][:20],k['ms'],k['frac']) for k in d['roofline_hbm_kernels']])"
