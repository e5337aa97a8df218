mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --durations=5 2>&1 | tail -25) > gpurun_out/pytest_gpu.log 2>&1; tail -25 gpurun_out/pytest_gpu.log
