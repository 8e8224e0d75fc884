CLB_MM_MODE=3 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_m3.log 2>&1; echo pytest-mode3 rc=$?; tail -5 gpurun_out/pytest_gpu_m3.log
