timeout 150 python tools/tc_debug.py bf16 > gpurun_out/tc_debug_bf16c.log 2>&1; rc=$?; echo tc_debug rc=$rc; grep -E "wgrad rel err|done|rror" gpurun_out/tc_debug_bf16c.log | tail -14
if [ $rc -ne 0 ]; then exit 1; fi
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["value"], d["roofline"]["ms_per_step_in_kernel"])'
timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_m3b.err | python -c "$P"
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q -x > gpurun_out/pytest_gpu4.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_gpu4.log
