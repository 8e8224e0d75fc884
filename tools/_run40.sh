timeout 150 python tools/tc_debug.py bf16 > gpurun_out/tc_debug_bf16b.log 2>&1; rc=$?; echo tc_debug rc=$rc; grep -E "rel err|done|rror" gpurun_out/tc_debug_bf16b.log | tail -34
if [ $rc -ne 0 ]; then exit 1; fi
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["value"], d["roofline"]["ms_per_step_in_kernel"])'
timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --mm-mode 3 2>gpurun_out/bench_m3.err | python -c "$P"
CLB_MM_MODE=3 timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_m3.log 2>&1; echo pytest-mode3 rc=$?; tail -8 gpurun_out/pytest_gpu_m3.log
