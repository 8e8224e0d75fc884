timeout 200 python tools/tc_debug.py bf16 > gpurun_out/tc_debug_bf16.log 2>&1; echo rc=$?; grep -E "rel err|done|Error|error" gpurun_out/tc_debug_bf16.log | tail -40
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["value"], d["roofline"]["ms_per_step_in_kernel"])'
for m in 3 1; do timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --mm-mode $m 2>gpurun_out/bench_m$m.err | python -c "$P"; done
