export CLB_LIB_PATH=$PWD/clsurvey_b200/_lib/libclb_g3p.so
timeout 150 python tools/tc_debug.py bf16 > gpurun_out/tc_debug_g3p.log 2>&1; rc=$?; echo tc_debug rc=$rc; grep -E "mode 3.*fwd rel err|dgrad rel err|done|rror" gpurun_out/tc_debug_g3p.log | tail -14
if [ $rc -ne 0 ]; then exit 1; fi
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["value"], d["roofline"]["ms_per_step_in_kernel"])'
timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_g3p.err | python -c "$P"
unset CLB_LIB_PATH
timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P"
