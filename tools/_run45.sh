timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke3.log 2>&1; echo smoke rc=$?; tail -1 gpurun_out/smoke3.log
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu5.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/pytest_gpu5.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final3.json 2> gpurun_out/bench_final3.err; echo bench rc=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_final3.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac_of_split_ceiling'], d['roofline']['ms_per_step_in_kernel'], d['cpu_baseline']['value'])"
timeout 100 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-graph --model alexnet 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('alexnet', d['ms_per_step'], d['value'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_bf16b.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > /dev/null 2>&1; echo ncu-list rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_bf16_kernel -c 4 -f -o gpurun_out/prof_bf16_wgrad python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > /dev/null 2>&1; echo ncu-full rc=$?
