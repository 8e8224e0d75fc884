timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke2.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/smoke2.log
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu3.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_gpu3.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; echo bench rc=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_final2.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac_of_split_ceiling'], d['roofline']['ms_per_step_in_kernel'], d['cpu_baseline']['value'], d['clocks'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_bf16.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > /dev/null 2>&1; echo ncu-list rc=$?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bf16_kernel -c 10 -f -o gpurun_out/prof_bf16 python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > /dev/null 2>&1; echo ncu-full rc=$?
