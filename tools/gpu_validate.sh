#!/bin/bash
# The round-end recipe on a B200 box (run from the repo root, e.g. `gpurun --timeout 1800 -- 'bash tools/gpu_validate.sh'`):
# smoke, the GPU parity suite, the bench line, the ncu launch list of the same command and one `--set full` capture of
# the dominant kernels.  Everything lands in gpurun_out/; copy what should be judged into profiles/.
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > /dev/null 2>&1; echo "ncu launch list rc=$?"
timeout 400 ncu --set full --clock-control none -k regex:conv_planes_kernel -c 27 -f -o gpurun_out/prof_conv \
    python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > /dev/null 2>&1; echo "ncu full rc=$?"
# read here with: ncu -i gpurun_out/prof_conv.ncu-rep --page raw --csv > profiles/rN_ncu_full_conv.csv
#                 python tools/ncu_traffic.py profiles/rN_ncu_full_conv.csv > profiles/rN_conv_traffic.json
timeout 300 python tools/stream_bench.py > gpurun_out/stream_kernels.jsonl 2> gpurun_out/stream.err; echo "stream bench rc=$?"
