timeout 90 python tools/bf16_probe.py > gpurun_out/bf16_probe.log 2>&1; echo probe rc=$?; cat gpurun_out/bf16_probe.log | tail -4
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"], d["config"].get("launch"))'
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/n4.err | tee gpurun_out/n4.json | python -c "$P"
echo rc=$?
