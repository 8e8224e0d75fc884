P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"], d["config"].get("launch"))'
i=0
for extra in "" "--no-graph-dp" "--no-graph-dp --no-overlap"; do
i=$((i+1))
timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$i bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline $extra 2>gpurun_out/n2_$i.err | tee gpurun_out/n2_$i.json | python -c "$P"
echo rc=$? "($extra)"
done
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/dp_parity.py > gpurun_out/dp_parity2.log 2>&1
echo dp_parity rc=$?; tail -5 gpurun_out/dp_parity2.log
