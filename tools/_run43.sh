P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["roofline"]["ms_per_step_in_kernel"])'
for v in noa nomma nosplit; do
echo "variant $v"
CLB_LIB_PATH=$PWD/clsurvey_b200/_lib/libclb_$v.so timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P"
done
