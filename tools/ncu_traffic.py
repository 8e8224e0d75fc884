"""profiles/r2_conv_traffic.json from an `ncu --set full` raw CSV of the bench's conv-stack launches (the capture may hold
more than one step: the first period of the kernel-name sequence is used):
    CLB_PLANES_LINEAR=0 ncu --set full --clock-control none -k regex:"conv_planes|conv1_|wgrad_reduce|bias_" -c 140 ... bench.py
    ncu -i gpurun_out/x/conv_full.ncu-rep --page raw --csv > /tmp/raw.csv ; python tools/ncu_traffic.py /tmp/raw.csv
Also prints the per-launch table (time, tensor-pipe %, DRAM bytes) that profiles/README.md quotes."""
import csv, json, os, re, sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
def get(r, name):
    v = r[col[name]].replace(",", "")
    return float(v) if v not in ("", "n/a") else 0.0
units = rows[1]
def to_bytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
ur, uw = units[col["dram__bytes_read.sum"]], units[col["dram__bytes_write.sum"]]
tot, out = 0.0, []
body = rows[2:]
names = [re.sub(r"\(.*", "", r[col["Kernel Name"]]) for r in body]
period = len(body)
for q in range(8, len(body) // 2 + 1):                    # one step = the shortest period of the launch sequence
    if names[:q] == names[q:2 * q]:
        period = q
        break
for r in body[:period]:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "")
    rd, wr = to_bytes(get(r, "dram__bytes_read.sum"), ur), to_bytes(get(r, "dram__bytes_write.sum"), uw)
    t = get(r, "gpu__time_duration.sum")
    tp = get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    tot += rd + wr
    out.append((name, t, tp, rd, wr))
    print("%-44s %8.1f %s  tensor %5.1f %%  dram rd %7.1f MB wr %7.1f MB" % (name[:44], t, units[col["gpu__time_duration.sum"]], tp, rd / 1e6, wr / 1e6))
# algorithmic bytes of the same launches (VGG-11, batch 200): every operand / result plane once
N = 200
layers = [(32, 64, 128), (16, 128, 256), (16, 256, 256), (8, 256, 512), (8, 512, 512), (4, 512, 512), (4, 512, 512)]
alg = N * 3 * 64 * 64 * 4 * 2 + N * 32 * 32 * 64 * (4 + 1) + N * 32 * 32 * 64 * (4 + 2 + 1)          # fused first layer fwd + bwd
for H, C, K in layers:
    px = N * H * H
    alg += px * C * 4 + K * 9 * C * 4 + px * K * 4                    # fwd: x planes, weight planes, y planes
    alg += px * K * 4 + K * 9 * C * 4 + px * C * 4 + px * C * 2       # dgrad: dy, weights, dx (+ mask hi plane)
    alg += px * C * 4 + px * K * 4 + K * 9 * C * 4                    # wgrad: x, dy, dw (split-K partials are extra traffic)
res = {"bytes_per_step": tot, "algorithmic_bytes_per_step": float(alg), "launches": len(out),
       "note": "sum of dram__bytes_read.sum + dram__bytes_write.sum over the %d conv-stack launches of one step (ncu --set full, "
               "profiles/r2_ncu_full_conv.csv); algorithmic = every operand / result plane once" % len(out)}
json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_conv_traffic.json"), "w"), indent=1)
print(json.dumps(res))
