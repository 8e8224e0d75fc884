"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time and share.
python tools/launch_summary.py gpurun_out/x/launches.csv [first_id last_id]"""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
iK, iV, iU, iID = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
for r in rd:
    v = float(r[iV].replace(",", ""))
    u = r[iU]
    v = v / 1000.0 if u in ("ns", "nsecond") else (v * 1000.0 if u in ("ms", "msecond") else v)   # -> us
    rows.append((int(r[iID]), re.sub(r"\(.*", "", r[iK]).replace("void ", "").replace("clb::", ""), v))
if len(sys.argv) > 3:
    rows = [r for r in rows if int(sys.argv[2]) <= r[0] <= int(sys.argv[3])]
agg = collections.OrderedDict()
for _, k, v in rows:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print("%d launches, %.1f us total" % (len(rows), tot))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%6.1f us %5.1f%% x%-3d %s" % (a[1], 100 * a[1] / tot, a[0], k[:110]))
