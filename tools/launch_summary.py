"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals
for the LAST complete training step (launches between two consecutive k_nchw_to_nhwc)."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        if unit in ("us", "usecond"): v *= 1e3
        if unit in ("ms", "msecond"): v *= 1e6
        rows.append((r["Kernel Name"], v))
starts = [i for i, (k, _) in enumerate(rows) if ("stem_im2col" in k or "nchw_to_nhwc" in k or "k_stem_pack_weights" in k)]
which = int(sys.argv[2]) if len(sys.argv) > 2 else -2
a, b = starts[which], starts[which + 1]
step = rows[a:b]
tot = collections.OrderedDict()
for k, v in step:
    k = re.sub(r"\(.*", "", k)
    t = tot.setdefault(k, [0, 0.0]); t[0] += 1; t[1] += v
total = sum(v for _, v in step)
print(f"launches in step: {len(step)}, summed kernel time {total/1e3:.1f} us")
for k, (n, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{v/1e3:9.1f} us {100*v/total:5.1f}%  x{n:<3d} {k[:90]}")
if "--list" in sys.argv:
    for k, v in step:
        print(f"{v/1e3:8.2f} {k[:100]}")
