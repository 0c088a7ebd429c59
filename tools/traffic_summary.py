"""Summarise an ncu launch list taken with
   --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
(tools/run_profile.sh): per-kernel duration and DRAM bytes for one complete training step, and the
JSON bench.py reads for roofline.traffic.   usage: traffic_summary.py list.csv out.json [step index]"""
import csv, sys, collections, re, json

def scale(v, unit):
    u = unit.lower()
    m = {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9,
         "byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}
    return v * m.get(u, 1)

with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
launches = collections.OrderedDict()
for r in csv.DictReader(lines):
    d = launches.setdefault(int(r["ID"]), {"name": re.sub(r"\(.*", "", r["Kernel Name"])})
    d[r["Metric Name"]] = scale(float(r["Metric Value"].replace(",", "")), r.get("Metric Unit", ""))
rows = list(launches.values())
starts = [i for i, r in enumerate(rows) if "nchw_to_nhwc" in r["name"] or "k_stem_pack_weights" in r["name"]]
if len(starts) < 2:   # fall back: a step starts at the first k_conv after a k_sgd
    starts = [i for i, r in enumerate(rows) if i and "k_sgd" in rows[i - 1]["name"]]
which = int(sys.argv[3]) if len(sys.argv) > 3 else -2
a, b = starts[which], starts[which + 1]
step = rows[a:b]
tot = collections.OrderedDict()
for r in step:
    t = tot.setdefault(r["name"], [0, 0.0, 0.0, 0.0])
    t[0] += 1; t[1] += r.get("gpu__time_duration.sum", 0); t[2] += r.get("dram__bytes_read.sum", 0); t[3] += r.get("dram__bytes_write.sum", 0)
T = sum(v[1] for v in tot.values()); R = sum(v[2] for v in tot.values()); W = sum(v[3] for v in tot.values())
print(f"launches in step: {len(step)}; summed kernel time {T/1e3:.1f} us; DRAM read {R/1e6:.1f} MB, write {W/1e6:.1f} MB")
print(f"{'us':>9s} {'%':>5s} {'n':>4s} {'rd MB':>8s} {'wr MB':>8s}  kernel")
for k, (n, t, rd, wr) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{t/1e3:9.1f} {100*t/T:5.1f} {n:4d} {rd/1e6:8.1f} {wr/1e6:8.1f}  {k[:80]}")
conv = [v for k, v in tot.items() if k.startswith("k_conv") or "k_conv<" in k]
cn = sum(v[0] for v in conv); cb = sum(v[2] + v[3] for v in conv)
out = {
    "source": sys.argv[1].split("/")[-1],
    "step_launches": len(step), "step_kernel_time_us": T / 1e3,
    "step_dram_read_bytes": R, "step_dram_write_bytes": W,
    "conv_launches_per_step": cn,
    "conv_dram_bytes_per_launch": cb / max(cn, 1),
    "per_kernel": {k: {"n": n, "us": t / 1e3, "dram_read_bytes": rd, "dram_write_bytes": wr} for k, (n, t, rd, wr) in tot.items()},
    "note": "warm graph-replay step under ncu --cache-control none (L2 state as the step leaves it): dram__bytes_read.sum + "
            "dram__bytes_write.sum averaged over the k_conv (fprop+dgrad, fused BatchNorm) launches of one training step",
}
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
