"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X`) of bench.py into per-kernel totals of ONE
training step: the launches between two consecutive `patchify4_kernel` launches (the first kernel of a forward pass).
Usage: ncu_launch_summary.py launches.csv > summary.csv"""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        rows.append((r["Kernel Name"], us))
starts = [i for i, (k, _) in enumerate(rows) if "patchify4_kernel" in k]
if len(starts) < 2:
    sys.exit("need two forward passes in the capture")
step = rows[starts[0]:starts[1]]
agg = collections.OrderedDict()
for k, us in step:
    k = re.sub(r"\(.*", "", k).strip()
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
lib = sum(a[0] for k, a in agg.items() if "msu::" in k)
print(f"# one training step = {len(step)} launches ({lib} from libmsunet_sm100.so), {tot / 1e3:.2f} ms serialised under ncu "
      f"(cold caches, no stream overlap: compare SHARES, not absolutes)")
print("kernel,launches,total_us,share")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k},{a[0]},{a[1]:.1f},{a[1] / tot:.4f}")
