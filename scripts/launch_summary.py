"""Aggregates an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: count, total, mean, share.
    python scripts/launch_summary.py gpurun_out/launches.csv [n_steps] > profiles/rNN_launches_*.txt"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
agg = OrderedDict()
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    n, t = agg.get(r[ki], (0, 0.0))
    agg[r[ki]] = (n + 1, t + float(r[vi].replace(",", "")) / 1e3)
total = sum(t for _, t in agg.values())
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 0
print(f"# {len(rows) - 1} launches, total {total:.1f} us" + (f" = {total / steps:.1f} us per step over {steps} steps" if steps else ""))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} n={n:4d} total={t:9.1f}us avg={t / n:8.2f}us share={100 * t / total:5.1f}%")
