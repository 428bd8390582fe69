"""Per-source-line instruction / stall-sample shares of one kernel in an .ncu-rep captured with --import-source on.
    python scripts/ncu_lines.py report.ncu-rep [top_n]"""
import collections
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
hdr = rows[hi]
ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
agg = collections.OrderedDict()
file = None
for r in rows:
    if r and r[0] == "File Path":
        file = r[1].split("/")[-1]
        continue
    if len(r) < len(hdr) or not r[0].strip().isdigit():
        continue
    try:
        agg[(file, int(r[0]), r[1].strip()[:120])] = (int(r[ie] or 0), int(r[isamp] or 0))
    except ValueError:
        pass
ti, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
print(f"# total warp instructions {ti}, stall samples {ts}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{k[0]}:{k[1]:5d} inst={100 * v[0] / ti:5.1f}% samples={100 * v[1] / max(ts, 1):5.1f}%  {k[2]}")
