"""Per-kernel totals from an ncu `--metrics gpu__time_duration.sum --csv` launch list:  python tools/launch_summary.py file.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, v, r.get("Grid Size", ""), r.get("Block Size", "")))
tot = sum(v for _, v, _, _ in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, v, _, _ in rows:
    agg[n][0] += 1
    agg[n][1] += v
print(f"{len(rows)} launches, {tot / 1e3:.3f} ms total (cold-cache, serialised)")
for n, (cnt, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v / 1e3:9.3f} ms {100 * v / tot:5.1f}%  x{cnt:<4d} {n}")
if len(sys.argv) > 2:
    print("--- top launches")
    for n, v, g, b in sorted(rows, key=lambda r: -r[1])[:int(sys.argv[2])]:
        print(f"{v:9.1f} us  grid {g} block {b}  {n}")
