"""aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name"""
import collections, csv, re, sys
path = sys.argv[1]
per_step = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1000.0 if row["Metric Unit"] in ("ns", "nsecond") else v
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"void |salun::", "", name)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches ({per_step} steps -> {tot/per_step:.1f} us/step)")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:64]:64s} n={c:4d} total={t:9.1f}us avg={t/c:7.1f}us per-step={t/per_step:8.1f}us share={t/tot*100:5.1f}%")
