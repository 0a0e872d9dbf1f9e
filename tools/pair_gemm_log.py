"""Pair the SALUN_GEMM_LOG=1 stderr lines of a run with the ncu launch list of the same run (launch order) and print a
per-shape table of the tensor-core kernels: launches, total time, TFLOP/s."""
import collections, csv, re, sys
log, path = sys.argv[1], sys.argv[2]
steps = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
gem = [l.split(None, 1)[1].strip() for l in open(log, errors="ignore") if l.startswith("GEMMLOG")]
wg = [l.split(None, 1)[1].strip() for l in open(log, errors="ignore") if l.startswith("WGLOG")]
lines = [l for l in open(path) if not l.startswith("==")]
gi = wi = 0
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1000.0 if row["Metric Unit"] in ("ns", "nsecond") else v
    name = row["Kernel Name"]
    if ("k_conv_gemm_p" in name or "k_gemm2" in name) and gi < len(gem):
        key = ("gemm", gem[gi]); gi += 1
    elif "k_wgrad" in name and "reduce" not in name:
        if wi >= len(wg):
            continue
        key = ("wgrad", wg[wi]); wi += 1
    else:
        continue
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += v
def flops(kind, desc):
    d = dict(kv.split("=") for kv in desc.split() if "=" in kv)
    if kind == "gemm":
        return 2.0 * int(d["M"]) * int(d["N"]) * int(d["K"])
    return 2.0 * int(d["pixels"]) * int(d["Cout"]) * int(d["Kc"])
tot = sum(a[1] for a in agg.values())
print(f"paired {gi} gemm / {wi} wgrad launches; total {tot/steps:.0f} us/step")
for (kind, desc), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{kind:5s} {desc:62s} n={n/steps:5.1f}/step {t/steps:8.1f} us/step {flops(kind, desc)*n/t/1e6:7.0f} TFLOP/s avg {t/n:7.1f} us")
