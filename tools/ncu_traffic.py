"""profiles/r2_ncu_full_*.csv (ncu --set full raw pages, one launch each) -> profiles/r2_traffic.json (what bench.py quotes
as roofline.traffic) and a one-screen summary of the metrics the roofline discussion uses.
    python tools/ncu_traffic.py > profiles/r2_ncu_summary.txt"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "lts__t_sectors.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic"]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}
out = {}
for tag, key in (("rw", "resnet_conv"), ("gemm", "resnet_gemm"), ("wgrad", "resnet_wgrad"), ("unet_conv", "unet_conv")):
    path = os.path.join(ROOT, "profiles", f"r2_ncu_full_{tag}.csv")
    if not os.path.exists(path):
        continue
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    name = vals[col["Kernel Name"]] if "Kernel Name" in col else "?"
    print(f"== {tag}: {name[:120]}")
    rec = {}
    for m in WANT:
        if m in col:
            v, u = vals[col[m]].replace(",", ""), units[col[m]]
            rec[m] = (float(v) if v not in ("", "n/a") else None, u)
            print(f"   {m:70s} {v:>16s} {u}")
    rd, wr = rec.get("dram__bytes_read.sum"), rec.get("dram__bytes_write.sum")
    if rd and wr and rd[0] is not None:
        b = rd[0] * UNIT.get(rd[1], 1.0) + wr[0] * UNIT.get(wr[1], 1.0)
        out[key] = {"bytes": int(b), "kernel": name[:160],
                    "source": f"dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full (profiles/r2_ncu_full_{tag}.csv)"}
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w"), indent=1)
