#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/probe_conv.py > gpurun_out/probe_conv.log 2>&1; cat gpurun_out/probe_conv.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_gemm_p --launch-skip 2 -c 1 -o gpurun_out/ncu_conv128 -f python tools/probe_conv.py 1 1 > gpurun_out/ncu_conv128.log 2>&1; tail -3 gpurun_out/ncu_conv128.log
ncu -i gpurun_out/ncu_conv128.ncu-rep --page raw --csv > gpurun_out/ncu_conv128_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_conv128.ncu-rep --page details 2>/dev/null | grep -E "Duration|Throughput|Pipe|Hit Rate|Stall|Issue|Warp Cycles|L2|DRAM|Tensor|No Eligible|Active Warps" | head -60
