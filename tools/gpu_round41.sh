#!/bin/bash
# final evidence: full GPU suite, bench.py line, ncu launch list of the same command, ncu --set full of the top kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest41.log 2>&1; echo "exit=$?" >> gpurun_out/pytest41.log
tail -4 gpurun_out/pytest41.log
timeout 600 python bench.py > gpurun_out/bench41.json 2> gpurun_out/bench41.err; cut -c1-400 gpurun_out/bench41.json; tail -2 gpurun_out/bench41.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches41.csv python bench.py --steps 2 --warmup 3 --ddpm-steps 2 --no-cpu-baseline > gpurun_out/b41.log 2>&1
python tools/agg_launches.py gpurun_out/launches41.csv 2>/dev/null | head -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm2 --launch-skip 40 -c 1 -o gpurun_out/ncu_gemm2 -f python tools/bench_ddpm_step.py 1 --profile > gpurun_out/ncu_gemm2.log 2>&1; tail -2 gpurun_out/ncu_gemm2.log
ncu -i gpurun_out/ncu_gemm2.ncu-rep --page raw --csv > gpurun_out/ncu_gemm2_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_gemm2.ncu-rep --page details 2>/dev/null | grep -E "Duration|DRAM Throughput|L2 Cache Throughput|Compute \(SM\) Throughput|Grid Size|Registers|Shared Memory Config|L2 Hit Rate" | head -12
