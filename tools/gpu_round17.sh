#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 340 -c 340 --csv --log-file gpurun_out/launches17_warm.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench17.log 2>&1
python tools/agg_launches.py gpurun_out/launches17_warm.csv 2 | head -40
