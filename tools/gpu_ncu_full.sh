#!/bin/bash
mkdir -p gpurun_out
for spec in "k_conv_rw:rw:2" "k_wgrad:wgrad:24" "k_conv_gemm:gemm:12"; do
  IFS=: read pat tag skip <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c 2 -f -o gpurun_out/prof_$tag python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$tag.log 2>&1
  echo "$tag exit=$?"
done
ls -la gpurun_out/*.ncu-rep
