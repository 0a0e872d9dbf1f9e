#!/bin/bash
# ncu --set full captures of the round's dominant kernels (one launch each, single GPU), raw pages exported as CSV.
#   ResNet-18 step : k_conv_rw<32,1> (layer1 3x3), k_conv_gemm_p<128,6> (layer3/4), k_wgrad
#   DDPM iteration : k_gemm2 (the 128->128 3x3 convolution at 32x32, batch 256)
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-maskgen --no-modes --no-torch --no-cpu-baseline"
for spec in "k_conv_rw:rw:2:--no-ddpm" "k_wgrad:wgrad:24:--no-ddpm" "k_conv_gemm_p:gemm:12:--no-ddpm" "k_gemm2:unet_conv:40:"; do
  IFS=: read pat tag skip extra <<< "$spec"
  SALUN_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c 1 -f -o gpurun_out/r2_prof_$tag $B $extra > gpurun_out/ncu_$tag.log 2>&1
  echo "$tag exit=$?"
  ncu -i gpurun_out/r2_prof_$tag.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_$tag.csv 2>/dev/null
done
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_ncu_full_*.csv
