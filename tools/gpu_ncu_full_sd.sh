#!/bin/bash
# ncu --set full captures of the SD U-Net forward's kernels (one launch each, eager replay of tools/sd_launches.py), raw pages as CSV:
#   k_flash_attn (a 4096-token self-attention, both builds), k_conv_gemm_p<160,5> (N = 320 convolution at 64x64), k_gn2_partial
mkdir -p gpurun_out
for spec in "k_flash_attn:sd_flash_bf16:bf16:0" "k_flash_attn:sd_flash_split:split:0" "k_conv_gemm_p<160:sd_conv160:bf16:3" "k_gn2_partial:sd_gn:bf16:1"; do
  IFS=: read pat tag prec skip <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$pat" -s $skip -c 1 -f -o gpurun_out/r2_prof_$tag python tools/sd_launches.py $prec > gpurun_out/ncu_$tag.log 2>&1
  echo "$tag exit=$?"
  ncu -i gpurun_out/r2_prof_$tag.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_$tag.csv 2>/dev/null
done
ls -la gpurun_out/r2_ncu_full_sd_*.csv
