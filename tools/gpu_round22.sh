#!/bin/bash
# U-Net engine bring-up: probes vs torch on three configs
mkdir -p gpurun_out
for cfg in "tiny 6" "small 8" "full 8"; do
  set -- $cfg
  timeout 600 python tools/gpu_unet_probe.py $1 $2 > gpurun_out/unet_probe_$1.log 2>&1; echo "exit=$?" >> gpurun_out/unet_probe_$1.log
  grep -E "==|<--|PROBE|whole|engine fwd|Error|error|exit=" gpurun_out/unet_probe_$1.log | head -40
done
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/gpu_unet_probe.py tiny 6 > gpurun_out/unet_memcheck.log 2>&1
grep -E "ERROR SUMMARY|Invalid|at salun|PROBE" gpurun_out/unet_memcheck.log | head -20
