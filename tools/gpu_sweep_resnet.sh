#!/bin/bash
# A/B sweep of the ResNet-18 step knobs (one bench line each, steps/s printed)
mkdir -p gpurun_out
run() {
  tag=$1; shift
  v=$(env "$@" timeout 300 python bench.py --steps 100 --warmup 10 --no-ddpm --no-maskgen --no-modes --no-torch --no-cpu-baseline 2>gpurun_out/sweep_$tag.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['achieved'],1), round(d['roofline']['other']['achieved'],1))")
  echo "$tag: $v"
}
run base SALUN_X=0
run bwdfuse SALUN_BN_BWD_FUSE=1
run pair1 SALUN_RESNET_PAIR=1
run pair2 SALUN_RESNET_PAIR=2
run norw SALUN_CONV_RW=0
run norw_pair1 SALUN_CONV_RW=0 SALUN_RESNET_PAIR=1
run rw32 SALUN_CONV_RW=2
run rw32_pair1 SALUN_CONV_RW=2 SALUN_RESNET_PAIR=1
