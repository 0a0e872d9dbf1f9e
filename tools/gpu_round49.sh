#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench49.json 2> gpurun_out/bench49.err; python -c "
import json; d=json.load(open('gpurun_out/bench49.json')); print('resnet', d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic']); dd=d['ddpm']; print('ddpm', dd['value'], dd['ms_per_it'], dd['e2e']['value'], dd['roofline']['frac'], dd['roofline']['achieved'], dd['roofline']['other'])"; tail -2 gpurun_out/bench49.err
