#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest48.log 2>&1; echo "exit=$?" >> gpurun_out/pytest44.log
tail -4 gpurun_out/pytest48.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke48.log 2>&1; tail -2 gpurun_out/smoke44.log | cut -c1-200
timeout 600 python bench.py > gpurun_out/bench48.json 2> gpurun_out/bench48.err; python -c "
import json; d=json.load(open('gpurun_out/bench48.json')); print('resnet', d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks']); dd=d['ddpm']; print('ddpm', dd['value'], dd['ms_per_it'], dd['e2e']['value'], dd['roofline']['frac'], dd['launches_per_it'])"; tail -2 gpurun_out/bench48.err
