#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest14.log 2>&1; echo "exit=$?" >> gpurun_out/pytest14.log
timeout 600 python tools/bench_ddpm.py 10 > gpurun_out/bench_ddpm.json 2> gpurun_out/bench_ddpm.err
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench14.json 2> gpurun_out/bench14.err
tail -6 gpurun_out/pytest14.log; cat gpurun_out/bench_ddpm.json; tail -3 gpurun_out/bench_ddpm.err
python -c "
import json
d=json.load(open('gpurun_out/bench14.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
