#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/pytest30.log 2>&1; echo "exit=$?" >> gpurun_out/pytest30.log
tail -4 gpurun_out/pytest30.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench30.json 2> gpurun_out/bench30.err; cat gpurun_out/bench30.json; tail -3 gpurun_out/bench30.err
