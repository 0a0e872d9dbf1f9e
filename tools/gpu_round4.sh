#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_all.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_all.log
timeout 500 python -m pytest tests/test_resnet_gpu.py -q -m gpu -s > gpurun_out/pytest_resnet.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_resnet.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/smoke.log
timeout 300 python bench.py --steps 100 --warmup 10 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench exit=$?" >> gpurun_out/bench2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench2_ref.json 2>> gpurun_out/bench2.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 450 --csv --log-file gpurun_out/launches2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -5 gpurun_out/pytest_all.log; grep -E "worst|Jaccard|passed|failed" gpurun_out/pytest_resnet.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench2.json gpurun_out/bench2_ref.json; tail -3 gpurun_out/bench2.err
