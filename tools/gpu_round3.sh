#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_probe_resnet.py > gpurun_out/probe_resnet.log 2>&1; echo "exit=$?" >> gpurun_out/probe_resnet.log
timeout 500 python -m pytest tests/test_resnet_gpu.py -q -m gpu -s > gpurun_out/pytest_resnet.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_resnet.log
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench exit=$?" >> gpurun_out/bench1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/smoke.log
cat gpurun_out/probe_resnet.log | tail -30; tail -15 gpurun_out/pytest_resnet.log; cat gpurun_out/bench1.json; tail -3 gpurun_out/bench1.err; tail -3 gpurun_out/smoke.log
