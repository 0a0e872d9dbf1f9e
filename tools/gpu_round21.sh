#!/bin/bash
# re-entry sanity: full GPU suite, smoke, bench, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/pytest21.log 2>&1; echo "exit=$?" >> gpurun_out/pytest21.log
tail -5 gpurun_out/pytest21.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke21.log 2>&1; tail -2 gpurun_out/smoke21.log
timeout 400 python bench.py --steps 100 --warmup 10 > gpurun_out/bench21.json 2> gpurun_out/bench21.err; cat gpurun_out/bench21.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches21.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b21.log 2>&1
python tools/agg_launches.py gpurun_out/launches21.csv 2>/dev/null | head -40
