#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/probe_shortk.py > gpurun_out/probe_shortk.log 2>&1; cat gpurun_out/probe_shortk.log
