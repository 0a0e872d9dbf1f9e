#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/probe_two_streams.py > gpurun_out/probe_two_streams.log 2>&1; cat gpurun_out/probe_two_streams.log
