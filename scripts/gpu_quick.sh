#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/bench_attn.py 2>&1 | head -2
timeout 100 python scripts/trace_attn.py 5 > gpurun_out/tc5_trace.log 2>&1; echo "trace rc=$?"; head -34 gpurun_out/tc5_trace.log
