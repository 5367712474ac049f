#!/bin/bash
mkdir -p gpurun_out
timeout 100 python scripts/trace_attn.py 6 > gpurun_out/tc6_trace.log 2>&1; echo "trace rc=$?"; head -22 gpurun_out/tc6_trace.log; grep -A20 "^SMX" gpurun_out/tc6_trace.log | head -21
