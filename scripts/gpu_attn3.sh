#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_attn_impls_gpu.py -x -q 2>&1 | tail -1
for gap in 0 300 600 900; do
  timeout 120 python scripts/trace_attn.py 1500 $gap > gpurun_out/trace_gap$gap.log 2>&1; head -1 gpurun_out/trace_gap$gap.log
done
sed -n 2,32p gpurun_out/trace_gap0.log
