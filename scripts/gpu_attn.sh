#!/bin/bash
# attention kernel iteration: tests of the kernel, micro-benchmark, CTA-0 timeline
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_attn_impls_gpu.py -x -q > gpurun_out/pytest_attn.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_attn.log
timeout 120 python scripts/bench_attn.py 2>&1 | tee gpurun_out/bench_attn.log | tail -3
timeout 120 python scripts/trace_attn.py > gpurun_out/trace_attn.log 2>&1; tail -34 gpurun_out/trace_attn.log
