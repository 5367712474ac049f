#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/bench_attn.py > gpurun_out/attn_bench.log 2>&1; echo "attn rc=$?"; cat gpurun_out/attn_bench.log
timeout 300 python -m pytest tests/test_attn_impls_gpu.py -x -q > gpurun_out/pytest_attn.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_attn.log
timeout 300 python scripts/trace_attn.py > gpurun_out/pp_trace.log 2>&1; echo "trace rc=$?"; cat gpurun_out/pp_trace.log
