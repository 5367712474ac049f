#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/bench_attn.py > gpurun_out/attn_bench.log 2>&1; echo "attn rc=$?"; cat gpurun_out/attn_bench.log
timeout 600 python -m pytest tests/test_attn_impls_gpu.py tests/test_model_gpu.py -x -q > gpurun_out/pytest_attn.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_attn.log
