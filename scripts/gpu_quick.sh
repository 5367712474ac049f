#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_eval_gpu.py tests/test_model_gpu.py -x -q > gpurun_out/pytest_fe.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_fe.log
timeout 100 python scripts/bench_frontend.py 2>&1 | tail -3
