#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_eval_gpu.py -q -x > gpurun_out/pytest_fe.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_fe.log
timeout 300 python scripts/bench_frontend.py 2>&1 | tail -4
