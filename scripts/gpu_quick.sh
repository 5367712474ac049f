#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_eval_gpu.py tests/test_model_gpu.py -x -q > gpurun_out/pytest_fe.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_fe.log
timeout 100 python scripts/bench_frontend.py 2>&1 | tail -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:frontend_kernel -s 4 -c 1 -f -o gpurun_out/prof_fe python scripts/bench_frontend.py > gpurun_out/ncu_fe.log 2>&1; echo "ncu fe rc=$?"
