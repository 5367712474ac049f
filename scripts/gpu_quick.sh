#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_eval_gpu.py -x -q > gpurun_out/pytest_eval.log 2>&1; echo "pytest eval rc=$?"; tail -40 gpurun_out/pytest_eval.log
