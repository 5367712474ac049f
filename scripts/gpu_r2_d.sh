#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -3; grep -E "decoder logits|FAILED|Error" gpurun_out/pytest_gpu.log | head -20
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python scripts/bench_frontend.py 2>&1 | tail -2
