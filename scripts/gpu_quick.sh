#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_eval_gpu.py -x -q 2>&1 | tail -4
