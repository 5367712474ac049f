#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -x -q -k "full_bench or graph" 2>&1 | tail -6
