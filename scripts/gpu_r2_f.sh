#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -x -k "resid_ln or split or resid_in_place" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_model_gpu.py -q -x 2>&1 | tail -3
timeout 600 python scripts/ab_step.py fused=fuse_ln:1 unfused=fuse_ln:0 2>&1 | tail -2
timeout 600 python scripts/bench_text.py 2>&1 | tail -1
