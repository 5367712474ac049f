#!/bin/bash
# round 2, call A: GPU tests after the options refactor, attention exp2-split bench, L2-blocking chunk sweep
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python scripts/bench_attn.py > gpurun_out/attn_bench.log 2>&1; echo "attn rc=$?"; cat gpurun_out/attn_bench.log
timeout 900 python scripts/sweep_chunks.py > gpurun_out/sweep_chunks.log 2>&1; echo "sweep rc=$?"; tail -4 gpurun_out/sweep_chunks.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
