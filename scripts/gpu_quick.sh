#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/sanitizer_smoke.log 2>&1; echo "memcheck smoke rc=$?"; tail -6 gpurun_out/sanitizer_smoke.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_eval_gpu.py -x -q -k "topk or retrieval_metric or avg_pool or ragged_frontend or resample" > gpurun_out/sanitizer_eval.log 2>&1; echo "memcheck eval rc=$?"; tail -6 gpurun_out/sanitizer_eval.log
