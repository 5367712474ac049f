#!/bin/bash
# round 2 (second session): full GPU test suite incl. the KV-cached decode, decode throughput by arm, memcheck of the cached decode
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -3; grep -E "kv-cached|Error|assert" gpurun_out/pytest_gpu.log | head -20
timeout 400 python scripts/bench_decode.py > gpurun_out/bench_decode.jsonl 2> gpurun_out/bench_decode.err; echo "decode bench rc=$?"; cat gpurun_out/bench_decode.jsonl; tail -5 gpurun_out/bench_decode.err
timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_model_gpu.py -q -x -k "kv_cached" > gpurun_out/sanitizer_kv.log 2>&1; echo "sanitizer rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_kv.log
