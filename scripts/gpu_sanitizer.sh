#!/bin/bash
# compute-sanitizer memcheck over the whole path (smoke), the captioning head, the ragged / TMA frontend and the attention tests
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python __graft_entry__.py smoke > gpurun_out/sanitizer_smoke.log 2>&1; echo "smoke rc=$?"; grep -E "ERROR SUMMARY|smoke ok" gpurun_out/sanitizer_smoke.log
timeout 1200 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_attn_impls_gpu.py tests/test_ops_gpu.py -q -x -k "attention or frontend or pool or split" > gpurun_out/sanitizer_ops.log 2>&1; echo "ops rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_ops.log
timeout 1200 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_model_gpu.py -q -x -k "decoder" > gpurun_out/sanitizer_decoder.log 2>&1; echo "decoder rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_decoder.log
