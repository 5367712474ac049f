#!/bin/bash
# quick GPU round: GPU tests, GEMM micro-bench, one bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -i "rel\|err" gpurun_out/pytest_gpu.log | head -20; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python scripts/bench_gemm.py cg2_n256 > gpurun_out/bench_gemm.log 2>&1; echo "gemm rc=$?"; cat gpurun_out/bench_gemm.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
