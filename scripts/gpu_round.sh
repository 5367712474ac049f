#!/bin/bash
# One gpurun call = many isolated experiments: each step runs in its own process under its own timeout, so a hang or
# a sticky CUDA error in one kernel variant cannot take the others (or the box) down.  Logs land in gpurun_out/.
mkdir -p gpurun_out
step() { name=$1; shift; echo "=== $name"; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n ${TAILN:-6} gpurun_out/$name.log; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
step gemm_cg1_n256 300 python -m pytest tests/test_gemm_gpu.py -q -x -k "cg1_n256 or rejects"
step gemm_cg1_n128 300 python -m pytest tests/test_gemm_gpu.py -q -x -k "cg1_n128"
step gemm_cg2_n256 300 python -m pytest tests/test_gemm_gpu.py -q -x -k "cg2_n256"
step ops 600 python -m pytest tests/test_ops_gpu.py -q
TAILN=25 step model 900 python -m pytest tests/test_model_gpu.py -q -s
TAILN=40 step bench_gemm 600 python scripts/bench_gemm.py "$@"
