#!/bin/bash
mkdir -p gpurun_out
step() { name=$1; shift; echo "=== $name"; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n ${TAILN:-6} gpurun_out/$name.log; }
step gemm_e16 300 python -m pytest tests/test_gemm_gpu.py -q -x -k "cg2_e16"
TAILN=30 step bench_gemm 300 python scripts/bench_gemm.py cg2_e16 cg2_n256
