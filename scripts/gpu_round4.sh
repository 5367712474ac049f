#!/bin/bash
mkdir -p gpurun_out
step() { name=$1; shift; echo "=== $name"; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n ${TAILN:-6} gpurun_out/$name.log; }
step gemm_all 400 python -m pytest tests/test_gemm_gpu.py -q -x
TAILN=10 step bench_gemm 300 python scripts/bench_gemm.py cg2_n256
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 3 -c 1 -o gpurun_out/prof_attn python scripts/bench_attn.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
