#!/bin/bash
mkdir -p gpurun_out
step() { name=$1; shift; echo "=== $name"; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=20 step ops_attn 240 python -m pytest tests/test_ops_gpu.py -q -x -k "attention"
TAILN=4 step attn_bench 240 python scripts/bench_attn.py
TAILN=8 step model 600 python -m pytest tests/test_model_gpu.py -q -s
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --serial-towers > gpurun_out/bench_serial.json 2> gpurun_out/bench.err; echo "bench serial rc=$?"; cut -c1-330 gpurun_out/bench_serial.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
