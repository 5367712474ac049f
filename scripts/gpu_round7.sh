#!/bin/bash
mkdir -p gpurun_out
step() { name=$1; shift; echo "=== $name"; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=25 step ops_attn 180 python -m pytest tests/test_ops_gpu.py -q -x -k "attention_audio"
TAILN=5 step attn_bench 180 python scripts/bench_attn.py
TAILN=8 step model 400 python -m pytest tests/test_model_gpu.py -q -s
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
