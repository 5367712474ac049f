#!/bin/bash
mkdir -p gpurun_out
step() { name=$1; shift; echo "=== $name"; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n ${TAILN:-6} gpurun_out/$name.log; }
TAILN=20 step ops 300 python -m pytest tests/test_ops_gpu.py tests/test_attn_impls_gpu.py -q -x
TAILN=8 step model 400 python -m pytest tests/test_model_gpu.py -q -s
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu --profile --serial-towers > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
python scripts/summarize_launches.py gpurun_out/launches.csv | tee gpurun_out/launch_summary.txt
