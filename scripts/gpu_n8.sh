#!/bin/bash
# N = 1 and N = 8 back to back on the same box (scaling is only meaningful on one box), + config 5 on 8 GPUs + sharded checks
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_n1_samebox.json 2> gpurun_out/bench_n1_samebox.err; echo "bench n1 rc=$?"; tail -1 gpurun_out/bench_n1_samebox.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['clocks'])"
bash scripts/gpu_n2.sh 8
