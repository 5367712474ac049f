#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,power.limit,temperature.gpu,clocks.max.sm --format=csv
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['achieved'], d['roofline_hbm']['achieved'])"
timeout 600 python scripts/ab_step.py pdl=caco_set_pdl:1 nopdl=caco_set_pdl:0 2>&1 | tail -1
