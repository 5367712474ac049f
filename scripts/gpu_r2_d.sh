#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -3; grep -E "logits row-relative|^model_s|split_weights|FAILED" gpurun_out/pytest_gpu.log | head -20
timeout 600 python bench.py --workload zeroshot --steps 10 --warmup 3 > gpurun_out/bench_zeroshot.json 2> gpurun_out/bench_zeroshot.err; echo "zeroshot rc=$?"; cut -c1-200 gpurun_out/bench_zeroshot.json; python -c "
import json; d=json.load(open('gpurun_out/bench_zeroshot.json')); print(d['value'], d['trim_padding'], d['e2e'], d['oracle_agreement'])"
