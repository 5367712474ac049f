#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/bench_gemm.py cg2_n256 cg2_e16 > gpurun_out/bench_gemm.log 2>&1; tail -12 gpurun_out/bench_gemm.log
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py -q -x 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline']['share_of_step'], d['roofline']['timed_on'][:60], d['clocks'])"
