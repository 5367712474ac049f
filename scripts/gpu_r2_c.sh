#!/bin/bash
# round 2: full GPU test suite, smoke, headline bench (with the reference CPU arm + torch library baseline), config-5 bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -3; grep -E "logits row-relative|model_s|split_weights" gpurun_out/pytest_gpu.log | head -12
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv 2>&1 &
SMI=$!
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
kill $SMI
timeout 600 python bench.py --workload zeroshot --steps 10 --warmup 3 > gpurun_out/bench_zeroshot.json 2> gpurun_out/bench_zeroshot.err; echo "zeroshot rc=$?"; cat gpurun_out/bench_zeroshot.json; tail -3 gpurun_out/bench_zeroshot.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm rc=$?"; cut -c1-400 gpurun_out/bench_reference.json; tail -2 gpurun_out/bench_reference.err
timeout 300 python scripts/bench_decode.py > gpurun_out/bench_decode.jsonl 2>&1; cat gpurun_out/bench_decode.jsonl
timeout 300 python scripts/bench_attn.py 2>&1 | tail -1
timeout 300 python scripts/bench_frontend.py 2>&1 | tail -1
