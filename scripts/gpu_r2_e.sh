#!/bin/bash
# round 2, final-code pass on one B200: full GPU suite, smoke, headline + config-5 benches, decode / attention / frontend
# micro-benches, ncu launch list of one serial step, ncu --set full of the attention kernel and the frontend
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv 2>&1 &
SMI=$!
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
kill $SMI
timeout 600 python bench.py --workload zeroshot --steps 10 --warmup 3 > gpurun_out/bench_zeroshot.json 2> gpurun_out/bench_zeroshot.err; echo "zeroshot rc=$?"; cut -c1-600 gpurun_out/bench_zeroshot.json; tail -3 gpurun_out/bench_zeroshot.err
timeout 300 python scripts/bench_decode.py > gpurun_out/bench_decode.jsonl 2>&1; cat gpurun_out/bench_decode.jsonl
timeout 100 python scripts/bench_attn.py 2>&1 | tee gpurun_out/bench_attn.log | tail -1
timeout 100 python scripts/trace_attn.py 1500 > gpurun_out/trace_attn.log 2>&1; head -1 gpurun_out/trace_attn.log
timeout 100 python scripts/bench_frontend.py 2>&1 | tail -1
P="python bench.py --steps 1 --warmup 3 --no-cpu --profile --serial-towers"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $P > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
python scripts/summarize_launches.py gpurun_out/launches.csv | tee gpurun_out/launch_summary.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attention_pp -s 14 -c 1 -f -o gpurun_out/prof_attn $P > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:frontend_kernel -s 3 -c 1 -f -o gpurun_out/prof_fe $P > gpurun_out/ncu_fe.log 2>&1; echo "ncu fe rc=$?"
python scripts/ncu_summary.py gpurun_out/ncu_final_summary.csv gpurun_out/prof_attn.ncu-rep gpurun_out/prof_fe.ncu-rep 2>&1 | tail -2
