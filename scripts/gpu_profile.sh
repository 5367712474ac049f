#!/bin/bash
# Round profile pass on one B200: bench line (with clocks), ncu launch list of one serial step, and ncu --set full captures of
# the GEMM family (4 launches of one audio layer), the audio attention kernel, LayerNorm and the frontend.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv 2>&1 &
SMI=$!
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
kill $SMI
cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm rc=$?"; cut -c1-300 gpurun_out/bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu --profile --serial-towers > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
python scripts/summarize_launches.py gpurun_out/launches.csv | tee gpurun_out/launch_summary.txt
P="python bench.py --steps 1 --warmup 3 --no-cpu --profile --serial-towers"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_f16_kernel -s 130 -c 4 -f -o gpurun_out/prof_gemm $P > gpurun_out/ncu_full.log 2>&1; echo "ncu gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc3 -s 14 -c 1 -f -o gpurun_out/prof_attn $P > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:layernorm_kernel -s 60 -c 1 -f -o gpurun_out/prof_ln $P > gpurun_out/ncu_ln.log 2>&1; echo "ncu ln rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:frontend_kernel -s 3 -c 1 -f -o gpurun_out/prof_fe $P > gpurun_out/ncu_fe.log 2>&1; echo "ncu fe rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_pool_kernel -s 6 -c 1 -f -o gpurun_out/prof_pool $P > gpurun_out/ncu_pool.log 2>&1; echo "ncu pool rc=$?"
ls -la gpurun_out/*.ncu-rep
