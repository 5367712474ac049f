#!/bin/bash
# bench + ncu launch list (+ optional full capture of the GEMM) on one B200.  Usage: gpu_bench.sh [full]
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/clocks.csv 2>&1 &
SMI=$!
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
kill $SMI
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu --profile > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
python scripts/summarize_launches.py gpurun_out/launches.csv | tee gpurun_out/launch_summary.txt
if [ "$1" == "full" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_f16_kernel -s 130 -c 4 -o gpurun_out/prof_gemm \
     python bench.py --steps 1 --warmup 3 --no-cpu --profile > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
