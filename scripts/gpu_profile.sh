#!/bin/bash
# Round profile pass on one B200: ncu launch list of one serial step and ncu --set full captures of the GEMM family (4 launches
# of one audio layer), the audio attention kernel, LayerNorm, the frontend and the pooler.  (Bench lines come from gpu_r2_c.sh.)
mkdir -p gpurun_out
P="python bench.py --steps 1 --warmup 3 --no-cpu --profile --serial-towers"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $P > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
python scripts/summarize_launches.py gpurun_out/launches.csv | tee gpurun_out/launch_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_f16_kernel -s 130 -c 4 -f -o gpurun_out/prof_gemm $P > gpurun_out/ncu_full.log 2>&1; echo "ncu gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_pp -s 14 -c 1 -f -o gpurun_out/prof_attn $P > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:layernorm_kernel -s 60 -c 1 -f -o gpurun_out/prof_ln $P > gpurun_out/ncu_ln.log 2>&1; echo "ncu ln rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:frontend_kernel -s 3 -c 1 -f -o gpurun_out/prof_fe $P > gpurun_out/ncu_fe.log 2>&1; echo "ncu fe rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_pool_kernel -s 6 -c 1 -f -o gpurun_out/prof_pool $P > gpurun_out/ncu_pool.log 2>&1; echo "ncu pool rc=$?"
ls -la gpurun_out/*.ncu-rep
