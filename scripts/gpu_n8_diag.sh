#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2951$i bench.py --gpus 8 --steps 10 --warmup 3 --diag-local --no-cpu 2> gpurun_out/diag_n8.err | tail -1 | tee -a gpurun_out/diag_n8.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2952$i bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu 2>> gpurun_out/diag_n8.err | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('value','ms_per_step','n_gpus','tail')}))" | tee -a gpurun_out/diag_n8.jsonl
done
nvidia-smi --query-gpu=index,power.limit,clocks.max.sm,temperature.gpu --format=csv | tee -a gpurun_out/diag_n8.jsonl
