#!/bin/bash
# final-code 2-GPU pass: sharded logits / predictions / rankings against the single-GPU results, headline bench at N = 2
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_sharded.py > gpurun_out/check_sharded_n2.log 2>&1; echo "check_sharded rc=$?"; grep -v "^W\|^\[W" gpurun_out/check_sharded_n2.log | tail -6
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; tail -1 gpurun_out/bench_n2.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','n_gpus','tail','per_rank_ms_per_step')}, d['e2e']['value'], d['clocks'])"
tail -2 gpurun_out/bench_n2.err
