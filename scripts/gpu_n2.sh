#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/check_sharded.py > gpurun_out/check_sharded_n$N.log 2>&1; echo "check_sharded rc=$?"; grep -v "^W\|^\[W" gpurun_out/check_sharded_n$N.log | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"; tail -1 gpurun_out/bench_n$N.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','n_gpus','tail','per_rank_ms_per_step')}, d['e2e']['value'], d['clocks'])"
tail -3 gpurun_out/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload zeroshot --steps 10 --warmup 3 > gpurun_out/bench_zeroshot_n$N.json 2> gpurun_out/bench_zeroshot_n$N.err; echo "zeroshot n$N rc=$?"; tail -1 gpurun_out/bench_zeroshot_n$N.json | cut -c1-900
timeout 300 python -m pytest tests/test_model_gpu.py -q -k second_device 2>&1 | tail -2
