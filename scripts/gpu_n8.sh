#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/check_sharded.py > gpurun_out/check_sharded_n$N.log 2>&1; echo "check_sharded rc=$?"; grep "sharded\|config-5" gpurun_out/check_sharded_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"; tail -1 gpurun_out/bench_n$N.json | cut -c1-220
