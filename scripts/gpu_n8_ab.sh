#!/bin/bash
mkdir -p gpurun_out
run() { # name, nproc, extra args
  if [ "$2" == "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu $3 > gpurun_out/$1.json 2> gpurun_out/$1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $2 --steps 10 --warmup 3 --no-cpu $3 > gpurun_out/$1.json 2> gpurun_out/$1.err
  fi
  tail -1 gpurun_out/$1.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', {k:d.get(k) for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], 'tail', d['tail']['ms'], d['tail'].get('exchange'), d['tail'].get('pipelined'), d['clocks']['sm_mhz'])"
}
run ab_n1_a 1 ""
run ab_n8_pipe_a 8 ""
run ab_n8_nopipe_a 8 "--no-pipeline"
run ab_n8_pipe_b 8 ""
run ab_n8_nopipe_b 8 "--no-pipeline"
run ab_n1_b 1 ""
run ab_n4_pipe 4 ""
run ab_n2_pipe 2 ""
