#!/bin/bash
# A/B of two library builds on one box, interleaved process by process.  The "old" build is made beforehand from an earlier
# revision of the kernel under test, e.g.:  git show <rev>:cacophony_b200/csrc/attention_pp.cu > scratch/old/attention_pp.cu;
# nvcc (flags of cacophony_b200/build.py) -c it; nvcc -shared -o scratch/libcaco_b200_old.so <the other objects of _build/> + it
mkdir -p gpurun_out
for r in 1 2; do
  timeout 200 python scripts/ab_lib.py scratch/libcaco_b200_old.so old 2>&1 | tail -1 | tee -a gpurun_out/ab_lib.jsonl
  timeout 200 python scripts/ab_lib.py cacophony_b200/libcaco_b200.so new 2>&1 | tail -1 | tee -a gpurun_out/ab_lib.jsonl
done
