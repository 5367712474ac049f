#!/bin/bash
# attention: timeline with wall-clock anchors (cold and right after a sustained load), old-vs-new output equality
mkdir -p gpurun_out
timeout 120 python scripts/trace_attn.py 2>&1 | head -3
timeout 200 python scripts/trace_attn.py 3000 2>&1 | head -3
timeout 100 python scripts/ab_lib.py scratch/libcaco_b200_old.so old /tmp/o_old.pt 2>&1 | tail -1
timeout 100 python scripts/ab_lib.py cacophony_b200/libcaco_b200.so new /tmp/o_new.pt 2>&1 | tail -1
python -c "
import torch
a, b = torch.load('/tmp/o_old.pt'), torch.load('/tmp/o_new.pt')
print('attention outputs old vs new kernel bit-equal:', torch.equal(a, b), float((a.float()-b.float()).abs().max()))"
