"""Micro-benchmark of the frontend kernel (K1) at the bench shape: 256 clips x 160000 samples -> [256, 500, 256] fp16 patches."""
import json
import sys

import torch

sys.path.insert(0, ".")
from cacophony_b200 import ops

B, n, P = 256, 160000, 500
w = (0.1 * (2 * torch.rand(B, n, device="cuda") - 1)).float()
for want_f16 in (False, True):
    for _ in range(3):
        ops.frontend(w, P, want_f16=want_f16)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.frontend(w, P, want_f16=want_f16)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    algo = B * (n * 4 + P * 256 * 4 + 3 * P * 4) + (B * P * 256 * 2 if want_f16 else 0)
    print(json.dumps({"kernel": "frontend", "also_f16": want_f16, "ms": round(ms, 4), "algorithmic_GB_per_s": round(algo / ms / 1e6, 1)}))
