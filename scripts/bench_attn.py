"""Micro-benchmark of the audio attention kernel at the bench shape (256 clips x 8 heads x 500 tokens, head_dim 96)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from cacophony_b200 import ops

B, S, H, dh = 256, 500, 8, 96
qkv = (torch.randn(B, S, 3 * H * dh, device="cuda") * 1.0).half()
mask = torch.ones(B, S, device="cuda")
mask[:, 496:] = 0
for rnd in range(3):
    for _ in range(3):
        o = ops.attention_audio(qkv, mask, H)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        o = ops.attention_audio(qkv, mask, H)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    fl = 4.0 * B * H * S * S * dh
    print(json.dumps({"round": rnd, "kernel": "attention_pp_kernel", "ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1)}), flush=True)
