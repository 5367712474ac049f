"""Micro-benchmark of the audio attention kernel at the bench shape for every exp2 split (attn_poly 0..3)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from cacophony_b200 import _lib as L
from cacophony_b200 import ops

B, S, H, dh = 256, 500, 8, 96
qkv = (torch.randn(B, S, 3 * H * dh, device="cuda") * 1.0).half()
mask = torch.ones(B, S, device="cuda")
mask[:, 496:] = 0
lib = L.load()
res = {}
for rnd in range(2):
    for poly, name in ((0, "mufu_only"), (1, "poly_1_of_4"), (3, "poly_3_of_8"), (2, "poly_1_of_2")):
        lib.caco_set_default_option(b"attn_poly", poly)
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
        res[name] = o.float()
        print(json.dumps({"round": rnd, "impl": name, "ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1)}), flush=True)
ref = res["mufu_only"]
print(json.dumps({k: float((v - ref).abs().max()) for k, v in res.items()}))
lib.caco_set_default_option(b"attn_poly", 0)
