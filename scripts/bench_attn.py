"""Micro-benchmark of the audio attention kernels (tcgen05 vs warp-level mma.sync) at the bench shape."""
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
for impl, name in ((4, "tcgen05_pingpong"), (3, "tcgen05_persistent"), (2, "tcgen05"), (1, "mma_sync")):
    lib.caco_set_attention_impl(impl)
    for _ in range(3):
        o = ops.attention_audio(qkv, mask, H)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        o = ops.attention_audio(qkv, mask, H)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 4.0 * B * H * S * S * dh
    res[name] = o.float()
    print(json.dumps({"impl": name, "ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1)}), flush=True)
d = (res["tcgen05"] - res["mma_sync"]).abs().max().item()
d2 = (res["tcgen05_persistent"] - res["mma_sync"]).abs().max().item()
d3 = (res["tcgen05_pingpong"] - res["mma_sync"]).abs().max().item()
print(json.dumps({"max_abs_diff_between_impls": d, "persistent_vs_mma_sync": d2, "pingpong_vs_mma_sync": d3}))
lib.caco_set_attention_impl(0)
