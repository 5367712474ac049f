"""Time the full bench step and the audio-attention kernel with a given build of the library (A/B across two .so files,
interleaved process by process on one box).  Usage: python scripts/ab_lib.py path/to/libcaco_b200.so [label]"""
import json
import sys

import torch

sys.path.insert(0, ".")
from cacophony_b200 import _lib as L

L.LIB_PATH = sys.argv[1]
label = sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
import bench
import cacophony_b200 as cb
from cacophony_b200 import ops

torch.manual_seed(0)
model = cb.create_caco_model().cuda()
wave, ids, mask = [t.cuda() for t in bench.synth_inputs(256, 0)]


def timed(fn, n, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n, 4)


def step():
    a, t = model.encode_pairs(wave, ids, mask, max_patches=bench.MAX_PATCHES)
    return model.similarity(a, t)


qkv = torch.randn(256, 500, 3 * 768, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1)).half()
am = torch.ones(256, 500, device="cuda")
am[:, 496:] = 0
out = {"lib": label, "step_ms": [timed(step, 10) for _ in range(3)],
       "attn_ms": [timed(lambda: ops.attention_audio(qkv, am, 8), 20) for _ in range(2)],
       "attn_long_ms": timed(lambda: ops.attention_audio(qkv, am, 8), 2000)}
if len(sys.argv) > 3:                      # dump the attention output (bit-equality check between builds)
    torch.save(ops.attention_audio(qkv, am, 8).cpu(), sys.argv[3])
print(json.dumps(out), flush=True)
