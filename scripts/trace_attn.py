"""Timeline of CTA 0 of the persistent ping-pong attention kernel (clock64 stamps, see PP_STAMP in attention_pp.cu): the
MMA-issuing thread and the first softmax warp of tile A and of tile B, plus two wall-clock anchors (SM clock during the kernel).
Usage: python scripts/trace_attn.py [launches before the traced one]"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from cacophony_b200 import _lib as L
from cacophony_b200 import ops

B, S, H, dh = 256, 500, 8, 96
qkv = torch.randn(B, S, 3 * H * dh, device="cuda").half()
mask = torch.ones(B, S, device="cuda")
mask[:, 496:] = 0
lib = L.load()
trace_fn = lib.caco_attn_trace
trace_fn.argtypes = [C.c_void_p]
PREHEAT = int(sys.argv[1]) if len(sys.argv) > 1 else 2      # launches before the traced one (3000 = ~1 s of sustained load)
for _ in range(PREHEAT):
    ops.attention_audio(qkv, mask, H)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    ops.attention_audio(qkv, mask, H)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 200
buf = torch.zeros(3 * 64 * 8, dtype=torch.int64, device="cuda")
trace_fn(buf.data_ptr())
ops.attention_audio(qkv, mask, H)
torch.cuda.synchronize()
trace_fn(None)
t = buf.cpu().view(3, 64, 8)
ns = int(t[0, 63, 5] - t[0, 0, 5])
cyc = int(t[0, 63, 0] - t[0, 0, 0])
t[0, :, 5] = 0
print(f"{ms:.4f} ms per launch (200 launches after {PREHEAT}); blocks 0..63 of CTA 0: {cyc} cycles in {ns} ns -> "
      f"SM clock {1e3 * cyc / max(ns, 1):.0f} MHz, {cyc / 63:.0f} cycles per block")
t0 = int(t[t > 0].min())
names = [["iter", "waited", "A_issued", "pB_seen", "B_issued", "-", "-", "-"],
         ["wait_s", "s_ready", "s_loaded", "max_done", "p_written", "pv_done", "stored", "-"]]
for role, rn in ((0, "MMA"), (1, "SMX_A"), (2, "SMX_B")):
    print(rn, " ".join(f"{n:>10s}" for n in names[min(role, 1)]))
    for g in range(4, 13):
        row = t[role, g]
        print(f"g={g:2d}", " ".join(f"{(int(v) - t0) if v > 0 else -1:10d}" for v in row))
