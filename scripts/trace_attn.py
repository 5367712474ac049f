"""Timeline of CTA 0 of the persistent ping-pong attention kernel (clock64 stamps, see PP_STAMP in attention_pp.cu)."""
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
for _ in range(2):
    ops.attention_audio(qkv, mask, H)
buf = torch.zeros(2 * 64 * 8, dtype=torch.int64, device="cuda")
trace_fn(buf.data_ptr())
ops.attention_audio(qkv, mask, H)
torch.cuda.synchronize()
trace_fn(None)
t = buf.cpu().view(2, 64, 8)
t0 = int(t[t > 0].min())
names = [["iter", "waited", "A_issued", "pB_seen", "B_issued", "-", "-", "-"],
         ["wait_s", "s_ready", "s_loaded", "max_done", "p_written", "pv_done", "stored", "-"]]
for role, rn in ((0, "MMA"), (1, "SMX")):
    print(rn, " ".join(f"{n:>10s}" for n in names[role]))
    for g in range(14):
        row = t[role, g]
        print(f"g={g:2d}", " ".join(f"{(int(v) - t0) if v > 0 else -1:10d}" for v in row))
