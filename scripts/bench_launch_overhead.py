"""Fixed cost per launch of the tower kernels: tiny problems back to back on one stream (CUDA events)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from cacophony_b200 import _lib as L
from cacophony_b200 import ops


def timed(fn, iters=200, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for M, N, K in ((256, 256, 64), (256, 768, 768), (8192, 768, 768), (8192, 768, 3072), (8192, 2304, 768), (8192, 3072, 768)):
    a = torch.randn(M, K, device="cuda").half()
    w = torch.randn(N, K, device="cuda").half()
    b = torch.randn(N, device="cuda")
    o16 = torch.empty(M, N, device="cuda", dtype=torch.float16)
    x = torch.zeros(M, N, device="cuda")
    r = torch.zeros(M, N, device="cuda")
    for name, fn in (("bias_f16", lambda: ops.gemm_f16(a, w, b, L.EPI_BIAS_F16, out=o16)),
                     ("gelu_f16", lambda: ops.gemm_f16(a, w, b, L.EPI_BIAS_GELU_F16, out=o16)),
                     ("resid_inplace", lambda: ops.gemm_f16(a, w, b, L.EPI_BIAS_RESID_F32, resid=x, out=x)),
                     ("resid_separate", lambda: ops.gemm_f16(a, w, b, L.EPI_BIAS_RESID_F32, resid=r, out=x))):
        us = timed(fn)
        print(json.dumps({"kernel": "gemm", "M": M, "N": N, "K": K, "epi": name, "us": round(us, 2),
                          "tflops": round(2.0 * M * N * K / us / 1e6, 1)}), flush=True)
x = torch.randn(64, 768, device="cuda")
g = torch.ones(768, device="cuda")
print(json.dumps({"kernel": "layernorm 64 rows", "us": round(timed(lambda: ops.layernorm(x, g, g, want_f32=False, want_f16=True)), 2)}))
x = torch.randn(8192, 768, device="cuda")
print(json.dumps({"kernel": "layernorm 8192 rows", "us": round(timed(lambda: ops.layernorm(x, g, g, want_f32=False, want_f16=True)), 2)}))
