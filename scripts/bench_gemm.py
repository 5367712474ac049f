"""Micro-benchmark of the tcgen05 GEMM variants at the towers' real shapes (CUDA events, L2 flushed by size).
Usage: python scripts/bench_gemm.py [variant ...]   -> one line per (variant, shape): ms, TFLOP/s."""
import json
import math
import sys
import time

import torch

sys.path.insert(0, ".")
from cacophony_b200 import _lib as L
from cacophony_b200 import ops

SHAPES = [  # (name, M, N, K, epi)
    ("qkv", 128000, 2304, 768, L.EPI_BIAS_F16),
    ("out", 128000, 768, 768, L.EPI_BIAS_RESID_F32),
    ("fc1", 128000, 3072, 768, L.EPI_BIAS_SILU_F16),
    ("fc2", 128000, 768, 3072, L.EPI_BIAS_RESID_F32),
    ("in", 128000, 768, 256, L.EPI_BIAS_F32),
    ("t_qkv", 8192, 2304, 768, L.EPI_BIAS_F16),
    ("t_out", 8192, 768, 768, L.EPI_BIAS_RESID_F32),
    ("t_fc1", 8192, 3072, 768, L.EPI_BIAS_GELU_F16),
    ("t_fc2", 8192, 768, 3072, L.EPI_BIAS_RESID_F32),
]
VARIANTS = {"cg1_n256": L.GEMM_CG1_N256, "cg1_n128": L.GEMM_CG1_N128, "cg2_n256": L.GEMM_CG2_N256,
            "cg2_e16": L.GEMM_CG2_N256_E16}


def main():
    names = [a for a in sys.argv[1:] if a in VARIANTS] or list(VARIANTS)
    if "nored" in sys.argv[1:]:
        L.load().caco_set_default_option(b"resid_red", 0)
    for vn in names:
        v = VARIANTS[vn]
        for name, M, N, K, epi in SHAPES:
            a = torch.randn(M, K, device="cuda").half()
            w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).half()
            bias = torch.randn(N, device="cuda")
            f32 = epi in (L.EPI_BIAS_F32, L.EPI_BIAS_RESID_F32)
            out = torch.empty(M, N, device="cuda", dtype=torch.float32 if f32 else torch.float16)
            resid = out if epi == L.EPI_BIAS_RESID_F32 else None
            if resid is not None:
                out.zero_()
            for _ in range(3):
                ops.gemm_f16(a, w, bias, epi, resid=resid, variant=v, out=out)
            torch.cuda.synchronize()
            iters = 10
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                ops.gemm_f16(a, w, bias, epi, resid=resid, variant=v, out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            tf = 2.0 * M * N * K / ms / 1e9
            print(json.dumps({"variant": vn, "shape": name, "M": M, "N": N, "K": K, "ms": round(ms, 4), "tflops": round(tf, 1)}), flush=True)
            del a, w, out
    # cuBLAS reference point for the same shapes (library bar, fp16 in / fp16 out, no epilogue)
    for name, M, N, K, epi in SHAPES[:4]:
        a = torch.randn(M, K, device="cuda").half()
        w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).half()
        for _ in range(3):
            torch.matmul(a, w.t())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            torch.matmul(a, w.t())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(json.dumps({"variant": "cublas_f16", "shape": name, "ms": round(ms, 4), "tflops": round(2.0 * M * N * K / ms / 1e9, 1)}), flush=True)


if __name__ == "__main__":
    main()
