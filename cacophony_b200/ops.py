"""Operator-level Python surface over the C ABI: torch CUDA tensors in, torch CUDA tensors out.

Every function validates device / dtype / contiguity, allocates the output with torch (PyTorch owns all memory),
and enqueues ONE library call on the current CUDA stream.  Errors are Python exceptions (ValueError for bad
arguments, CacoError for a failing library call); nothing falls back to torch math.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

import functools

from . import _lib as L


def _on_tensor_device(fn):
    """Run `fn` with the CUDA device of its tensor arguments current (the library enqueues on the current device's
    current stream), and require all tensor arguments to live on ONE device — a model on cuda:1 must work while cuda:0 is the
    process's current device."""
    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        dev = None
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if dev is None:
                    dev = a.device
                elif a.device != dev:
                    raise ValueError(f"{fn.__name__}: tensor arguments live on different devices ({dev} and {a.device})")
        if dev is None:
            return fn(*args, **kwargs)          # argument validation below raises the proper error
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapped


def _need(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise ValueError(f"{name}: expected a torch.Tensor")
    if not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA tensor (cacophony_b200 has no CPU path)")
    if t.dtype != dtype:
        raise ValueError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous tensor")
    return t


@_on_tensor_device
def cast_f16(x: torch.Tensor) -> torch.Tensor:
    _need(x, torch.float32, "x")
    out = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    L.check(L.load().caco_cast_f32_f16(L.ptr(x), L.ptr(out), x.numel(), L.stream_ptr()), "caco_cast_f32_f16")
    return out


@_on_tensor_device
def gemm_f16(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], epi: int,
             resid: Optional[torch.Tensor] = None, variant: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """epilogue(a[M,K] @ w[N,K].T + bias); a, w fp16; out fp16 (EPI_*_F16) or fp32."""
    _need(a, torch.float16, "a")
    _need(w, torch.float16, "w")
    if a.dim() != 2 or w.dim() != 2 or a.shape[1] != w.shape[1]:
        raise ValueError("gemm_f16: a[M,K], w[N,K] expected")
    M, K = a.shape
    N = w.shape[0]
    if bias is not None:
        _need(bias, torch.float32, "bias")
        if bias.numel() != N:
            raise ValueError("gemm_f16: bias must have N elements")
    out_dtype = torch.float32 if epi in (L.EPI_BIAS_F32, L.EPI_BIAS_RESID_F32) else torch.float16
    if epi == L.EPI_BIAS_RESID_F32:
        if resid is None:
            raise ValueError("gemm_f16: EPI_BIAS_RESID_F32 needs resid")
        _need(resid, torch.float32, "resid")
        if tuple(resid.shape) != (M, N):
            raise ValueError("gemm_f16: resid must be [M,N]")
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    else:
        _need(out, out_dtype, "out")
    rc = L.load().caco_gemm_f16(L.ptr(a), K, L.ptr(w), K, L.ptr(bias), L.ptr(resid), N, L.ptr(out), N, M, N, K, epi,
                                variant, L.stream_ptr())
    L.check(rc, "caco_gemm_f16")
    return out


@_on_tensor_device
def cast_f16_split(w: torch.Tensor) -> torch.Tensor:
    """w [N, K] fp32 -> [N, 2K] fp16 = fp16(w) | fp16(w - fp16(w)): the split-weight packing."""
    _need(w, torch.float32, "w")
    if w.dim() != 2:
        raise ValueError("cast_f16_split: w must be [N, K]")
    out = torch.empty((w.shape[0], 2 * w.shape[1]), dtype=torch.float16, device=w.device)
    L.check(L.load().caco_cast_f32_f16_split(L.ptr(w), L.ptr(out), w.shape[0], w.shape[1], L.stream_ptr()),
            "caco_cast_f32_f16_split")
    return out


@_on_tensor_device
def gemm_f16_wsplit(a: torch.Tensor, w2: torch.Tensor, bias: Optional[torch.Tensor], epi: int,
                    resid: Optional[torch.Tensor] = None, variant: int = 0) -> torch.Tensor:
    """epilogue(a[M,K] @ (hi + lo)[N,K].T + bias) with w2 = cast_f16_split(w) [N, 2K]."""
    _need(a, torch.float16, "a")
    _need(w2, torch.float16, "w2")
    M, K = a.shape
    N = w2.shape[0]
    if w2.shape[1] != 2 * K:
        raise ValueError("gemm_f16_wsplit: w2 must be [N, 2K]")
    out_dtype = torch.float32 if epi in (L.EPI_BIAS_F32, L.EPI_BIAS_RESID_F32) else torch.float16
    out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    L.check(L.load().caco_gemm_f16_wsplit(L.ptr(a), K, L.ptr(w2), 2 * K, L.ptr(bias), L.ptr(resid), N, L.ptr(out), N, M, N, K,
                                          epi, variant, L.stream_ptr()), "caco_gemm_f16_wsplit")
    return out


@_on_tensor_device
def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5, want_f32: bool = True,
              want_f16: bool = False) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    _need(x, torch.float32, "x"); _need(gamma, torch.float32, "gamma"); _need(beta, torch.float32, "beta")
    dim = x.shape[-1]
    rows = x.numel() // dim
    o32 = torch.empty_like(x) if want_f32 else None
    o16 = torch.empty(x.shape, dtype=torch.float16, device=x.device) if want_f16 else None
    L.check(L.load().caco_layernorm(L.ptr(x), L.ptr(gamma), L.ptr(beta), eps, L.ptr(o32), L.ptr(o16), rows, dim,
                                    L.stream_ptr()), "caco_layernorm")
    return o32, o16


@_on_tensor_device
def audio_add_pos(x: torch.Tensor, time_inds: torch.Tensor, freq_inds: torch.Tensor, freq_emb: torch.Tensor) -> torch.Tensor:
    """In place: x[m] += sincos(time_inds[m]) + freq_emb[freq_inds[m]] (mae.py:135-142)."""
    _need(x, torch.float32, "x"); _need(time_inds, torch.float32, "time_inds"); _need(freq_inds, torch.float32, "freq_inds")
    _need(freq_emb, torch.float32, "freq_emb")
    dim = x.shape[-1]
    rows = x.numel() // dim
    L.check(L.load().caco_audio_add_pos(L.ptr(x), L.ptr(time_inds), L.ptr(freq_inds), L.ptr(freq_emb), freq_emb.shape[0],
                                        rows, dim, L.stream_ptr()), "caco_audio_add_pos")
    return x


@_on_tensor_device
def attention_audio(qkv: torch.Tensor, mask: torch.Tensor, heads: int) -> torch.Tensor:
    """qkv [B,S,3*D] fp16 (q|k|v), mask [B,S] fp32 (1 = keep) -> [B,S,D] fp16."""
    _need(qkv, torch.float16, "qkv"); _need(mask, torch.float32, "mask")
    B, S, D3 = qkv.shape
    D = D3 // 3
    out = torch.empty((B, S, D), dtype=torch.float16, device=qkv.device)
    L.check(L.load().caco_attention_audio(L.ptr(qkv), L.ptr(mask), L.ptr(out), B, S, heads, D // heads, L.stream_ptr()),
            "caco_attention_audio")
    return out


@_on_tensor_device
def attention_text(qkv: torch.Tensor, key_mask: torch.Tensor, heads: int) -> torch.Tensor:
    _need(qkv, torch.float16, "qkv"); _need(key_mask, torch.float32, "key_mask")
    B, T, D3 = qkv.shape
    D = D3 // 3
    out = torch.empty((B, T, D), dtype=torch.float16, device=qkv.device)
    L.check(L.load().caco_attention_text(L.ptr(qkv), L.ptr(key_mask), L.ptr(out), B, T, heads, D // heads, L.stream_ptr()),
            "caco_attention_text")
    return out


@_on_tensor_device
def text_embed_ln(ids, position_ids, word, pos, type0, gamma, beta, eps: float = 1e-5):
    _need(ids, torch.int64, "ids")
    if position_ids is not None:
        _need(position_ids, torch.int64, "position_ids")
    for n, t in (("word", word), ("pos", pos), ("type0", type0), ("gamma", gamma), ("beta", beta)):
        _need(t, torch.float32, n)
    B, T = ids.shape
    dim = word.shape[1]
    o32 = torch.empty((B, T, dim), dtype=torch.float32, device=ids.device)
    o16 = torch.empty((B, T, dim), dtype=torch.float16, device=ids.device)
    L.check(L.load().caco_text_embed_ln(L.ptr(ids), L.ptr(position_ids), L.ptr(word), L.ptr(pos), L.ptr(type0), L.ptr(gamma),
                                        L.ptr(beta), eps, L.ptr(o32), L.ptr(o16), B, T, dim, word.shape[0], pos.shape[0],
                                        L.stream_ptr()), "caco_text_embed_ln")
    return o32, o16


@_on_tensor_device
def attn_pool(hid, mask, u, c, ln_gamma=None, ln_beta=None, ln_eps: float = 1e-5, want_hidden: bool = False):
    _need(hid, torch.float32, "hid"); _need(mask, torch.float32, "mask"); _need(u, torch.float32, "u"); _need(c, torch.float32, "c")
    B, S, dim = hid.shape
    heads = u.shape[0]
    pooled = torch.empty((B, heads, dim), dtype=torch.float32, device=hid.device)
    hid_out = torch.empty_like(hid) if (want_hidden and ln_gamma is not None) else None
    L.check(L.load().caco_attn_pool(L.ptr(hid), L.ptr(mask), L.ptr(u), L.ptr(c), L.ptr(ln_gamma), L.ptr(ln_beta), ln_eps,
                                    L.ptr(hid_out), L.ptr(pooled), B, S, heads, dim, L.stream_ptr()), "caco_attn_pool")
    return pooled, hid_out


@_on_tensor_device
def sgemm_nt(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, alpha: float = 1.0) -> torch.Tensor:
    _need(a, torch.float32, "a"); _need(w, torch.float32, "w")
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    L.check(L.load().caco_sgemm_nt(L.ptr(a), K, L.ptr(w), K, L.ptr(bias), alpha, L.ptr(out), N, M, N, K, L.stream_ptr()),
            "caco_sgemm_nt")
    return out


@_on_tensor_device
def l2norm(x: torch.Tensor, eps: float = 1e-10) -> torch.Tensor:
    _need(x, torch.float32, "x")
    out = torch.empty_like(x)
    L.check(L.load().caco_l2norm(L.ptr(x), L.ptr(out), x.shape[0], x.shape[1], eps, L.stream_ptr()), "caco_l2norm")
    return out


@_on_tensor_device
def sim_logits(a: torch.Tensor, t: torch.Tensor, logit_scale: torch.Tensor, want_ta: bool = True):
    """(exp(logit_scale)*a) @ t.T and its transpose counterpart (caco.py:208-210)."""
    _need(a, torch.float32, "a"); _need(t, torch.float32, "t"); _need(logit_scale, torch.float32, "logit_scale")
    na, dim = a.shape
    nt = t.shape[0]
    at = torch.empty((na, nt), dtype=torch.float32, device=a.device)
    ta = torch.empty((nt, na), dtype=torch.float32, device=a.device) if want_ta else None
    L.check(L.load().caco_sim_logits(L.ptr(a), L.ptr(t), L.ptr(logit_scale), L.ptr(at), L.ptr(ta), na, nt, dim,
                                     L.stream_ptr()), "caco_sim_logits")
    return at, ta


@_on_tensor_device
def frontend(wave: torch.Tensor, max_patches: int, want_log_mel: bool = False, want_f16: bool = False):
    """wave [B, L] fp32 CUDA -> dict(audio_patches [B,P,256], audio_time_inds, audio_freq_inds, audio_mask [B,P])."""
    _need(wave, torch.float32, "wave")
    if wave.dim() != 2:
        raise ValueError("frontend: wave must be [batch, n_samples]")
    B, n = wave.shape
    dev = wave.device
    patches = torch.empty((B, max_patches, 256), dtype=torch.float32, device=dev)
    p16 = torch.empty((B, max_patches, 256), dtype=torch.float16, device=dev) if want_f16 else None
    ti = torch.empty((B, max_patches), dtype=torch.float32, device=dev)
    fi = torch.empty_like(ti)
    mk = torch.empty_like(ti)
    n_frames = (n + 159) // 160
    mel = torch.empty((B, n_frames, 128), dtype=torch.float32, device=dev) if want_log_mel else None
    L.check(L.load().caco_frontend(L.ptr(wave), B, n, max_patches, L.ptr(patches), L.ptr(p16), L.ptr(ti), L.ptr(fi), L.ptr(mk),
                                   L.ptr(mel), L.stream_ptr()), "caco_frontend")
    out = {"audio_patches": patches, "audio_time_inds": ti, "audio_freq_inds": fi, "audio_mask": mk}
    if want_log_mel:
        out["log_mel"] = mel
    if want_f16:
        out["audio_patches_f16"] = p16
    return out


@_on_tensor_device
def frontend_ragged(wave: torch.Tensor, lengths: torch.Tensor, max_patches: int, want_f16: bool = False):
    """Ragged batch: wave [B, stride] fp32 (clip b = wave[b, :lengths[b]]), lengths [B] int32 CUDA -> the same dict as
    `frontend`, every clip treated as eval_caco_torch.py:181-206 treats a single clip of its own length."""
    _need(wave, torch.float32, "wave"); _need(lengths, torch.int32, "lengths")
    if wave.dim() != 2 or lengths.dim() != 1 or lengths.shape[0] != wave.shape[0]:
        raise ValueError("frontend_ragged: wave [batch, stride], lengths [batch] expected")
    B, stride = wave.shape
    dev = wave.device
    patches = torch.empty((B, max_patches, 256), dtype=torch.float32, device=dev)
    p16 = torch.empty((B, max_patches, 256), dtype=torch.float16, device=dev) if want_f16 else None
    ti = torch.empty((B, max_patches), dtype=torch.float32, device=dev)
    fi = torch.empty_like(ti)
    mk = torch.empty_like(ti)
    L.check(L.load().caco_frontend_ragged(L.ptr(wave), L.ptr(lengths), B, stride, max_patches, L.ptr(patches), L.ptr(p16),
                                          L.ptr(ti), L.ptr(fi), L.ptr(mk), L.stream_ptr()), "caco_frontend_ragged")
    out = {"audio_patches": patches, "audio_time_inds": ti, "audio_freq_inds": fi, "audio_mask": mk}
    if want_f16:
        out["audio_patches_f16"] = p16
    return out


@_on_tensor_device
def topk_rows(x: torch.Tensor, k: int, want_values: bool = False):
    """argsort(-x, dim=-1)[:, :k] (ties: lower column first; NaN last) for x [rows, cols] fp32, k <= 32 -> int32 [rows, k]."""
    _need(x, torch.float32, "x")
    if x.dim() != 2:
        raise ValueError("topk_rows: x must be [rows, cols]")
    rows, cols = x.shape
    if not (1 <= k <= 32 and k <= cols):
        raise ValueError("topk_rows: need 1 <= k <= min(32, cols)")
    idx = torch.empty((rows, k), dtype=torch.int32, device=x.device)
    val = torch.empty((rows, k), dtype=torch.float32, device=x.device) if want_values else None
    if rows:
        L.check(L.load().caco_topk_rows(L.ptr(x), rows, cols, cols, k, L.ptr(idx), L.ptr(val), L.stream_ptr()), "caco_topk_rows")
    return (idx, val) if want_values else idx


@_on_tensor_device
def retrieval_hits(topk: torch.Tensor, key_id: torch.Tensor, gt_id: torch.Tensor, gt_pairs: Optional[torch.Tensor] = None,
                   n_key_ids: int = 0) -> torch.Tensor:
    """Hit bit-mask per query (bit j = rank j+1 is a hit), eval_utils.py:26-41.  gt_pairs None = 'ta' mode."""
    _need(topk, torch.int32, "topk"); _need(key_id, torch.int32, "key_id"); _need(gt_id, torch.int32, "gt_id")
    if topk.dim() != 2 or topk.shape[1] < 10:
        raise ValueError("retrieval_hits: topk must be [queries, >= 10]")
    Q = topk.shape[0]
    if gt_id.numel() != Q:
        raise ValueError("retrieval_hits: gt_id must have one entry per query")
    mode = 0
    if gt_pairs is not None:
        _need(gt_pairs, torch.int64, "gt_pairs")
        mode = 1
    out = torch.empty((Q,), dtype=torch.int32, device=topk.device)
    L.check(L.load().caco_retrieval_hits(L.ptr(topk), topk.shape[1], Q, L.ptr(key_id), L.ptr(gt_id), L.ptr(gt_pairs),
                                         0 if gt_pairs is None else gt_pairs.numel(), int(n_key_ids), mode, L.ptr(out),
                                         L.stream_ptr()), "caco_retrieval_hits")
    return out


@_on_tensor_device
def avg_pool_tokens(hid: torch.Tensor, group: int = 8) -> torch.Tensor:
    """[B, S, D] -> [B, S // group, D] mean over groups of `group` consecutive tokens (caco_embeddings.py:124-125)."""
    _need(hid, torch.float32, "hid")
    B, S, D = hid.shape
    out = torch.empty((B, S // group, D), dtype=torch.float32, device=hid.device)
    L.check(L.load().caco_avg_pool_tokens(L.ptr(hid), B, S, D, group, L.ptr(out), L.stream_ptr()), "caco_avg_pool_tokens")
    return out
