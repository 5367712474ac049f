"""CPU oracle for the Cacophony inference hot path.  TEST INFRASTRUCTURE ONLY.

A functional fp32 restatement (torch CPU tensors used as plain arrays + numpy) of the reference
algorithm, written from the reference's behaviour; every function cites the reference lines it
follows (paths relative to ``/root/reference``).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import this module; the product package
``cacophony_b200`` never does and has no CPU fallback.

Parity status: PINNED against the reference itself — ``oracle/make_golden.py`` imports
``/root/reference/src`` in the authoring container, runs it on the synthetic weights/inputs of
``oracle/weights.py`` and commits the outputs under ``tests/golden/``; ``tests/test_oracle.py``
checks this restatement against those vectors.  (The reference ships no tests or golden vectors
of its own — SURVEY.md §4.)

``Rounding`` lets a test emulate the operand rounding of the CUDA path (fp16 GEMM inputs, fp32
accumulate) on the CPU, to separate "rounding by design" from "kernel bug".
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, Optional, Tuple

import numpy as np
import torch

NORM_EPS = 1e-10            # src/caco_torch/caco.py:14
LN_EPS = 1e-5               # nn.LayerNorm default (mae.py:68,76,123) and roberta.py:22
SR, HOP, WIN, NFFT, NMELS = 16000, 160, 400, 512, 128
LOG_EPS, MEL_SCALE, MEL_BIAS = 1e-5, 0.2, 0.9


# ------------------------------------------------------------------------------------------
# operand rounding emulation
# ------------------------------------------------------------------------------------------
@dataclass
class Rounding:
    """How GEMM operands are rounded before the fp32-accumulated product (None = exact fp32)."""
    act: Optional[Callable[[torch.Tensor], torch.Tensor]] = None
    wgt: Optional[Callable[[torch.Tensor], torch.Tensor]] = None

    @staticmethod
    def fp16(split_weights: bool = False) -> "Rounding":
        h = lambda x: x.half().float()
        if split_weights:      # w ≈ hi + lo, both fp16  (two MMA passes on the device)
            return Rounding(act=h, wgt=lambda w: h(w) + h(w - h(w)))
        return Rounding(act=h, wgt=h)

    @staticmethod
    def bf16() -> "Rounding":
        b = lambda x: x.bfloat16().float()
        return Rounding(act=b, wgt=b)


def _ra(x, r: Optional[Rounding]):
    return x if r is None or r.act is None else r.act(x)


def _rw(w, r: Optional[Rounding]):
    return w if r is None or r.wgt is None else r.wgt(w)


def linear(x, w, b, r: Optional[Rounding] = None):
    """y = x·wᵀ + b  (nn.Linear)."""
    return _ra(x, r) @ _rw(w, r).t() + b


def layer_norm(x, g, b, eps: float = LN_EPS):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


# ------------------------------------------------------------------------------------------
# frontend: waveform -> log-mel -> patches      (src/eval/eval_caco_torch.py:41-151)
# ------------------------------------------------------------------------------------------
def hann_periodic(n: int = WIN) -> torch.Tensor:
    """torch.hann_window(400) (periodic) as used at eval_caco_torch.py:86."""
    k = torch.arange(n, dtype=torch.float32)
    return (0.5 - 0.5 * torch.cos(2.0 * math.pi * k / n)).float()


def mel_filterbank(n_freqs: int = NFFT // 2 + 1, f_min: float = 0.0, f_max: float = SR / 2,
                   n_mels: int = NMELS, sample_rate: int = SR) -> torch.Tensor:
    """HTK-scale triangular filterbank, norm=None  — the published algorithm of
    ``torchaudio.functional.melscale_fbanks`` (torchaudio 2.5.1 pinned in requirements_torch.txt:51;
    called at eval_caco_torch.py:94-101).  Returns [n_freqs, n_mels] fp32."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


def num_frames(n_samples: int) -> int:
    return (n_samples + HOP - 1) // HOP            # eval_caco_torch.py:67


def log_mel(wave: torch.Tensor) -> torch.Tensor:
    """[L] fp32 -> [frames,128] fp32.  eval_caco_torch.py:63-104: ceil(L/160) frames, zero tail pad
    to (frames-1)*160+512, 512-point frames with the 400-tap periodic Hann window centred inside
    the FFT frame (torch.stft pads the window by (512-400)/2 = 56 on both sides), |rfft|, HTK mel,
    log(x+1e-5)*0.2+0.9."""
    x = wave.reshape(-1).float()
    n = num_frames(x.numel())
    need = (n - 1) * HOP + NFFT
    if need > x.numel():
        x = torch.cat([x, torch.zeros(need - x.numel())])
    win = torch.zeros(NFFT)
    win[(NFFT - WIN) // 2:(NFFT - WIN) // 2 + WIN] = hann_periodic()
    frames = x.unfold(0, NFFT, HOP)[:n] * win                       # [n,512]
    spec = torch.fft.rfft(frames, dim=-1).abs()                      # [n,257]
    mel = spec @ mel_filterbank()
    return torch.log(mel + LOG_EPS) * MEL_SCALE + MEL_BIAS


def patchify(mel: np.ndarray, max_patches: int, tp: int = 16, fp: int = 16) -> Dict[str, np.ndarray]:
    """eval_caco_torch.py:108-151: 16x16 patches, token p = 8*t + f, element dt*16+df; keep the first
    ``max_patches`` or zero-pad; padded slots carry time/freq index 0 and mask 0; all float32."""
    mel = np.asarray(mel, dtype=np.float32)
    nt = mel.shape[0] // tp
    nf = mel.shape[1] // fp
    full = nt * nf
    x = mel[: nt * tp].reshape(nt, tp, nf, fp).transpose(0, 2, 1, 3).reshape(full, tp * fp)
    p = np.arange(max_patches)
    if full > max_patches:
        x = x[:max_patches]
        mask = np.ones(max_patches, np.float32)
        t_ind, f_ind = p // nf, p % nf
    else:
        mask = (p < full).astype(np.float32)
        live = (mask * p).astype(np.int64)
        t_ind, f_ind = live // nf, live % nf
        x = np.concatenate([x, np.zeros((max_patches - full, tp * fp), np.float32)], 0)
    return {"audio_patches": x.astype(np.float32), "audio_time_inds": t_ind.astype(np.float32),
            "audio_freq_inds": f_ind.astype(np.float32), "audio_mask": mask}


def prepare_audio_batch(waves, max_patches: int = 500) -> Dict[str, torch.Tensor]:
    """Batched form of eval_caco_torch.py:181-206 (the reference does one clip per call)."""
    outs = [patchify(log_mel(torch.as_tensor(w)).numpy(), max_patches) for w in waves]
    return {k: torch.from_numpy(np.stack([o[k] for o in outs])) for k in outs[0]}


# ------------------------------------------------------------------------------------------
# audio tower    (src/caco_torch/audio_models/mae.py:47-148, caco.py:24-79,123-150)
# ------------------------------------------------------------------------------------------
def sincos_time_embed(t: torch.Tensor, dim: int) -> torch.Tensor:
    """mae.py:102-109: cat[sin(t·ω_i), cos(t·ω_i)], ω_i = exp(2i·(−ln 1e4)/dim), i < dim/2."""
    w = torch.exp(2 * torch.arange(dim // 2, dtype=torch.float32) * -math.log(10000.0) / dim)
    a = t.unsqueeze(-1) * w
    return torch.cat([torch.sin(a), torch.cos(a)], -1)


def mha_self(x, mask, w_in, b_in, w_out, b_out, heads: int, r=None):
    """nn.MultiheadAttention(batch_first, key_padding_mask=(mask==0))  (mae.py:69-74,89-92):
    packed in-proj rows q|k|v, q scaled by 1/sqrt(dh), masked keys -> -inf, softmax, out-proj."""
    B, S, D = x.shape
    dh = D // heads
    qkv = linear(x, w_in, b_in, r)
    q, k, v = [z.reshape(B, S, heads, dh).transpose(1, 2) for z in qkv.split(D, -1)]
    q = q * (1.0 / math.sqrt(dh))
    s = _ra(q, r) @ _ra(k, r).transpose(-1, -2)
    s = s.masked_fill((mask == 0)[:, None, None, :], float("-inf"))
    p = torch.softmax(s, -1)
    o = (_ra(p, r) @ _ra(v, r)).transpose(1, 2).reshape(B, S, D)
    return linear(o, w_out, b_out, r)


def audio_encoder(sd, patches, t_inds, f_inds, mask, heads: int = 8, r=None,
                  taps: Optional[dict] = None) -> torch.Tensor:
    """AudioEncoder.forward (mae.py:125-148): input proj + sin/cos time + learned freq pos-emb,
    pre-LN blocks (mae.py:80-99) with SiLU MLP (mae.py:47-61), final LN."""
    P = "audio_module."
    x = linear(patches, sd[P + "input_proj.weight"], sd[P + "input_proj.bias"], r)
    x = x + sincos_time_embed(t_inds, x.shape[-1])
    x = x + sd[P + "freq_positional_embedding"][f_inds.long()]
    if taps is not None:
        taps["audio_embed"] = x
    i = 0
    while f"{P}layers.{i}.norm1.weight" in sd:
        L = f"{P}layers.{i}."
        h = layer_norm(x, sd[L + "norm1.weight"], sd[L + "norm1.bias"])
        x = x + mha_self(h, mask, sd[L + "attn.in_proj_weight"], sd[L + "attn.in_proj_bias"],
                         sd[L + "attn.out_proj.weight"], sd[L + "attn.out_proj.bias"], heads, r)
        h = layer_norm(x, sd[L + "norm2.weight"], sd[L + "norm2.bias"])
        h = torch.nn.functional.silu(linear(h, sd[L + "mlp.fc1.weight"], sd[L + "mlp.fc1.bias"], r))
        x = x + linear(h, sd[L + "mlp.fc2.weight"], sd[L + "mlp.fc2.bias"], r)
        if taps is not None:
            taps[f"audio_layer{i}"] = x
        i += 1
    return layer_norm(x, sd[P + "norm.weight"], sd[P + "norm.bias"])


def audio_pool(sd, hidden, mask, heads: int = 2, r=None) -> torch.Tensor:
    """AudioAttentionPooler.forward (caco.py:41-79)."""
    P = "audio_attention_pool."
    B, S, D = hidden.shape
    dh = D // heads
    kv = linear(hidden, sd[P + "kv_proj.weight"], sd[P + "kv_proj.bias"], r)
    k, v = kv[..., :D].reshape(B, S, heads, dh), kv[..., D:].reshape(B, S, heads, dh)
    q = sd[P + "query"].reshape(heads, dh) * (1.0 / math.sqrt(dh))
    a = torch.einsum("hd,bjhd->bhj", q, k)
    a = a.masked_fill((mask == 0)[:, None, :], float("-inf"))
    a = torch.softmax(a, -1)
    o = torch.einsum("bhj,bjhd->bhd", a, v).reshape(B, D)
    return linear(o, sd[P + "out_proj.weight"], sd[P + "out_proj.bias"], r)


def l2_normalize(e):
    """caco.py:146,173: e / ||e + 1e-10||₂ (epsilon added inside the norm, per element)."""
    return e / torch.norm(e + NORM_EPS, dim=-1, keepdim=True)


def get_audio_embedding(sd, patches, t_inds, f_inds, mask, normalize=False, r=None, taps=None,
                        audio_heads: int = 8, pool_heads: int = 2):
    """CACO.get_audio_embedding (caco.py:123-150) -> (embedding, hidden)."""
    hid = audio_encoder(sd, patches, t_inds, f_inds, mask, audio_heads, r, taps)
    emb = audio_pool(sd, hid, mask, pool_heads, r)
    if normalize:
        emb = l2_normalize(emb)
    return emb, hid


# ------------------------------------------------------------------------------------------
# text tower    (src/caco_torch/text_models/roberta.py:26-326, caco.py:152-177)
# ------------------------------------------------------------------------------------------
def text_encoder(sd, ids, mask, heads: int = 12, r=None, taps=None, position_ids=None):
    """RobertaModel.forward (roberta.py:283-326): position ids arange(T) (:292-293), causal AND
    key-padding mask as additive 0/-inf bias (:297-310), embeddings+LN (:35-53), post-LN layers
    (:191-215) with erf-GELU (:156-157), 1-query attention pooler (:253-271)."""
    P = "text_module."
    B, T = ids.shape
    pos = torch.arange(T).unsqueeze(0).expand(B, T) if position_ids is None else position_ids
    E = P + "embeddings."
    x = sd[E + "word_embeddings.weight"][ids] + sd[E + "position_embeddings.weight"][pos] \
        + sd[E + "token_type_embeddings.weight"][torch.zeros_like(ids)]
    x = layer_norm(x, sd[E + "LayerNorm.weight"], sd[E + "LayerNorm.bias"])
    if taps is not None:
        taps["text_embed"] = x
    allow = torch.tril(torch.ones(T, T, dtype=torch.bool))[None, None] & mask[:, None, None, :].bool()
    D = x.shape[-1]
    dh = D // heads
    i = 0
    while f"{P}encoder.layers.{i}.attention.self.query.weight" in sd:
        L = f"{P}encoder.layers.{i}."
        A = L + "attention.self."
        q = linear(x, sd[A + "query.weight"], sd[A + "query.bias"], r)
        k = linear(x, sd[A + "key.weight"], sd[A + "key.bias"], r)
        v = linear(x, sd[A + "value.weight"], sd[A + "value.bias"], r)
        q, k, v = [z.reshape(B, T, heads, dh).transpose(1, 2) for z in (q, k, v)]
        s = (_ra(q, r) @ _ra(k, r).transpose(-1, -2)) / math.sqrt(dh)
        s = s.masked_fill(~allow, float("-inf"))
        p = torch.softmax(s, -1)
        o = (_ra(p, r) @ _ra(v, r)).transpose(1, 2).reshape(B, T, D)
        a = layer_norm(linear(o, sd[L + "attention.output.dense.weight"],
                              sd[L + "attention.output.dense.bias"], r) + x,
                       sd[L + "attention.output.LayerNorm.weight"],
                       sd[L + "attention.output.LayerNorm.bias"])
        h = torch.nn.functional.gelu(linear(a, sd[L + "intermediate.dense.weight"],
                                            sd[L + "intermediate.dense.bias"], r))
        x = layer_norm(linear(h, sd[L + "output.dense.weight"], sd[L + "output.dense.bias"], r) + a,
                       sd[L + "output.LayerNorm.weight"], sd[L + "output.LayerNorm.bias"])
        if taps is not None:
            taps[f"text_layer{i}"] = x
        i += 1
    Q = P + "pooler."
    key = linear(x, sd[Q + "key_proj.weight"], sd[Q + "key_proj.bias"], r) / math.sqrt(D)
    val = linear(x, sd[Q + "value_proj.weight"], sd[Q + "value_proj.bias"], r)
    w = torch.einsum("mh,bnh->bmn", sd[Q + "attention_pool_query"], key)
    w = w.masked_fill((mask == 0)[:, None, :], float("-inf"))
    w = torch.softmax(w, -1)
    pooled = torch.einsum("bmn,bnh->bmh", w, val)[:, 0]
    return pooled, x


def get_text_embedding(sd, ids, mask, normalize=False, r=None, taps=None, heads: int = 12):
    """CACO.get_text_embedding (caco.py:152-177) -> (embedding, hidden)."""
    pooled, hid = text_encoder(sd, ids, mask, heads, r, taps)
    emb = linear(pooled, sd["text_proj.weight"], sd["text_proj.bias"], r)
    if normalize:
        emb = l2_normalize(emb)
    return emb, hid


# ------------------------------------------------------------------------------------------
# captioning decoder    (src/caco_torch/text_models/roberta.py:329-373, caco.py:214-240)
# ------------------------------------------------------------------------------------------
def _bert_attention(sd, P, x, kv_src, bias, heads, r):
    """RobertaAttention (roberta.py:56-147): q from x, k / v from kv_src, additive bias, dense + post-LN residual."""
    B, Tq, D = x.shape
    dh = D // heads
    q = linear(x, sd[P + "self.query.weight"], sd[P + "self.query.bias"], r).reshape(B, Tq, heads, dh).transpose(1, 2)
    k = linear(kv_src, sd[P + "self.key.weight"], sd[P + "self.key.bias"], r).reshape(B, -1, heads, dh).transpose(1, 2)
    v = linear(kv_src, sd[P + "self.value.weight"], sd[P + "self.value.bias"], r).reshape(B, -1, heads, dh).transpose(1, 2)
    w = torch.softmax(_ra(q, r) @ _ra(k, r).transpose(-1, -2) / math.sqrt(dh) + bias, dim=-1)
    o = (_ra(w, r) @ _ra(v, r)).transpose(1, 2).reshape(B, Tq, D)
    o = linear(o, sd[P + "output.dense.weight"], sd[P + "output.dense.bias"], r)
    return layer_norm(o + x, sd[P + "output.LayerNorm.weight"], sd[P + "output.LayerNorm.bias"])


def decoder_logits(sd, text_hidden, text_mask, audio_hidden, audio_mask, heads: int = 12, r=None):
    """RobertaDecoder.forward (roberta.py:337-373): causal & padded self-attention (:346-355), cross-attention to the audio
    tokens with the audio key mask (:358-361, layer :205-211), GELU MLP (:150-178), decoder_proj (:372)."""
    B, T, _ = text_hidden.shape
    neg = float("-inf")
    allow = torch.tril(torch.ones(T, T, dtype=torch.bool))[None, None] & (text_mask != 0)[:, None, None, :]
    self_bias = torch.zeros(B, 1, T, T).masked_fill(~allow, neg)
    cross_bias = torch.zeros(B, 1, 1, audio_mask.shape[1]).masked_fill((audio_mask == 0)[:, None, None, :], neg)
    x = text_hidden
    i = 0
    while f"decoder_module.encoder.layers.{i}.intermediate.dense.weight" in sd:
        P = f"decoder_module.encoder.layers.{i}."
        a = _bert_attention(sd, P + "attention.", x, x, self_bias, heads, r)
        c = _bert_attention(sd, P + "crossattention.", a, audio_hidden, cross_bias, heads, r)
        h = torch.nn.functional.gelu(linear(c, sd[P + "intermediate.dense.weight"], sd[P + "intermediate.dense.bias"], r))
        y = linear(h, sd[P + "output.dense.weight"], sd[P + "output.dense.bias"], r)
        x = layer_norm(y + c, sd[P + "output.LayerNorm.weight"], sd[P + "output.LayerNorm.bias"])
        i += 1
    return linear(x, sd["decoder_module.decoder_proj.weight"], sd["decoder_module.decoder_proj.bias"], r)


def get_decoder_logits(sd, audio_hidden, audio_mask, ids, text_mask, r=None):
    """CACO.get_decoder_logits (caco.py:214-240): text tower hidden state -> decoder."""
    _, text_hidden = get_text_embedding(sd, ids, text_mask, r=r)
    return decoder_logits(sd, text_hidden, text_mask, audio_hidden, audio_mask, r=r)


class IncrementalDecoder:
    """The decode loop's call (eval_caco_torch.py:411-472 -> CACO.get_decoder_logits on the prefix so far) restated
    incrementally: text tower (roberta.py:283-326) and decoder (roberta.py:337-373) are causal, so the keys / values of the
    tokens already pushed never change — ``step`` pushes ONE token per sequence, appends its keys / values to per-layer caches
    and returns the next-token logits ``get_decoder_logits(...)[:, t]``; the cross-attention keys / values of the audio tokens
    are computed once (what the JAX twin caches, src/caco/caco.py:154-230).  This is the algorithm of the library's
    caco_model_decode_begin / caco_model_decode_step; tests/test_oracle.py checks it against the full-prefix oracle."""

    def __init__(self, sd, audio_hidden, audio_mask, heads: int = 12, r=None):
        self.sd, self.heads, self.r = sd, heads, r
        self.B, _, self.D = audio_hidden.shape
        self.cross_bias = torch.zeros(self.B, 1, 1, audio_mask.shape[1]).masked_fill((audio_mask == 0)[:, None, None, :], float("-inf"))
        self.t = 0
        self.text_kv, self.dec_kv, self.cross_kv = {}, {}, {}
        i = 0
        while f"decoder_module.encoder.layers.{i}.intermediate.dense.weight" in sd:
            P = f"decoder_module.encoder.layers.{i}.crossattention.self."
            self.cross_kv[i] = (self._heads(linear(audio_hidden, sd[P + "key.weight"], sd[P + "key.bias"], r)),
                                self._heads(linear(audio_hidden, sd[P + "value.weight"], sd[P + "value.bias"], r)))
            i += 1
        self.n_dec = i

    def _heads(self, z):
        return z.reshape(self.B, -1, self.heads, self.D // self.heads).transpose(1, 2)

    def _attend(self, P, x, k, v, bias):
        """RobertaAttention for one query row per sequence against given keys / values (roberta.py:86-123)."""
        dh = self.D // self.heads
        q = self._heads(linear(x, self.sd[P + "self.query.weight"], self.sd[P + "self.query.bias"], self.r))
        s = _ra(q, self.r) @ _ra(k, self.r).transpose(-1, -2) / math.sqrt(dh)
        w = torch.softmax(s if bias is None else s + bias, dim=-1)
        o = (_ra(w, self.r) @ _ra(v, self.r)).transpose(1, 2).reshape(self.B, 1, self.D)
        o = linear(o, self.sd[P + "output.dense.weight"], self.sd[P + "output.dense.bias"], self.r)
        return layer_norm(o + x, self.sd[P + "output.LayerNorm.weight"], self.sd[P + "output.LayerNorm.bias"])

    def _self_block(self, P, x, cache, i):
        k = self._heads(linear(x, self.sd[P + "self.key.weight"], self.sd[P + "self.key.bias"], self.r))
        v = self._heads(linear(x, self.sd[P + "self.value.weight"], self.sd[P + "self.value.bias"], self.r))
        if i in cache:
            k, v = torch.cat([cache[i][0], k], dim=2), torch.cat([cache[i][1], v], dim=2)
        cache[i] = (k, v)
        return self._attend(P, x, k, v, None)              # every cached key is at a position <= t: the causal mask is all-open

    def _mlp(self, L, x):
        sd, r = self.sd, self.r
        h = torch.nn.functional.gelu(linear(x, sd[L + "intermediate.dense.weight"], sd[L + "intermediate.dense.bias"], r))
        y = linear(h, sd[L + "output.dense.weight"], sd[L + "output.dense.bias"], r)
        return layer_norm(y + x, sd[L + "output.LayerNorm.weight"], sd[L + "output.LayerNorm.bias"])

    def step(self, ids):
        """ids [B] int64: the token at position t of every sequence -> next-token logits [B, vocab]."""
        sd = self.sd
        E = "text_module.embeddings."
        ids = ids.reshape(self.B, 1)
        x = sd[E + "word_embeddings.weight"][ids] + sd[E + "position_embeddings.weight"][torch.full_like(ids, self.t)] \
            + sd[E + "token_type_embeddings.weight"][torch.zeros_like(ids)]
        x = layer_norm(x, sd[E + "LayerNorm.weight"], sd[E + "LayerNorm.bias"])
        i = 0
        while f"text_module.encoder.layers.{i}.attention.self.query.weight" in sd:
            L = f"text_module.encoder.layers.{i}."
            x = self._mlp(L, self._self_block(L + "attention.", x, self.text_kv, i))
            i += 1
        for i in range(self.n_dec):
            L = f"decoder_module.encoder.layers.{i}."
            a = self._self_block(L + "attention.", x, self.dec_kv, i)
            c = self._attend(L + "crossattention.", a, self.cross_kv[i][0], self.cross_kv[i][1], self.cross_bias)
            x = self._mlp(L, c)
        self.t += 1
        return linear(x, sd["decoder_module.decoder_proj.weight"], sd["decoder_module.decoder_proj.bias"], self.r)[:, 0]


def contrastive_logits(sd, a_emb, t_emb):
    """caco.py:208-210: scale = exp(logit_scale); (scale·A)·Tᵀ and (scale·T)·Aᵀ."""
    s = torch.exp(sd["logit_scale"])
    return (s * a_emb) @ t_emb.t(), (s * t_emb) @ a_emb.t()


def forward(sd, patches, t_inds, f_inds, mask, ids, text_mask, r=None):
    """CACO.forward (caco.py:242-261) -> (at_logits, ta_logits, audio_emb, text_emb)."""
    a, _ = get_audio_embedding(sd, patches, t_inds, f_inds, mask, normalize=True, r=r)
    t, _ = get_text_embedding(sd, ids, text_mask, normalize=True, r=r)
    at, ta = contrastive_logits(sd, a, t)
    return at, ta, a, t


def zero_shot_top1(sd, a_emb, t_emb) -> torch.Tensor:
    """eval_caco_torch.py:330-331: argsort(-exp(logit_scale)·a·Tᵀ)[:, 0]."""
    logits = torch.exp(sd["logit_scale"]) * a_emb @ t_emb.t()
    return torch.argsort(-logits, dim=-1)[:, 0]
