"""Deterministic synthetic CACO weights (TEST INFRASTRUCTURE — not a product path).

The pretrained ``Cacophony.ckpt`` is not available offline, and the
reference cannot travel to the GPU box, so ``torch.manual_seed(k); create_caco_model()``
cannot be the shared weight source.  This module regenerates, from a numpy PCG64 stream,
a ``state_dict`` with exactly the keys/shapes of the reference encoder path
(``/root/reference/src/caco_torch/caco.py:100-121`` and the sub-modules it owns) so that

* here, the reference model ``load_state_dict``s it and produces the golden vectors, and
* on the GPU box, the CUDA path and the oracle port load the very same numbers.

The distributions follow PyTorch's default initialisers for the same layers (uniform
``±1/sqrt(fan_in)`` Linear, xavier-uniform MHA in-proj, N(0,1) embeddings, N(0,0.02²)
queries) but LayerNorm gains/biases and the MHA biases are randomised instead of 1/0 so
that every parameter on the path influences the result.  ``sharp > 1`` scales the q/k
projections so that softmax rows become peaky (harder numerics than near-uniform attention);
``outlier=True`` adds the outlier channels trained checkpoints have (see ``_apply_outliers``).
"""
from __future__ import annotations

import math
from typing import Dict, Iterator, List, Tuple

import numpy as np

AUDIO_LAYERS = 12
TEXT_LAYERS = 12
HIDDEN = 768
FFN = 3072
PATCH = 256
VOCAB = 50265
MAX_POS = 514
FREQ_PATCHES = 8
LOGIT_SCALE_INIT = 2.6592


def param_spec(audio_layers: int = AUDIO_LAYERS, text_layers: int = TEXT_LAYERS,
               vocab: int = VOCAB, decoder_layers: int = 0) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, shape, kind) for every tensor on the encoder path, in generation order.

    Key names: reference ``state_dict``; ``decoder_module.*`` is out of scope.
    kind: lin_w / lin_b (fan_in = last dim of the matching weight) / xavier / ln_w / ln_b /
    emb / small (N(0,0.02²)) / scalar.
    """
    D, F = HIDDEN, FFN
    s: List[Tuple[str, Tuple[int, ...], str]] = []
    s.append(("logit_scale", (), "scalar"))
    s.append(("audio_module.freq_positional_embedding", (FREQ_PATCHES, D), "small"))
    s.append(("audio_module.input_proj.weight", (D, PATCH), "lin_w"))
    s.append(("audio_module.input_proj.bias", (D,), f"lin_b:{PATCH}"))
    for i in range(audio_layers):
        p = f"audio_module.layers.{i}."
        s += [(p + "norm1.weight", (D,), "ln_w"), (p + "norm1.bias", (D,), "ln_b"),
              (p + "attn.in_proj_weight", (3 * D, D), "xavier"),
              (p + "attn.in_proj_bias", (3 * D,), f"lin_b:{D}"),
              (p + "attn.out_proj.weight", (D, D), "lin_w"),
              (p + "attn.out_proj.bias", (D,), f"lin_b:{D}"),
              (p + "norm2.weight", (D,), "ln_w"), (p + "norm2.bias", (D,), "ln_b"),
              (p + "mlp.fc1.weight", (F, D), "lin_w"), (p + "mlp.fc1.bias", (F,), f"lin_b:{D}"),
              (p + "mlp.fc2.weight", (D, F), "lin_w"), (p + "mlp.fc2.bias", (D,), f"lin_b:{F}")]
    s += [("audio_module.norm.weight", (D,), "ln_w"), ("audio_module.norm.bias", (D,), "ln_b")]
    s += [("audio_attention_pool.query", (D,), "small"),
          ("audio_attention_pool.kv_proj.weight", (2 * D, D), "lin_w"),
          ("audio_attention_pool.kv_proj.bias", (2 * D,), f"lin_b:{D}"),
          ("audio_attention_pool.out_proj.weight", (D, D), "lin_w"),
          ("audio_attention_pool.out_proj.bias", (D,), f"lin_b:{D}")]
    e = "text_module.embeddings."
    s += [(e + "word_embeddings.weight", (vocab, D), "emb"),
          (e + "position_embeddings.weight", (MAX_POS, D), "emb"),
          (e + "token_type_embeddings.weight", (1, D), "emb"),
          (e + "LayerNorm.weight", (D,), "ln_w"), (e + "LayerNorm.bias", (D,), "ln_b")]
    for i in range(text_layers):
        p = f"text_module.encoder.layers.{i}."
        for nm in ("query", "key", "value"):
            s += [(p + f"attention.self.{nm}.weight", (D, D), "lin_w"),
                  (p + f"attention.self.{nm}.bias", (D,), f"lin_b:{D}")]
        s += [(p + "attention.output.dense.weight", (D, D), "lin_w"),
              (p + "attention.output.dense.bias", (D,), f"lin_b:{D}"),
              (p + "attention.output.LayerNorm.weight", (D,), "ln_w"),
              (p + "attention.output.LayerNorm.bias", (D,), "ln_b"),
              (p + "intermediate.dense.weight", (F, D), "lin_w"),
              (p + "intermediate.dense.bias", (F,), f"lin_b:{D}"),
              (p + "output.dense.weight", (D, F), "lin_w"),
              (p + "output.dense.bias", (D,), f"lin_b:{F}"),
              (p + "output.LayerNorm.weight", (D,), "ln_w"),
              (p + "output.LayerNorm.bias", (D,), "ln_b")]
    s += [("text_module.pooler.attention_pool_query", (1, D), "small"),
          ("text_module.pooler.key_proj.weight", (D, D), "lin_w"),
          ("text_module.pooler.key_proj.bias", (D,), f"lin_b:{D}"),
          ("text_module.pooler.value_proj.weight", (D, D), "lin_w"),
          ("text_module.pooler.value_proj.bias", (D,), f"lin_b:{D}"),
          ("text_proj.weight", (D, D), "lin_w"), ("text_proj.bias", (D,), f"lin_b:{D}")]
    # captioning head (roberta.py:329-373), APPENDED so that the encoder-path tensors of a seed do not depend on it
    for i in range(decoder_layers):
        p = f"decoder_module.encoder.layers.{i}."
        for blk in ("attention", "crossattention"):
            for nm in ("query", "key", "value"):
                s += [(p + f"{blk}.self.{nm}.weight", (D, D), "lin_w"), (p + f"{blk}.self.{nm}.bias", (D,), f"lin_b:{D}")]
            s += [(p + f"{blk}.output.dense.weight", (D, D), "lin_w"), (p + f"{blk}.output.dense.bias", (D,), f"lin_b:{D}"),
                  (p + f"{blk}.output.LayerNorm.weight", (D,), "ln_w"), (p + f"{blk}.output.LayerNorm.bias", (D,), "ln_b")]
        s += [(p + "intermediate.dense.weight", (F, D), "lin_w"), (p + "intermediate.dense.bias", (F,), f"lin_b:{D}"),
              (p + "output.dense.weight", (D, F), "lin_w"), (p + "output.dense.bias", (D,), f"lin_b:{F}"),
              (p + "output.LayerNorm.weight", (D,), "ln_w"), (p + "output.LayerNorm.bias", (D,), "ln_b")]
    if decoder_layers:
        s += [("decoder_module.decoder_proj.weight", (vocab, D), "lin_w"), ("decoder_module.decoder_proj.bias", (vocab,), f"lin_b:{D}")]
    return s


def _gen(rng: np.random.Generator, shape, kind: str) -> np.ndarray:
    if kind == "scalar":
        return np.asarray(LOGIT_SCALE_INIT, dtype=np.float32)
    if kind == "small":
        return (0.02 * rng.standard_normal(shape, dtype=np.float32)).astype(np.float32)
    if kind == "emb":
        return rng.standard_normal(shape, dtype=np.float32)
    if kind == "ln_w":
        return (1.0 + 0.1 * rng.standard_normal(shape, dtype=np.float32)).astype(np.float32)
    if kind == "ln_b":
        return (0.05 * rng.standard_normal(shape, dtype=np.float32)).astype(np.float32)
    if kind == "lin_w":
        b = 1.0 / math.sqrt(shape[-1])
    elif kind == "xavier":
        b = math.sqrt(6.0 / (shape[0] + shape[1]))
    elif kind.startswith("lin_b:"):
        b = 1.0 / math.sqrt(int(kind.split(":")[1]))
    else:
        raise ValueError(kind)
    u = rng.random(shape, dtype=np.float32)
    return ((2.0 * u - 1.0) * np.float32(b)).astype(np.float32)


# "checkpoint-like" outliers: what trained transformers have and PyTorch's initialisers do not
OUTLIER_LN_CHANNELS = (7, 77, 161, 305, 449, 701)      # LayerNorm gains x 30 in these channels, every LayerNorm
OUTLIER_LN_GAIN = 30.0
OUTLIER_RESID = ((33, 300.0), (555, -300.0))           # audio residual stream: input_proj.bias pushes two channels to +-300
OUTLIER_TEXT_BETA = ((33, 20.0), (555, -20.0))         # text (post-LN): the same two channels of every LayerNorm bias
OUTLIER_FC1_UNITS = (5, 130, 1027, 2049, 3000)         # fc1 / intermediate biases: pre-activations near +1e3
OUTLIER_FC1_BIAS = 1000.0


def _apply_outliers(name: str, w: np.ndarray) -> np.ndarray:
    is_ln_w = name.endswith(("norm1.weight", "norm2.weight", "norm.weight", "LayerNorm.weight"))
    is_ln_b = name.endswith("LayerNorm.bias")
    if is_ln_w:
        w = w.copy()
        w[list(OUTLIER_LN_CHANNELS)] *= np.float32(OUTLIER_LN_GAIN)
    elif is_ln_b and name.startswith("text_module."):
        w = w.copy()
        for c, v in OUTLIER_TEXT_BETA:
            w[c] = np.float32(v)
    elif name == "audio_module.input_proj.bias":
        w = w.copy()
        for c, v in OUTLIER_RESID:
            w[c] = np.float32(v)
    elif name.endswith(("mlp.fc1.bias", "intermediate.dense.bias")):
        w = w.copy()
        w[list(OUTLIER_FC1_UNITS)] = np.float32(OUTLIER_FC1_BIAS)
    return w


def iter_weights(seed: int, sharp: float = 1.0, outlier: bool = False, **spec_kw) -> Iterator[Tuple[str, np.ndarray]]:
    rng = np.random.default_rng(np.random.PCG64(1000003 * seed + 17))
    D = HIDDEN
    for name, shape, kind in param_spec(**spec_kw):
        w = _gen(rng, shape, kind)
        if outlier:
            w = _apply_outliers(name, w)
        if sharp != 1.0:
            if name.endswith("attn.in_proj_weight") or name.endswith("attn.in_proj_bias"):
                w = w.copy()
                w[: 2 * D] *= np.float32(sharp)          # q and k rows of the packed in-proj
            elif ".attention.self.query." in name or ".attention.self.key." in name:
                w = w * np.float32(sharp)
        yield name, w


def make_state_dict(seed: int, sharp: float = 1.0, outlier: bool = False, **spec_kw) -> Dict[str, "torch.Tensor"]:
    """Synthetic encoder-path ``state_dict`` as CPU fp32 torch tensors.  outlier: checkpoint-like outlier structure (LayerNorm
    gains x 30 in six channels, two residual channels at +-300, a few fc1 pre-activations near 1e3)."""
    import torch
    return {k: torch.from_numpy(np.array(v, dtype=np.float32, copy=True, order="C"))          # keeps logit_scale 0-d
            for k, v in iter_weights(seed, sharp, outlier, **spec_kw)}


# ---------------------------------------------------------------------------------------------
# Synthetic inputs (SURVEY.md §8d "Value distributions")
# ---------------------------------------------------------------------------------------------

def make_waveforms(seed: int, batch: int, n_samples: int = 160000, kind: str = "noise") -> np.ndarray:
    """``noise``: uniform ±0.1 white noise; ``chirp``: 0.3·sin sweep 100–7000 Hz with a silent tail
    (exercises the log(1e-5) floor of the mel path)."""
    rng = np.random.default_rng(np.random.PCG64(7919 * seed + 1234))
    if kind == "noise":
        return (0.1 * (2.0 * rng.random((batch, n_samples), dtype=np.float32) - 1.0)).astype(np.float32)
    t = np.arange(n_samples, dtype=np.float64) / 16000.0
    out = np.zeros((batch, n_samples), dtype=np.float32)
    for b in range(batch):
        f0 = 100.0 + 300.0 * rng.random()
        f1 = 3000.0 + 4000.0 * rng.random()
        dur = n_samples / 16000.0
        phase = 2 * np.pi * (f0 * t + 0.5 * (f1 - f0) / dur * t * t)
        w = 0.3 * np.sin(phase)
        tail = int(n_samples * (0.6 + 0.3 * rng.random()))
        w[tail:] = 0.0
        out[b] = w.astype(np.float32)
    return out


def make_captions(seed: int, batch: int, max_len: int = 32, lens=None,
                  vocab: int = VOCAB) -> Tuple[np.ndarray, np.ndarray]:
    """Synthetic RoBERTa ids: ``<s>``=0 first, ``</s>``=2 last valid, pad=1, body uniform in [3,vocab).
    Returns (ids int64 [B,T], mask int64 [B,T])."""
    rng = np.random.default_rng(np.random.PCG64(104729 * seed + 99))
    ids = np.full((batch, max_len), 1, dtype=np.int64)
    mask = np.zeros((batch, max_len), dtype=np.int64)
    for b in range(batch):
        n = max_len if lens is None else int(lens[b % len(lens)])
        n = max(2, min(max_len, n))
        body = rng.integers(3, vocab, size=n, dtype=np.int64)
        body[0] = 0
        body[n - 1] = 2
        ids[b, :n] = body
        mask[b, :n] = 1
    return ids, mask
