"""Host-side mirror of the reference model API (src/caco_torch/caco.py, audio_models/mae.py, text_models/roberta.py).

Same class names, dataclass fields, method names, keyword arguments, defaults, return arity and ``state_dict``
keys as the reference, so ``from cacophony_b200 import create_caco_model`` replaces
``from src.caco_torch import create_caco_model`` for the inference path.  The modules below are parameter
containers only: all arithmetic happens in libcaco_b200.so (hand-written sm_100a kernels) behind a C handle
(``caco_model_*`` in include/caco_b200.h).  There is no torch-math forward and no CPU fallback: calling a
model whose parameters are not on a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional, Tuple, Union

import torch
import torch.nn as nn

from . import _lib as L
from . import ops

NORM_EPS = 1e-10                       # caco.py:14


@dataclass
class AudioTransformerConfig:          # mae.py:9-20
    hidden_size: int
    num_layers: int
    num_heads: int
    intermediate_size: int
    patch_size: int
    max_time_ind: int
    num_freq_patches: int
    dropout_rate: float
    drop_path_rate: float
    dtype: torch.dtype = torch.float32


@dataclass
class RobertaConfig:                   # roberta.py:11-23
    vocab_size: int = 50265
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    hidden_dropout_prob: float = 0.1
    attention_probs_dropout_prob: float = 0.1
    max_position_embeddings: int = 514
    type_vocab_size: int = 1
    layer_norm_eps: float = 1e-5
    pad_token_id: int = 1


@dataclass
class CACOConfig:                      # caco.py:17-21
    projection_size: int = 768
    num_attention_pool_heads: int = 2
    logit_scale_init_value: float = 2.6592


# ------------------------------------------------------------------------------------------------------
# parameter containers (same attribute tree => same state_dict keys as the reference)
# ------------------------------------------------------------------------------------------------------
class _Linear(nn.Module):
    def __init__(self, fan_in: int, fan_out: int):
        super().__init__()
        k = 1.0 / fan_in ** 0.5
        self.weight = nn.Parameter(torch.empty(fan_out, fan_in).uniform_(-k, k), requires_grad=False)
        self.bias = nn.Parameter(torch.empty(fan_out).uniform_(-k, k), requires_grad=False)


class _LayerNorm(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim), requires_grad=False)
        self.bias = nn.Parameter(torch.zeros(dim), requires_grad=False)


class _MHA(nn.Module):                 # nn.MultiheadAttention's parameter names (mae.py:69-74)
    def __init__(self, dim: int):
        super().__init__()
        k = (6.0 / (4 * dim)) ** 0.5
        self.in_proj_weight = nn.Parameter(torch.empty(3 * dim, dim).uniform_(-k, k), requires_grad=False)
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * dim), requires_grad=False)
        self.out_proj = _Linear(dim, dim)


class _MLP(nn.Module):                 # mae.py:47-61
    def __init__(self, dim: int, ffn: int):
        super().__init__()
        self.fc1 = _Linear(dim, ffn)
        self.fc2 = _Linear(ffn, dim)


class AudioEncoderLayer(nn.Module):    # mae.py:64-99
    def __init__(self, cfg: AudioTransformerConfig):
        super().__init__()
        self.norm1 = _LayerNorm(cfg.hidden_size)
        self.attn = _MHA(cfg.hidden_size)
        self.norm2 = _LayerNorm(cfg.hidden_size)
        self.mlp = _MLP(cfg.hidden_size, cfg.intermediate_size)


class AudioEncoder(nn.Module):
    """Parameter tree of mae.py:112-148.  Executed by CACO.get_audio_embedding through the C handle."""

    def __init__(self, config: AudioTransformerConfig):
        super().__init__()
        self.config = config
        self.input_proj = _Linear(config.patch_size, config.hidden_size)
        self.freq_positional_embedding = nn.Parameter(torch.randn(config.num_freq_patches, config.hidden_size) * 0.02,
                                                      requires_grad=False)
        self.layers = nn.ModuleList([AudioEncoderLayer(config) for _ in range(config.num_layers)])
        self.norm = _LayerNorm(config.hidden_size)


class AudioAttentionPooler(nn.Module):
    """Parameter tree of caco.py:24-79."""

    def __init__(self, hidden_size: int, num_heads: int, projection_size: Optional[int] = None):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = hidden_size // num_heads
        self.kv_proj = _Linear(hidden_size, 2 * hidden_size)
        self.out_proj = _Linear(hidden_size, projection_size or hidden_size)
        self.query = nn.Parameter(torch.randn(hidden_size) * 0.02, requires_grad=False)


class _Embedding(nn.Module):
    def __init__(self, n: int, dim: int):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(n, dim), requires_grad=False)


class RobertaEmbeddings(nn.Module):    # roberta.py:26-53
    def __init__(self, cfg: RobertaConfig):
        super().__init__()
        self.word_embeddings = _Embedding(cfg.vocab_size, cfg.hidden_size)
        self.position_embeddings = _Embedding(cfg.max_position_embeddings, cfg.hidden_size)
        self.token_type_embeddings = _Embedding(cfg.type_vocab_size, cfg.hidden_size)
        self.LayerNorm = _LayerNorm(cfg.hidden_size)


class _SelfAttention(nn.Module):       # roberta.py:56-104
    def __init__(self, cfg: RobertaConfig):
        super().__init__()
        self.query = _Linear(cfg.hidden_size, cfg.hidden_size)
        self.key = _Linear(cfg.hidden_size, cfg.hidden_size)
        self.value = _Linear(cfg.hidden_size, cfg.hidden_size)


class _DenseLN(nn.Module):             # roberta.py:107-124 / :161-178
    def __init__(self, fan_in: int, dim: int):
        super().__init__()
        self.dense = _Linear(fan_in, dim)
        self.LayerNorm = _LayerNorm(dim)


class _Dense(nn.Module):               # roberta.py:150-158
    def __init__(self, fan_in: int, fan_out: int):
        super().__init__()
        self.dense = _Linear(fan_in, fan_out)


class _Attention(nn.Module):
    def __init__(self, cfg: RobertaConfig):
        super().__init__()
        self.self = _SelfAttention(cfg)
        self.output = _DenseLN(cfg.hidden_size, cfg.hidden_size)


class RobertaLayer(nn.Module):         # roberta.py:181-215 (self-attention branch only)
    def __init__(self, cfg: RobertaConfig):
        super().__init__()
        self.attention = _Attention(cfg)
        self.intermediate = _Dense(cfg.hidden_size, cfg.intermediate_size)
        self.output = _DenseLN(cfg.intermediate_size, cfg.hidden_size)


class RobertaEncoder(nn.Module):       # roberta.py:218-242
    def __init__(self, cfg: RobertaConfig):
        super().__init__()
        self.layers = nn.ModuleList([RobertaLayer(cfg) for _ in range(cfg.num_hidden_layers)])


class AttentionPooler(nn.Module):      # roberta.py:245-271
    def __init__(self, cfg: RobertaConfig):
        super().__init__()
        self.attention_pool_query = nn.Parameter(torch.randn(1, cfg.hidden_size) * 0.02, requires_grad=False)
        self.key_proj = _Linear(cfg.hidden_size, cfg.hidden_size)
        self.value_proj = _Linear(cfg.hidden_size, cfg.hidden_size)


class RobertaModel(nn.Module):
    """Parameter tree of roberta.py:274-326.  Executed by CACO.get_text_embedding through the C handle."""

    def __init__(self, config: RobertaConfig):
        super().__init__()
        self.config = config
        self.embeddings = RobertaEmbeddings(config)
        self.encoder = RobertaEncoder(config)
        self.pooler = AttentionPooler(config)


class _DecoderLayer(RobertaLayer):     # roberta.py:181-215 with the cross-attention branch (:185-187)
    def __init__(self, cfg: RobertaConfig):
        super().__init__(cfg)
        self.crossattention = _Attention(cfg)


class _DecoderEncoder(nn.Module):      # RobertaEncoder(config, has_cross_attention=True), roberta.py:218-224
    def __init__(self, cfg: RobertaConfig):
        super().__init__()
        self.layers = nn.ModuleList([_DecoderLayer(cfg) for _ in range(cfg.num_hidden_layers)])


class RobertaDecoder(nn.Module):
    """Parameter tree of the captioning head (roberta.py:329-373: layers with cross-attention + ``decoder_proj``), same
    ``state_dict`` keys as the reference's, exported so ``from cacophony_b200 import *`` offers the reference's names and a
    checkpoint's ``decoder_module.*`` tensors have somewhere to live.  Captioning is not on the inference hot path
    (SURVEY.md 8, row f-4): there is no forward."""

    def __init__(self, config: RobertaConfig):
        super().__init__()
        self.config = config
        self.encoder = _DecoderEncoder(config)
        self.decoder_proj = _Linear(config.hidden_size, config.vocab_size)

    def forward(self, *args, **kwargs):
        raise NotImplementedError("cacophony_b200 implements the contrastive inference path; the captioning decoder is a "
                                  "parameter container only")


# ------------------------------------------------------------------------------------------------------
# CACO
# ------------------------------------------------------------------------------------------------------
class CACO(nn.Module):
    """Drop-in for src/caco_torch/caco.py:82-261 (inference path, including ``get_decoder_logits`` of the captioning head when
    the model was built with a ``decoder_config``, as ``create_caco_model`` does)."""

    def __init__(self, audio_config: AudioTransformerConfig, text_config: RobertaConfig, caco_config: CACOConfig,
                 decoder_config: Optional[RobertaConfig] = None):
        super().__init__()
        self.audio_config, self.text_config, self.caco_config = audio_config, text_config, caco_config
        if caco_config.projection_size != audio_config.hidden_size or text_config.hidden_size != audio_config.hidden_size:
            raise ValueError("cacophony_b200 supports projection_size == hidden_size for both towers (the checkpoint's shape)")
        self.audio_module = AudioEncoder(audio_config)
        self.audio_attention_pool = AudioAttentionPooler(audio_config.hidden_size, caco_config.num_attention_pool_heads,
                                                         caco_config.projection_size)
        self.text_module = RobertaModel(text_config)
        self.text_proj = _Linear(text_config.hidden_size, caco_config.projection_size)
        self.logit_scale = nn.Parameter(torch.tensor(caco_config.logit_scale_init_value), requires_grad=False)
        # caco.py:118-121: the captioning head exists iff a decoder config is given (create_caco_model gives one)
        self.decoder_module = RobertaDecoder(decoder_config) if decoder_config is not None else None
        self._handle: Optional[int] = None
        self._packed_key = None
        self._side_stream = None
        self._options: Dict[str, int] = {}
        self.eval()

    # ---- state handling ---------------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        """Accepts the reference checkpoint layouts' tensors.  ``decoder_module.*`` tensors are loaded when this
        model has the captioning head and dropped otherwise; an encoder-only state_dict leaves the head as it is."""
        has_dec = any(k.startswith("decoder_module.") for k in state_dict)
        if self.decoder_module is None:
            sd = {k: v for k, v in state_dict.items() if not k.startswith("decoder_module.")}
        elif not has_dec:
            sd = dict(state_dict)
            sd.update({"decoder_module." + k: v for k, v in self.decoder_module.state_dict().items()})
        else:
            sd = state_dict
        out = super().load_state_dict(sd, strict=strict, assign=assign)
        self._packed_key = None
        return out

    def _apply(self, fn, *a, **kw):      # .to() / .cuda() move the parameters -> re-pack lazily
        self._packed_key = None
        return super()._apply(fn, *a, **kw)

    def __del__(self):
        try:
            if self._handle is not None:
                L.load().caco_model_destroy(self._handle)
        except Exception:
            pass

    # The C handle (packed weights, workspaces) belongs to ONE Python object: copies and unpickled models start without one
    # and build their own at first use — never two objects destroying the same handle, never a copy running on the
    # original's packed weights.
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_handle"], state["_packed_key"], state["_side_stream"] = None, None, None
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._handle, self._packed_key, self._side_stream = None, None, None

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k in ("_handle", "_packed_key", "_side_stream") else copy.deepcopy(v, memo)
        return new

    # ---- execution options (per model; see caco_set_default_option in include/caco_b200.h for the names) -------------
    def set_option(self, name: str, value: int) -> None:
        """E.g. ``set_option("split_weights", 1)``: GEMM weights as fp16 hi + lo (two accumulating tensor-core passes, the
        precision escape hatch of SURVEY.md 7.3), ``"audio_chunk_rows"``, ``"pdl"``."""
        self._options[name] = int(value)
        if self._handle is not None:
            L.check(L.load().caco_model_set_option(self._handle, name.encode(), int(value)), f"caco_model_set_option({name})")

    def generation(self) -> int:
        """Changes whenever the handle released device memory an earlier enqueued / captured call may reference."""
        return 0 if self._handle is None else int(L.load().caco_model_generation(self._handle))

    def _device(self) -> torch.device:
        return self.logit_scale.device

    def _ensure_packed(self) -> int:
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("cacophony_b200.CACO runs on a CUDA device only (no CPU fallback): call model.to('cuda')")
        if self._handle is not None and self._packed_key:
            return self._handle          # invalidated by load_state_dict / .to(); call repack() after in-place edits
        sd = self.state_dict()
        lib = L.load()
        if self._handle is None:
            a, t = self.audio_config, self.text_config
            cfg = L.CacoConfig(a.hidden_size, a.intermediate_size, a.patch_size, a.num_layers, a.num_heads,
                               a.num_freq_patches, self.caco_config.num_attention_pool_heads, t.num_hidden_layers,
                               t.num_attention_heads, t.vocab_size, t.max_position_embeddings, float(t.layer_norm_eps),
                               1e-5)        # audio tower: nn.LayerNorm's default eps (mae.py:68,76,123)
            h = C.c_void_p()
            L.check(lib.caco_model_create(C.byref(cfg), C.byref(h)), "caco_model_create")
            self._handle = h.value
            for name, value in self._options.items():
                L.check(lib.caco_model_set_option(self._handle, name.encode(), value), f"caco_model_set_option({name})")
        with torch.cuda.device(dev):
            for k, v in sd.items():
                if v.dtype != torch.float32 or not v.is_contiguous():
                    raise ValueError(f"parameter {k}: expected contiguous float32")
                rc = lib.caco_model_set_tensor(self._handle, k.encode(), v.data_ptr(), v.numel())
                if rc < 0:
                    L.check(rc, f"caco_model_set_tensor({k})")
            L.check(lib.caco_model_pack(self._handle, L.stream_ptr()), "caco_model_pack")
        self._packed_key = True
        return self._handle

    def repack(self) -> None:
        """Re-read the parameters (needed only after modifying them in place)."""
        self._packed_key = None
        self._ensure_packed()

    # ---- reference API ----------------------------------------------------------------------------
    @torch.no_grad()
    def get_audio_embedding(self, audio_patches: torch.Tensor, audio_time_inds: torch.Tensor,
                            audio_freq_inds: torch.Tensor, audio_mask: torch.Tensor, deterministic: bool = True,
                            return_hidden_state: bool = True, normalize: bool = False
                            ) -> Union[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]:
        """caco.py:123-150."""
        if not deterministic:
            raise ValueError("cacophony_b200 implements the inference path only (deterministic=True)")
        h = self._ensure_packed()
        dev = self._device()
        p = _as(audio_patches, torch.float32, dev, "audio_patches")
        if p.dim() != 3 or p.shape[-1] != self.audio_config.patch_size:
            raise ValueError(f"audio_patches: expected [batch, seq, {self.audio_config.patch_size}]")
        B, S, _ = p.shape
        ti = _as(audio_time_inds, torch.float32, dev, "audio_time_inds")
        fi = _as(audio_freq_inds, torch.float32, dev, "audio_freq_inds")
        mk = _as(audio_mask, torch.float32, dev, "audio_mask")
        for n, t in (("audio_time_inds", ti), ("audio_freq_inds", fi), ("audio_mask", mk)):
            if tuple(t.shape) != (B, S):
                raise ValueError(f"{n}: expected shape {(B, S)}, got {tuple(t.shape)}")
        D = self.audio_config.hidden_size
        emb = torch.empty((B, D), dtype=torch.float32, device=dev)
        hid = torch.empty((B, S, D), dtype=torch.float32, device=dev) if return_hidden_state else None
        with torch.cuda.device(dev):
            L.check(L.load().caco_model_audio_embedding(h, L.ptr(p), L.ptr(ti), L.ptr(fi), L.ptr(mk), B, S, int(normalize),
                                                        L.ptr(emb), L.ptr(hid), L.stream_ptr()), "caco_model_audio_embedding")
        return (emb, hid) if return_hidden_state else emb

    @torch.no_grad()
    def get_text_embedding(self, text_input_ids: torch.Tensor, text_mask: torch.Tensor,
                           position_ids: Optional[torch.Tensor] = None, deterministic: bool = True,
                           return_hidden_state: bool = True, normalize: bool = False
                           ) -> Union[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]:
        """caco.py:152-177."""
        if not deterministic:
            raise ValueError("cacophony_b200 implements the inference path only (deterministic=True)")
        h = self._ensure_packed()
        dev = self._device()
        ids = _as(text_input_ids, torch.int64, dev, "text_input_ids")
        if ids.dim() != 2:
            raise ValueError("text_input_ids: expected [batch, seq]")
        B, T = ids.shape
        if T > self.text_config.max_position_embeddings:
            raise ValueError(f"text sequence length must be <= {self.text_config.max_position_embeddings}")
        mk = _as(text_mask, torch.float32, dev, "text_mask")
        if tuple(mk.shape) != (B, T):
            raise ValueError(f"text_mask: expected shape {(B, T)}")
        pids = None if position_ids is None else _as(position_ids, torch.int64, dev, "position_ids").expand(B, T).contiguous()
        D = self.text_config.hidden_size
        emb = torch.empty((B, D), dtype=torch.float32, device=dev)
        hid = torch.empty((B, T, D), dtype=torch.float32, device=dev) if return_hidden_state else None
        with torch.cuda.device(dev):
            L.check(L.load().caco_model_text_embedding(h, L.ptr(ids), L.ptr(mk), L.ptr(pids), B, T, int(normalize), L.ptr(emb),
                                                       L.ptr(hid), L.stream_ptr()), "caco_model_text_embedding")
        return (emb, hid) if return_hidden_state else emb

    @torch.no_grad()
    def get_contrastive_logits(self, audio_patches, audio_time_inds, audio_freq_inds, audio_mask, text_input_ids,
                               text_mask, deterministic: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
        """caco.py:179-212."""
        a = self.get_audio_embedding(audio_patches, audio_time_inds, audio_freq_inds, audio_mask,
                                     deterministic=deterministic, return_hidden_state=False, normalize=True)
        t = self.get_text_embedding(text_input_ids, text_mask, deterministic=deterministic,
                                    return_hidden_state=False, normalize=True)
        return self.similarity(a, t)

    @torch.no_grad()
    def get_decoder_logits(self, audio_hidden_state: torch.Tensor, audio_mask: torch.Tensor, text_input_ids: torch.Tensor,
                           text_mask: torch.Tensor, deterministic: bool = True) -> torch.Tensor:
        """caco.py:214-240: captioning logits [batch, T, vocab] — text tower hidden state -> RobertaDecoder (causal
        self-attention, cross-attention to ``audio_hidden_state`` [batch, S, hidden], GELU MLP, vocabulary projection)."""
        if self.decoder_module is None:
            raise ValueError("Decoder module not initialized")       # caco.py:223-224
        if not deterministic:
            raise ValueError("cacophony_b200 implements the inference path only (deterministic=True)")
        h = self._ensure_packed()
        dev = self._device()
        _, text_hidden = self.get_text_embedding(text_input_ids, text_mask, deterministic=True, return_hidden_state=True)
        ah = _as(audio_hidden_state, torch.float32, dev, "audio_hidden_state")
        am = _as(audio_mask, torch.float32, dev, "audio_mask")
        tm = _as(text_mask, torch.float32, dev, "text_mask")
        if ah.dim() != 3 or ah.shape[-1] != self.audio_config.hidden_size or tuple(am.shape) != tuple(ah.shape[:2]):
            raise ValueError("audio_hidden_state: expected [batch, seq, hidden] with audio_mask [batch, seq]")
        B, T, _ = text_hidden.shape
        if ah.shape[0] != B:
            raise ValueError("audio and text batch sizes differ")
        lib = L.load()
        V = int(lib.caco_model_decoder_vocab(h))
        if V <= 0:
            raise ValueError("Decoder module not initialized")
        logits = torch.empty((B, T, V), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            L.check(lib.caco_model_decoder_logits(h, L.ptr(text_hidden), L.ptr(tm), L.ptr(ah), L.ptr(am), B, T, int(ah.shape[1]),
                                                  L.ptr(logits), L.stream_ptr()), "caco_model_decoder_logits")
        return logits

    # ---- KV-cached captioning decode -------------------------------------------
    @torch.no_grad()
    def decode_begin(self, audio_hidden_state: torch.Tensor, audio_mask: torch.Tensor, capacity: int,
                     cache: Optional["DecodeCache"] = None) -> "DecodeCache":
        """Start an incremental decode for a batch of clips: allocates the key / value cache for sequences of up to ``capacity``
        tokens and stores the cross-attention keys / values of ``audio_hidden_state`` [batch, S, hidden].  Text tower and
        decoder are causal (roberta.py:297-310, 346-355), so ``decode_step`` on token t against the cache gives the logits
        ``get_decoder_logits(...)[:, t]`` of the full-prefix call the reference makes every step (eval_caco_torch.py:411-472).
        ``cache``: an earlier cache of the same (batch, S, capacity) to reset and reuse instead of allocating a new one."""
        if self.decoder_module is None:
            raise ValueError("Decoder module not initialized")       # caco.py:223-224
        h = self._ensure_packed()
        dev = self._device()
        ah = _as(audio_hidden_state, torch.float32, dev, "audio_hidden_state")
        am = _as(audio_mask, torch.float32, dev, "audio_mask")
        if ah.dim() != 3 or ah.shape[-1] != self.audio_config.hidden_size or tuple(am.shape) != tuple(ah.shape[:2]):
            raise ValueError("audio_hidden_state: expected [batch, seq, hidden] with audio_mask [batch, seq]")
        B, S = int(ah.shape[0]), int(ah.shape[1])
        capacity = int(capacity)
        if not 1 <= capacity <= self.text_config.max_position_embeddings:
            raise ValueError(f"capacity must be in [1, {self.text_config.max_position_embeddings}]")
        lib = L.load()
        nbytes = int(lib.caco_model_decode_cache_bytes(h, B, S, capacity))
        if nbytes <= 0:
            raise ValueError("Decoder module not initialized")
        if cache is not None:
            if (cache.batch, cache.seq, cache.capacity) != (B, S, capacity) or cache.view.numel() != nbytes or cache.view.device != dev:
                raise ValueError("decode_begin: the cache to reuse was made for another shape or device")
        else:
            buf = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
            shift = (-buf.data_ptr()) % 256
            cache = DecodeCache(buf, buf[shift:shift + nbytes], B, S, capacity, int(lib.caco_model_decoder_vocab(h)))
        with torch.cuda.device(dev):
            L.check(lib.caco_model_decode_begin(h, L.ptr(cache.view), nbytes, L.ptr(ah), L.ptr(am), B, S, capacity,
                                                L.stream_ptr()), "caco_model_decode_begin")
        return cache

    @torch.no_grad()
    def decode_step(self, cache: "DecodeCache", token_ids: torch.Tensor, positions: torch.Tensor,
                    logits_out: Optional[torch.Tensor] = None, next_out: Optional[torch.Tensor] = None,
                    want_logits: bool = True, want_next: bool = False):
        """Push one token per sequence (``token_ids`` [batch] int64 at ``positions`` [batch] int64, device tensors) and return
        the next-token logits [batch, vocab] and / or their arg-max [batch] int32.  Positions must arrive in order from 0."""
        h = self._ensure_packed()
        dev = self._device()
        ids = _as(token_ids, torch.int64, dev, "token_ids").reshape(-1)
        pos = _as(positions, torch.int64, dev, "positions").reshape(-1)
        if ids.numel() != cache.batch or pos.numel() != cache.batch:
            raise ValueError(f"token_ids / positions: expected {cache.batch} entries")
        if logits_out is None and want_logits:
            logits_out = torch.empty((cache.batch, cache.vocab), dtype=torch.float32, device=dev)
        if next_out is None and want_next:
            next_out = torch.empty((cache.batch,), dtype=torch.int32, device=dev)
        if logits_out is None and next_out is None:
            raise ValueError("decode_step: ask for logits, the arg-max, or both")
        with torch.cuda.device(dev):
            L.check(L.load().caco_model_decode_step(h, L.ptr(cache.view), L.ptr(ids), L.ptr(pos), cache.batch, cache.seq,
                                                    cache.capacity, L.ptr(logits_out), L.ptr(next_out), L.stream_ptr()),
                    "caco_model_decode_step")
        if logits_out is not None and next_out is not None:
            return logits_out, next_out
        return logits_out if logits_out is not None else next_out

    def forward(self, audio_patches, audio_time_inds, audio_freq_inds, audio_mask, text_input_ids, text_mask,
                deterministic: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
        """caco.py:242-261."""
        return self.get_contrastive_logits(audio_patches, audio_time_inds, audio_freq_inds, audio_mask, text_input_ids,
                                           text_mask, deterministic=deterministic)

    # ---- additions (north-star aliases) -------------------------------------------------------------
    @torch.no_grad()
    def similarity(self, audio_embedding: torch.Tensor, text_embedding: torch.Tensor, want_ta: bool = True
                   ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """(exp(logit_scale)·A)·Tᵀ and (exp(logit_scale)·T)·Aᵀ (caco.py:208-210) for already-normalised embeddings."""
        dev = self._device()
        a = _as(audio_embedding, torch.float32, dev, "audio_embedding")
        t = _as(text_embedding, torch.float32, dev, "text_embedding")
        with torch.cuda.device(dev):
            return ops.sim_logits(a, t, self.logit_scale.data.reshape(1), want_ta=want_ta)

    @torch.no_grad()
    def encode_audio(self, waveform: torch.Tensor, max_patches: int = 500, normalize: bool = True,
                     lengths: Optional[torch.Tensor] = None, return_hidden_state: bool = False,
                     trim_padding: bool = False):
        """waveform [batch, n_samples] (16 kHz fp32) -> L2-normalised audio embeddings [batch, 768]:
        prepare_audio_batch (eval_caco_torch.py:181-206) + get_audio_embedding in one library call.

        lengths [batch] (int): ragged batch — clip b is waveform[b, :lengths[b]], each clip framed / patched / masked as the
        reference would treat it alone.  return_hidden_state: also return (hidden [batch, P, 768], mask [batch, P]).
        trim_padding: run the tower on P = the largest valid-patch count in the batch (rounded up to 8) instead of
        max_patches; masked keys get probability exactly 0 and padded tokens are never pooled, so the embeddings are the
        same — only the padded rows of the hidden state are not produced (5 s clips: 248 instead of 500 tokens)."""
        h = self._ensure_packed()
        dev = self._device()
        w = _as(waveform, torch.float32, dev, "waveform")
        if w.dim() == 1:
            w = w[None]
        B, n = w.shape
        lens = None
        if lengths is not None:
            lens = _as(torch.as_tensor(lengths), torch.int32, dev, "lengths")
            if tuple(lens.shape) != (B,):
                raise ValueError(f"lengths: expected shape {(B,)}")
        P = int(max_patches)
        if trim_padding:
            longest = n if lengths is None else int(min(n, max(0, int(torch.as_tensor(lengths).max()))))
            valid = (((longest + 159) // 160) // 16) * 8          # eval_caco_torch.py:67,116-117
            P = max(8, min(P, valid))
        emb = torch.empty((B, self.audio_config.hidden_size), dtype=torch.float32, device=dev)
        hid = torch.empty((B, P, self.audio_config.hidden_size), dtype=torch.float32, device=dev) if return_hidden_state else None
        mk = torch.empty((B, P), dtype=torch.float32, device=dev) if return_hidden_state else None
        with torch.cuda.device(dev):
            L.check(L.load().caco_model_encode_audio_ex(h, L.ptr(w), L.ptr(lens), B, n, P, int(normalize), L.ptr(emb),
                                                        L.ptr(hid), L.ptr(mk), L.stream_ptr()), "caco_model_encode_audio_ex")
        return (emb, hid, mk) if return_hidden_state else emb

    def side_stream(self) -> "torch.cuda.Stream":
        """The stream the text tower runs on next to the audio tower (created on first use, follows the model's device)."""
        dev = self._device()
        if getattr(self, "_side_stream", None) is None or self._side_stream.device != dev:
            self._side_stream = torch.cuda.Stream(device=dev)
        return self._side_stream

    @torch.no_grad()
    def encode_text(self, text_input_ids: torch.Tensor, text_mask: torch.Tensor, normalize: bool = True) -> torch.Tensor:
        return self.get_text_embedding(text_input_ids, text_mask, return_hidden_state=False, normalize=normalize)

    @torch.no_grad()
    def encode_pairs(self, waveform: torch.Tensor, text_input_ids: torch.Tensor, text_mask: torch.Tensor,
                     max_patches: int = 500) -> Tuple[torch.Tensor, torch.Tensor]:
        """Both towers of a batch of pairs, L2-normalised.  The text tower (5 % of the FLOPs, small GEMMs) is enqueued on
        a side stream so its kernels fill the tails of the audio tower's waves; the towers share no buffers (separate
        workspaces in the C handle).  Returns (audio_embeddings, text_embeddings)."""
        dev = self._device()
        self._ensure_packed()
        self.side_stream()
        cur = torch.cuda.current_stream(dev)
        ids = _as(text_input_ids, torch.int64, dev, "text_input_ids")
        mk = _as(text_mask, torch.float32, dev, "text_mask")
        self._side_stream.wait_stream(cur)
        with torch.cuda.stream(self._side_stream):
            t = self.encode_text(ids, mk)
        a = self.encode_audio(waveform, max_patches=max_patches)
        cur.wait_stream(self._side_stream)
        for x in (ids, mk, t):
            x.record_stream(cur)
        return a, t


class DecodeCache:
    """Key / value cache of one incremental captioning decode (``CACO.decode_begin``): a caller-owned device buffer laid out by
    the library (caco_model_decode_cache_bytes), plus the shape it was made for."""

    def __init__(self, storage: torch.Tensor, view: torch.Tensor, batch: int, seq: int, capacity: int, vocab: int):
        self.storage, self.view = storage, view          # view: the 256-byte aligned window the library uses
        self.batch, self.seq, self.capacity, self.vocab = batch, seq, capacity, vocab


def _as(t: torch.Tensor, dtype, dev, name: str) -> torch.Tensor:
    """Move/cast an argument the way the reference's torch ops would accept it (e.g. int64 masks), contiguous."""
    if not isinstance(t, torch.Tensor):
        raise ValueError(f"{name}: expected a torch.Tensor")
    if t.device != dev:
        t = t.to(dev)
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def create_caco_model(num_attention_pool_heads: int = 2) -> CACO:
    """caco.py:264-317: the checkpoint's configuration (num_attention_pool_heads: 2 as the torch port builds it, caco.py:292;
    the JAX loader uses 8, load_model.py:47 — same parameter shapes)."""
    audio_config = AudioTransformerConfig(hidden_size=768, num_layers=12, num_heads=8, intermediate_size=3072,
                                          patch_size=256, max_time_ind=512, num_freq_patches=8, dropout_rate=0.0,
                                          drop_path_rate=0.0)
    text_config = RobertaConfig()
    caco_config = CACOConfig(num_attention_pool_heads=num_attention_pool_heads)
    decoder_config = RobertaConfig(num_hidden_layers=4)
    return CACO(audio_config=audio_config, text_config=text_config, caco_config=caco_config, decoder_config=decoder_config)
