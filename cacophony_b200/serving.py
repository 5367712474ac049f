"""Low-latency replay of the hot path for a fixed request shape: the whole step — frontend, both towers (text on its side
stream), L2 normalisation and the similarity matrix, ~180 kernel launches — is captured once into a CUDA graph and
replayed with one launch.  At small batches the step is launch-bound (one pair: 183 launches in 1.26 ms), which is what
a graph removes; at the bench batch (256 pairs) the step is power-bound and a graph changes nothing.

The captured kernels read and write fixed device buffers (tensor maps and pointers are baked into the graph), so requests
are copied into static input buffers and results are returned as views of static output buffers (valid until the next
replay).  The graph shares the model handle's workspaces with eager calls on the same model: replays and eager calls must be
issued on the same stream (or otherwise ordered), as any two calls on one model must.  The reference has no counterpart (it issues one eager model call per clip, eval_caco_torch.py:315-336).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .model import CACO, DecodeCache


class GraphedPairs:
    """encode_pairs + similarity for a fixed (batch, n_samples, text_len) as one CUDA graph."""

    def __init__(self, model: CACO, batch: int, n_samples: int, text_len: int, max_patches: int = 500, warmup: int = 3):
        dev = model._device()
        if dev.type != "cuda":
            raise RuntimeError("cacophony_b200 runs on a CUDA device only (no CPU fallback)")
        self.model, self.max_patches = model, max_patches
        self.wave = torch.zeros((batch, n_samples), dtype=torch.float32, device=dev)
        self.ids = torch.ones((batch, text_len), dtype=torch.int64, device=dev)
        self.mask = torch.ones((batch, text_len), dtype=torch.float32, device=dev)
        self.ids[:, 0] = 0
        self._warmup = max(1, warmup)
        self._capture()

    def _capture(self) -> None:
        """(Re)capture.  The graph bakes in raw pointers into the model handle's workspaces and packed weights; the handle
        counts every release of such memory (workspace growth by a larger eager call, re-pack after load_state_dict / .to() /
        an option change) in ``generation``, and a stale graph is re-captured before it is replayed — never replayed."""
        dev = self.model._device()
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                       # warm-up off the capture: workspaces, function attributes
            for _ in range(self._warmup):
                self._step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        with torch.cuda.graph(self.graph):
            self.audio_emb, self.text_emb, self.at, self.ta = self._step()
        self._generation = self.model.generation()
        self.recaptures = getattr(self, "recaptures", -1) + 1

    def _step(self):
        a, t = self.model.encode_pairs(self.wave, self.ids, self.mask, max_patches=self.max_patches)
        at, ta = self.model.similarity(a, t)
        return a, t, at, ta

    @torch.no_grad()
    def __call__(self, waveform: torch.Tensor, text_input_ids: torch.Tensor, text_mask: torch.Tensor
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
        """(audio -> text logits, text -> audio logits) for one request of the captured shape (views of static buffers)."""
        if tuple(waveform.shape) != tuple(self.wave.shape) or tuple(text_input_ids.shape) != tuple(self.ids.shape):
            raise ValueError(f"GraphedPairs was captured for waveform {tuple(self.wave.shape)} / ids {tuple(self.ids.shape)}")
        if self.model._packed_key is None or self.model.generation() != self._generation:
            self._capture()                                 # the handle released memory the old graph points into
        self.wave.copy_(waveform, non_blocking=True)
        self.ids.copy_(text_input_ids, non_blocking=True)
        self.mask.copy_(text_mask, non_blocking=True)
        self.graph.replay()
        return self.at, self.ta


class GraphedDecodeStep:
    """One KV-cached captioning decode step (``CACO.decode_step``: one token per sequence through 12 text layers, 4 decoder layers
    and the vocabulary projection, ~140 launches) as one CUDA graph for a fixed (batch, audio tokens, capacity).  Token ids and
    positions are read from static device buffers, so the same graph serves every step of every request; ``begin`` resets the
    (static) cache for a new batch of clips outside the graph.  Same re-capture rule as ``GraphedPairs``."""

    def __init__(self, model: CACO, batch: int, seq: int, capacity: int, warmup: int = 2):
        dev = model._device()
        if dev.type != "cuda":
            raise RuntimeError("cacophony_b200 runs on a CUDA device only (no CPU fallback)")
        self.model, self.batch, self.seq, self.capacity = model, batch, seq, capacity
        D = model.audio_config.hidden_size
        self.ids = torch.zeros(batch, dtype=torch.int64, device=dev)
        self.pos = torch.zeros(batch, dtype=torch.int64, device=dev)
        self.cache: DecodeCache = model.decode_begin(torch.zeros((batch, seq, D), dtype=torch.float32, device=dev),
                                                     torch.ones((batch, seq), dtype=torch.float32, device=dev), capacity)
        self.logits = torch.empty((batch, self.cache.vocab), dtype=torch.float32, device=dev)
        self.next = torch.empty(batch, dtype=torch.int32, device=dev)
        self._warmup = max(1, warmup)
        self._capture()

    def _run(self) -> None:
        self.model.decode_step(self.cache, self.ids, self.pos, logits_out=self.logits, next_out=self.next)

    def _capture(self) -> None:
        dev = self.model._device()
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                       # warm-up off the capture: workspace, function attributes
            for _ in range(self._warmup):
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        with torch.cuda.graph(self.graph):
            self._run()
        self._generation = self.model.generation()
        self.recaptures = getattr(self, "recaptures", -1) + 1

    @torch.no_grad()
    def begin(self, audio_hidden_state: torch.Tensor, audio_mask: torch.Tensor) -> None:
        """New batch of clips: empty the cache, store their cross-attention keys / values (eager, once per request)."""
        self.model.decode_begin(audio_hidden_state, audio_mask, self.capacity, cache=self.cache)

    @torch.no_grad()
    def step(self, token_ids: torch.Tensor, positions: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """(next-token logits [batch, vocab], their arg-max [batch] int32) — views of static buffers, valid until the next step."""
        if self.model._packed_key is None or self.model.generation() != self._generation:
            self._capture()                                 # the handle released memory the old graph points into
        self.ids.copy_(token_ids.reshape(-1), non_blocking=True)
        self.pos.copy_(positions.reshape(-1), non_blocking=True)
        self.graph.replay()
        return self.logits, self.next
