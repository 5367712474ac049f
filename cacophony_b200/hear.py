"""Row (f-4): HEAR-style embedding API (src/eval/heareval/embeddings/audio_embedding/caco_embeddings.py:41-131).

The reference wrapper is JAX/TF-only (``load_caco`` + ``jax.pmap``); its two products are restated on this path:
  scene embedding      = the L2-normalised audio embedding                          (caco_embeddings.py:130-131)
  timestamp embeddings = tf.nn.avg_pool(hidden_state, ksize=8, strides=8, 'VALID')  (caco_embeddings.py:124-125)
                         i.e. the mean over the 8 frequency patches of each 160 ms patch row, with
                         timestamps = linspace(0, audio_max_len * 1000, n_rows) ms   (caco_embeddings.py:127)
Audio goes through this package's frontend (the torch evaluation frontend, eval_caco_torch.py:41-151), not the TF STFT of
``compute_mel_spec_audiomae``; the patch budget is the reference's ``max_patches`` formula (caco_embeddings.py:72-73).
"""
from __future__ import annotations

from typing import Any, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import loader, ops
from .model import CACO


class Embedding:
    """Mirror of caco_embeddings.Embedding for an already-constructed CACO model (checkpoint loading: eval.load_caco_torch)."""

    def __init__(self, model: CACO, audio_max_len: float = 10, batch_size: int = 1, sample_rate: int = 16000):
        if sample_rate != 16000:
            raise ValueError("the CUDA frontend is specialised to 16 kHz input (resample with loader.resample_to_16k)")
        self.model = model
        self.audio_max_len = audio_max_len
        self.batch_size = batch_size
        self.sample_rate = sample_rate
        segment = int(audio_max_len * sample_rate)
        self.max_patches = (segment // 160 // 16) * (128 // 16)                # caco_embeddings.py:72-73

    @torch.no_grad()
    def embed(self, waves: Sequence[Any]) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """(scene [B, 768], hidden [B, P, 768], mask [B, P]) for a list of clips (device tensors)."""
        buf, lens = loader.pad_ragged(waves, stride=int(self.audio_max_len * self.sample_rate))
        w = buf.to(self.model._device(), non_blocking=True)
        return self.model.encode_audio(w, max_patches=self.max_patches, lengths=lens, return_hidden_state=True)

    @torch.no_grad()
    def get_scene_embeddings(self, waves: Sequence[Any]) -> torch.Tensor:
        return self.embed(waves)[0]

    @torch.no_grad()
    def get_timestamp_embeddings(self, waves: Sequence[Any]) -> Tuple[torch.Tensor, np.ndarray]:
        """([B, P // 8, 768] patch-row embeddings, timestamps in ms [P // 8])."""
        _, hid, _ = self.embed(waves)
        pooled = ops.avg_pool_tokens(hid, 8)
        return pooled, np.linspace(0, self.audio_max_len * 1000, pooled.shape[-2])

    def get_embedding_as_numpy(self, audio: Any, embedding_type: Optional[str] = None):
        """caco_embeddings.py:99-131 for one clip (array of 16 kHz samples; the reference takes a file name)."""
        if embedding_type == "event":
            pooled, ts = self.get_timestamp_embeddings([audio])
            return pooled[0].cpu().numpy(), [ts]
        return self.get_scene_embeddings([audio])[0].cpu().numpy()
