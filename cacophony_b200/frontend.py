"""Audio frontend with the reference's function names (src/eval/eval_caco_torch.py:30-38, :41-151, :181-206),
executed by the K1 CUDA kernel (csrc/frontend.cu).  The reference handles one clip per call and round-trips through
numpy; here a [batch, n_samples] tensor is processed in one launch and everything stays on the device.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Union

import numpy as np
import torch

from . import ops


@dataclass
class DatasetConfig:                   # eval_caco_torch.py:30-38
    batch_size: int = 1
    patches_seq_len: int = 512
    time_patch_size: int = 16
    freq_patch_size: int = 16
    max_text_len: int = 100
    synthetic_prob: float = 0.8


def _wave(audio, device) -> torch.Tensor:
    w = torch.as_tensor(audio)
    if w.dim() == 1:
        w = w[None]
    if w.dim() != 2:
        raise ValueError("audio: expected [n_samples] or [batch, n_samples]")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("cacophony_b200 frontend runs on a CUDA device only (no CPU fallback)")
    return w.to(device=dev, dtype=torch.float32).contiguous()


def compute_mel_spectrogram(waveform, sample_rate: int = 16000, window_size: int = 400, hop_size: int = 160,
                            n_fft: int = 512, n_mels: int = 128, device: Union[str, torch.device] = "cuda") -> np.ndarray:
    """eval_caco_torch.py:41-105 -> log-mel [frames, 128] (numpy, like the reference) for ONE clip [1, L]."""
    if (sample_rate, window_size, hop_size, n_fft, n_mels) != (16000, 400, 160, 512, 128):
        raise ValueError("the CUDA frontend is specialised to 16 kHz / 400 / 160 / 512 / 128 (the checkpoint's settings)")
    w = _wave(waveform, device)
    if w.shape[0] != 1:
        raise ValueError("compute_mel_spectrogram takes one clip; use prepare_audio_batch for batches")
    with torch.cuda.device(w.device):
        out = ops.frontend(w, max_patches=8, want_log_mel=True)
    return out["log_mel"][0].cpu().numpy()


def spectrogram_to_patches(spectrogram, time_patch_size: int = 16, freq_patch_size: int = 16,
                           max_patches: int = 512, device: Union[str, torch.device] = "cuda") -> Dict[str, torch.Tensor]:
    """eval_caco_torch.py:108-151 on an existing log-mel [frames, 128]: pure index shuffling (no arithmetic), done with
    torch views on the device.  Returns device tensors (the reference returns numpy)."""
    if (time_patch_size, freq_patch_size) != (16, 16):
        raise ValueError("16x16 patches only")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("cacophony_b200 runs on a CUDA device only")
    mel = torch.as_tensor(spectrogram, dtype=torch.float32, device=dev)
    nt, nf = mel.shape[0] // 16, mel.shape[1] // 16
    full = nt * nf
    x = mel[: nt * 16].reshape(nt, 16, nf, 16).permute(0, 2, 1, 3).reshape(full, 256)
    p = torch.arange(max_patches, device=dev)
    if full > max_patches:
        x = x[:max_patches]
        mask = torch.ones(max_patches, device=dev)
        live = p
    else:
        mask = (p < full).float()
        live = (mask * p).long()
        x = torch.cat([x, torch.zeros(max_patches - full, 256, device=dev)], 0)
    return {"audio_patches": x.contiguous(), "audio_time_inds": (live // nf).float(), "audio_freq_inds": (live % nf).float(),
            "audio_mask": mask}


def prepare_audio_batch(audio, datasetconfig: DatasetConfig = None, device: Union[str, torch.device] = "cuda"
                        ) -> Dict[str, torch.Tensor]:
    """eval_caco_torch.py:181-206, batched: audio [batch, n_samples] (or [n_samples]) ->
    audio_patches [batch, P, 256], audio_time_inds / audio_freq_inds / audio_mask [batch, P], all float32 on `device`."""
    cfg = datasetconfig or DatasetConfig()
    w = _wave(audio, device)
    with torch.cuda.device(w.device):
        return ops.frontend(w, max_patches=cfg.patches_seq_len)
