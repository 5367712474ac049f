"""Row (f-1): the step right before the hot path — getting ragged clips onto the GPU in one batch.

The reference handles one file at a time: ``load_audio`` (src/eval/eval_utils.py:6-16: soundfile read, channel mean,
``scipy.signal.resample`` to 16 kHz) -> numpy -> ``prepare_audio_batch`` (eval_caco_torch.py:181-206) with batch dimension 1.
Here clips of different lengths are packed into ONE pinned, zero-padded ``[batch, stride]`` host buffer plus a ``lengths``
vector, copied to the device asynchronously, and framed / patched / masked per clip by ``caco_frontend_ragged`` in a
single launch (each clip gets exactly the treatment the reference gives it alone).

``resample_to_16k`` restates ``scipy.signal.resample``'s Fourier method for real input (scipy 1.x, ``signal/_signaltools.py``)
with ``torch.fft`` on the device, i.e. cuFFT — a LIBRARY call, like cuBLAS, used for an off-hot-path convenience; clip
lengths here have arbitrary prime factors (441 000 -> 160 000), which is what a general FFT library is for.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import ops
from .frontend import DatasetConfig

ArrayLike = Union[np.ndarray, torch.Tensor, Sequence[float]]


def pad_ragged(waves: Sequence[ArrayLike], stride: Optional[int] = None, pin: bool = True,
               multiple: int = 160) -> Tuple[torch.Tensor, torch.Tensor]:
    """Pack 1-D clips into a zero-padded fp32 host tensor [batch, stride] (pinned when CUDA is available) + int32 lengths.
    stride defaults to the longest clip rounded up to `multiple` samples; longer clips are cut to `stride`."""
    if len(waves) == 0:
        raise ValueError("pad_ragged: empty batch")
    arrs = []
    for i, w in enumerate(waves):
        a = w.detach().cpu().numpy() if isinstance(w, torch.Tensor) else np.asarray(w)
        if a.ndim == 2:                      # [samples, channels] as soundfile returns: channel mean (eval_utils.py:9-10)
            a = a.astype(np.float32).mean(axis=-1)
        if a.ndim != 1:
            raise ValueError(f"pad_ragged: clip {i} must be 1-D (or [samples, channels])")
        arrs.append(np.ascontiguousarray(a, dtype=np.float32))
    longest = max(a.shape[0] for a in arrs)
    if stride is None:
        stride = max(multiple, -(-longest // multiple) * multiple)
    buf = torch.zeros((len(arrs), stride), dtype=torch.float32)
    if pin and torch.cuda.is_available():
        buf = buf.pin_memory()
    lens = torch.empty((len(arrs),), dtype=torch.int32)
    view = buf.numpy()
    for i, a in enumerate(arrs):
        n = min(a.shape[0], stride)
        view[i, :n] = a[:n]
        lens[i] = n
    return buf, lens


def valid_patch_counts(lengths: ArrayLike, max_patches: int) -> np.ndarray:
    """Per-clip number of valid tokens, eval_caco_torch.py:67,116-117,132-138: floor(ceil(L/160)/16)*8 cut to max_patches."""
    L = np.asarray(lengths, dtype=np.int64)
    return np.minimum((((L + 159) // 160) // 16) * 8, max_patches)


def prepare_audio_batch_ragged(waves: Sequence[ArrayLike], datasetconfig: Optional[DatasetConfig] = None,
                               device: Union[str, torch.device] = "cuda") -> Dict[str, torch.Tensor]:
    """Batched, ragged ``prepare_audio_batch``: list of clips -> audio_patches [B, P, 256], audio_time_inds /
    audio_freq_inds / audio_mask [B, P] (float32, on `device`), P = datasetconfig.patches_seq_len."""
    cfg = datasetconfig or DatasetConfig()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("cacophony_b200 runs on a CUDA device only (no CPU fallback)")
    buf, lens = pad_ragged(waves)
    with torch.cuda.device(dev):
        w = buf.to(dev, non_blocking=True)
        ln = lens.to(dev, non_blocking=True)
        out = ops.frontend_ragged(w, ln, cfg.patches_seq_len)
    out["lengths"] = ln
    return out


@torch.no_grad()
def resample_to_16k(audio: ArrayLike, sampling_rate: int, device: Union[str, torch.device] = "cuda",
                    target_rate: int = 16000) -> torch.Tensor:
    """``scipy.signal.resample(x, round(len(x) * 16000 / sampling_rate))`` (eval_utils.py:12-14) on the device.
    audio: 1-D clip.  Returns a 1-D fp32 CUDA tensor.  Restated algorithm (real input): X = rfft(x); keep the lowest
    min(N, num)//2 + 1 bins; when that count's N is even, the shared Nyquist bin is doubled (down-sampling) or halved
    (up-sampling); y = irfft(Y, num) * num / N."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("cacophony_b200 runs on a CUDA device only (no CPU fallback)")
    x = torch.as_tensor(audio, dtype=torch.float32).to(dev)
    if x.dim() != 1:
        raise ValueError("resample_to_16k: one 1-D clip expected")
    nx = x.shape[0]
    if sampling_rate == target_rate:
        return x
    num = int(round(nx * float(target_rate) / sampling_rate))
    X = torch.fft.rfft(x)
    Y = torch.zeros(num // 2 + 1, dtype=X.dtype, device=dev)
    n = min(num, nx)
    nyq = n // 2 + 1
    Y[:nyq] = X[:nyq]
    if n % 2 == 0:
        if num < nx:
            Y[n // 2] *= 2.0
        elif nx < num:
            Y[n // 2] *= 0.5
    return torch.fft.irfft(Y, n=num) * (float(num) / float(nx))


def load_audio(audio_path: str, dataset_sampling_rate: int, device: Union[str, torch.device] = "cuda") -> torch.Tensor:
    """eval_utils.py:6-16 with the resampling on the device.  Reads with soundfile when it is installed, else PCM/float WAV
    through scipy.io.wavfile (int PCM scaled to [-1, 1) as soundfile does)."""
    try:
        import soundfile as sf
        wav, _ = sf.read(audio_path)
        wav = np.asarray(wav, dtype=np.float32)
    except ImportError:
        from scipy.io import wavfile
        _, raw = wavfile.read(audio_path)
        if np.issubdtype(raw.dtype, np.integer):
            wav = raw.astype(np.float32) / float(2 ** (8 * raw.dtype.itemsize - 1))
        else:
            wav = raw.astype(np.float32)
    if wav.ndim > 1:
        wav = wav.mean(axis=-1)
    return resample_to_16k(wav, dataset_sampling_rate, device)
