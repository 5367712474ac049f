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


def pad_ragged_device(waves: Sequence[torch.Tensor], stride: Optional[int] = None, multiple: int = 160
                      ) -> Tuple[torch.Tensor, torch.Tensor]:
    """`pad_ragged` for clips that already live on the GPU (e.g. the output of `load_audio` / `resample_to_16k`): one
    zero-filled [batch, stride] device buffer filled with device-to-device copies on the current stream, plus device int32
    lengths — no host round trip between the resampler and the frontend."""
    if len(waves) == 0:
        raise ValueError("pad_ragged_device: empty batch")
    dev = waves[0].device
    if dev.type != "cuda" or any((not isinstance(w, torch.Tensor)) or w.device != dev or w.dim() != 1 for w in waves):
        raise ValueError("pad_ragged_device: 1-D CUDA tensors on one device expected")
    longest = max(int(w.shape[0]) for w in waves)
    if stride is None:
        stride = max(multiple, -(-longest // multiple) * multiple)
    with torch.cuda.device(dev):
        buf = torch.zeros((len(waves), stride), dtype=torch.float32, device=dev)
        ns = []
        for i, w in enumerate(waves):
            n = min(int(w.shape[0]), stride)
            buf[i, :n].copy_(w[:n], non_blocking=True)
            ns.append(n)
        lens = torch.tensor(ns, dtype=torch.int32).to(dev, non_blocking=True)
    return buf, lens


class PinnedStager:
    """Double-buffered pinned staging for raw file samples: `put(array)` copies a host array into one of two reusable pinned
    buffers and enqueues the host-to-device copy on the current stream; a buffer is reused only after the copy that read it
    has completed (one event per buffer), so file reading for clip k+1 overlaps the transfer (and the GPU work) of clip k.
    Replaces the per-file pageable `tensor.to(device)` (a synchronous copy) of a naive loader."""

    def __init__(self, device: Union[str, torch.device], initial: int = 1 << 20):
        self.device = torch.device(device)
        self._buf = [torch.empty(initial, dtype=torch.float32).pin_memory() for _ in range(2)]
        self._ev: List[Optional[torch.cuda.Event]] = [None, None]
        self._k = 0

    def put(self, array: np.ndarray) -> torch.Tensor:
        a = np.ascontiguousarray(array, dtype=np.float32).reshape(-1)
        k = self._k
        self._k ^= 1
        if self._ev[k] is not None:
            self._ev[k].synchronize()
        if self._buf[k].numel() < a.shape[0]:
            self._buf[k] = torch.empty(max(a.shape[0], 2 * self._buf[k].numel()), dtype=torch.float32).pin_memory()
        self._buf[k].numpy()[: a.shape[0]] = a
        with torch.cuda.device(self.device):
            out = self._buf[k][: a.shape[0]].to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        self._ev[k] = ev
        return out


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
    x = audio.to(device=dev, dtype=torch.float32) if isinstance(audio, torch.Tensor) else \
        torch.as_tensor(audio, dtype=torch.float32).to(dev)
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


def load_audio(audio_path: str, dataset_sampling_rate: int, device: Union[str, torch.device] = "cuda",
               stager: Optional[PinnedStager] = None) -> torch.Tensor:
    """eval_utils.py:6-16 with the resampling on the device; returns a 1-D fp32 CUDA tensor that STAYS on the device (feed
    lists of them to `pad_ragged_device` / `eval.embed_waveforms`).  Reads with soundfile when it is installed, else PCM/float
    WAV through scipy.io.wavfile (int PCM scaled to [-1, 1) as soundfile does).  stager: a `PinnedStager` makes the
    host-to-device copy of the raw samples asynchronous."""
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
    if stager is not None:
        return resample_to_16k(stager.put(wav), dataset_sampling_rate, device)
    return resample_to_16k(wav, dataset_sampling_rate, device)
