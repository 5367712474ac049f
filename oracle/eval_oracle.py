"""CPU restatement (numpy, float64 where the reference uses it) of the evaluation-side steps around the hot path.
TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing in cacophony_b200/).

Follows, line by line:
  topk_indices              torch.argsort(-logits, dim=-1)[..., :k]            src/eval/eval_caco_torch.py:331,401,406
  retrieval_preds / retrieval_per_query / compute_retrieval_metric            src/eval/eval_utils.py:18-56
  jackknife_stats           astropy.stats.jackknife_stats as called at        src/eval/eval_utils.py:57-66
                            (astropy is a third-party dependency, unpinned in requirements_torch.txt and not installed here:
                            restated from its documented algorithm — leave-one-out resamples, bias, std error, normal
                            interval; "parity unpinned" for the interval only, the metric values themselves are pinned by
                            tests/golden/eval_retrieval.npz generated from the reference function)
  zero_shot_topk            logits = exp(logit_scale) * a @ t.T ; argsort     src/eval/eval_caco_torch.py:330-331
  avg_pool_tokens           tf.nn.avg_pool(x, ksize=8, strides=8, 'VALID')    src/eval/heareval/.../caco_embeddings.py:124-125
  resample                  scipy.signal.resample (third-party, installed: the oracle simply calls it, eval_utils.py:12-14)
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Sequence

import numpy as np


def topk_indices(x: np.ndarray, k: int) -> np.ndarray:
    """Descending order, ties to the lower column (stable sort), NaN last."""
    x = np.asarray(x, dtype=np.float32)
    key = np.where(np.isnan(x), -np.inf, x)
    order = np.argsort(-key, axis=-1, kind="stable")
    nan_last = np.argsort(np.take_along_axis(np.isnan(x), order, axis=-1), axis=-1, kind="stable")
    return np.take_along_axis(order, nan_last, axis=-1)[..., :k]


def retrieval_preds(indices, all_querys, all_keys, gt_query_key, retrieval_type="at") -> np.ndarray:
    """eval_utils.py:26-41: the boolean `preds` vector (10 ranks) of every query."""
    out = np.zeros((len(all_querys), 10), dtype=bool)
    for i, query in enumerate(all_querys):
        pred_keys = [all_keys[idx] for idx in indices[i, :10]]
        if retrieval_type == "at":
            seen = []
            for j, pred in enumerate(pred_keys):
                if pred not in seen and pred in gt_query_key[query]:
                    seen.append(pred)
                    out[i, j] = True
        elif retrieval_type == "ta":
            for j, pred in enumerate(pred_keys):
                out[i, j] = gt_query_key[query] == pred
    return out


def retrieval_per_query(preds: np.ndarray) -> Dict[str, np.ndarray]:
    """eval_utils.py:43-56."""
    R1, R5, R10, mAP10 = [], [], [], []
    for p in preds:
        R1.append(np.sum(np.any(p[:1]), dtype=float))
        R5.append(np.sum(np.any(p[:5]), dtype=float))
        R10.append(np.sum(np.any(p[:10]), dtype=float))
        positions = np.arange(1, 11, dtype=float)[p[:10] > 0]
        if len(positions) > 0:
            precisions = np.divide(np.arange(1, len(positions) + 1, dtype=float), positions)
            mAP10.append(np.mean(precisions, dtype=float))
        else:
            mAP10.append(0.0)
    return {"R1": np.asarray(R1), "R5": np.asarray(R5), "R10": np.asarray(R10), "mAP10": np.asarray(mAP10)}


def jackknife_stats(data: np.ndarray, statistic=np.mean, confidence_level: float = 0.95):
    """Generic O(n^2) leave-one-out form of astropy.stats.jackknife_stats."""
    from scipy.special import erfinv
    data = np.asarray(data, dtype=np.float64)
    n = data.shape[0]
    resamples = np.empty(n)
    for i in range(n):
        resamples[i] = statistic(np.delete(data, i))
    stat = statistic(data)
    mean_jack = np.mean(resamples)
    bias = (n - 1) * (mean_jack - stat)
    std_err = np.sqrt((n - 1) * np.mean((resamples - mean_jack) * (resamples - mean_jack)))
    estimate = stat - bias
    z = np.sqrt(2.0) * erfinv(confidence_level)
    return estimate, bias, std_err, estimate + z * np.array((-std_err, std_err))


def compute_retrieval_metric(indices, all_querys, all_keys, gt_query_key, retrieval_type="at") -> Dict[str, Any]:
    per_query = retrieval_per_query(retrieval_preds(indices, all_querys, all_keys, gt_query_key, retrieval_type))
    out: Dict[str, Any] = {"per_query": per_query}
    for name, vals in per_query.items():
        est, _, _, ci = jackknife_stats(vals, np.mean, 0.95)
        out[name] = (float(est), float(ci[0]), float(ci[1]))
    return out


def zero_shot_topk(logit_scale: float, a: np.ndarray, t: np.ndarray, k: int = 1) -> np.ndarray:
    logits = np.float32(math.exp(logit_scale)) * np.asarray(a, np.float32) @ np.asarray(t, np.float32).T
    return topk_indices(logits, k)


def avg_pool_tokens(hid: np.ndarray, group: int = 8) -> np.ndarray:
    B, S, D = hid.shape
    n = S // group
    return hid[:, :n * group].reshape(B, n, group, D).mean(axis=2, dtype=np.float32)


def resample(x: np.ndarray, sampling_rate: int, target_rate: int = 16000) -> np.ndarray:
    import scipy.signal
    num = round(x.shape[-1] * float(target_rate) / sampling_rate)
    return scipy.signal.resample(x, num)


def make_retrieval_case(seed: int, n_audio: int, caps_per_audio: int, dup_every: int = 0):
    """Deterministic synthetic retrieval problem: names, captions (optionally with duplicated caption strings and a repeated
    audio name, the two corner cases of eval_utils.py's string-keyed dictionaries), random ranked indices."""
    rng = np.random.default_rng(seed)
    names = [f"clip{i:04d}" for i in range(n_audio)]
    if dup_every:
        names[-1] = names[0]                                   # same file name twice: gt dict entries collide
    all_text, gt_at, gt_ta = [], {}, {}
    for i, nm in enumerate(names):
        gt_at[nm] = []
        for c in range(caps_per_audio):
            s = f"caption {i} {c}"
            if dup_every and (i * caps_per_audio + c) % dup_every == 0:
                s = f"shared caption {(i * caps_per_audio + c) // dup_every % 3}"
            gt_at[nm].append(s)
            gt_ta[s] = nm
            all_text.append(s)
    n_text = len(all_text)
    # rankings: mostly random, with the true keys planted at random ranks for a third of the queries
    at_idx = np.stack([rng.permutation(n_text)[:10] for _ in range(n_audio)]).astype(np.int64)
    ta_idx = np.stack([rng.permutation(n_audio)[:10] for _ in range(n_text)]).astype(np.int64)
    for i in range(0, n_audio, 3):
        at_idx[i, rng.integers(0, 10)] = i * caps_per_audio + rng.integers(0, caps_per_audio)
        at_idx[i, rng.integers(0, 10)] = i * caps_per_audio
    for j in range(0, n_text, 3):
        ta_idx[j, rng.integers(0, 10)] = j // caps_per_audio
    return names, all_text, gt_at, gt_ta, at_idx, ta_idx
