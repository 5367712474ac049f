"""Row (f-2 / f-3): the callers on either side of the hot path, with the reference's names
(src/eval/eval_caco_torch.py:154-408, src/eval/eval_utils.py:18-66), batched and with the O(N·M) parts on the device.

Reference flow (one clip and one caption per model call, batch dimension 1, argsort of whole logit rows on the
device then numpy on the host)            ->  here
  compute_all_class_embeddings :264-286   ->  one tokenizer call for all prompts, text tower in batches
  zs_classification            :289-340   ->  clips batched through the ragged frontend + audio tower, ONE logits matrix
                                              exp(logit_scale)·A·Tᵀ (caco_sim_logits), top-k per row (caco_topk_rows)
  audio_retrieval              :343-408   ->  both towers batched, T·Aᵀ and A·Tᵀ (caco_sgemm_nt), top-10 per row,
                                              per-query hit masks (caco_retrieval_hits); only [queries] int32 reach the host
  compute_retrieval_metric  eval_utils:18  ->  same R@1/5/10, mAP@10 and jackknife 95 % intervals, float64 on the host

Not available offline and therefore injected by the caller: the ``roberta-base`` tokenizer (any callable with the
``RobertaTokenizerFast.__call__`` keyword interface) and the dataset processors (anything with
``get_filepaths_and_descriptions`` and ``config.sampling_rate``).  The ``*_arrays`` entry points take in-memory
waveforms instead of file paths.
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import loader, ops
from .frontend import DatasetConfig, prepare_audio_batch
from .model import CACO, create_caco_model


# ----------------------------------------------------------------------------------------------------------- loading
def load_caco_torch(ckpt_path: Optional[str], device: Union[str, torch.device], tokenizer: Any = None,
                    state_dict: Optional[Dict[str, torch.Tensor]] = None, pool_heads: Optional[int] = None) -> Dict[str, Any]:
    """eval_caco_torch.py:154-178.  Accepts the three checkpoint layouts the reference accepts ('model_state_dict',
    'state_dict', or a bare state_dict) and, beyond the reference, the original Flax checkpoint file (msgpack, converted on
    the fly by ``checkpoint.convert_caco_checkpoint``); ``decoder_module.*`` tensors are ignored.  pool_heads: audio pooler
    head count if not the torch port's 2 (the JAX loader uses 8, load_model.py:47).  The tokenizer cannot be downloaded
    here: pass one in (``RobertaTokenizerFast.from_pretrained('roberta-base')`` where the files exist)."""
    from . import checkpoint as ckpt
    model = create_caco_model() if pool_heads is None else create_caco_model(num_attention_pool_heads=pool_heads)
    if state_dict is None:
        if ckpt_path is None:
            raise ValueError("load_caco_torch: ckpt_path or state_dict required")
        with open(ckpt_path, "rb") as f:
            head = f.read(4)
        if head[:2] == b"PK" or head[:1] == b"\x80":            # torch.save: zip archive (or legacy pickle)
            checkpoint = torch.load(ckpt_path, map_location="cpu")
        else:                                                    # flax.training.checkpoints: msgpack
            checkpoint = ckpt.convert_caco_checkpoint(ckpt_path)
    else:
        checkpoint = state_dict
    if "model_state_dict" in checkpoint:
        model.load_state_dict(checkpoint["model_state_dict"])
    elif "state_dict" in checkpoint:
        model.load_state_dict(checkpoint["state_dict"])
    else:
        model.load_state_dict(checkpoint)
    model = model.to(device)
    model.eval()
    if tokenizer is None:
        try:
            from transformers import RobertaTokenizerFast
            tokenizer = RobertaTokenizerFast.from_pretrained("roberta-base", local_files_only=True)
        except Exception:
            tokenizer = None          # offline: callers pass token ids / their own tokenizer
    return {"model": model, "tokenizer": tokenizer, "device": torch.device(device)}


def prepare_text_batch(text: Union[str, Sequence[str]], tokenizer: Any, max_text_len: int,
                       device: Union[str, torch.device]) -> Dict[str, torch.Tensor]:
    """eval_caco_torch.py:209-227 for one caption or a list of captions."""
    if tokenizer is None:
        raise ValueError("prepare_text_batch: a tokenizer is required (roberta-base files are not available offline)")
    texts = [text] if isinstance(text, str) else list(text)
    tokenized = tokenizer(texts, padding="max_length", truncation=True, max_length=max_text_len, return_tensors="pt")
    return {"text_input_ids": tokenized["input_ids"].to(device), "text_mask": tokenized["attention_mask"].to(device)}


@torch.no_grad()
def compute_audio_embedding(model: CACO, audio_batch: Dict[str, torch.Tensor]) -> torch.Tensor:
    """eval_caco_torch.py:230-245."""
    return model.get_audio_embedding(audio_patches=audio_batch["audio_patches"], audio_time_inds=audio_batch["audio_time_inds"],
                                     audio_freq_inds=audio_batch["audio_freq_inds"], audio_mask=audio_batch["audio_mask"],
                                     deterministic=True, return_hidden_state=False, normalize=True)


@torch.no_grad()
def compute_text_embedding(model: CACO, text_batch: Dict[str, torch.Tensor]) -> torch.Tensor:
    """eval_caco_torch.py:248-261."""
    return model.get_text_embedding(text_input_ids=text_batch["text_input_ids"], text_mask=text_batch["text_mask"],
                                    deterministic=True, return_hidden_state=False, normalize=True)


@torch.no_grad()
def embed_text_ids(model: CACO, ids: torch.Tensor, mask: torch.Tensor, batch_size: int = 512, trim_padding: bool = True
                   ) -> torch.Tensor:
    """L2-normalised text embeddings of already-tokenised captions, in batches.  trim_padding: run the tower on the columns
    up to the last valid token of the batch (rounded up to 8) instead of the full padded length — attention is causal and
    key-masked and the pooler is masked, so columns past every caption's end cannot influence any embedding (prompts
    padded to 100 tokens with ~10 valid: 6x fewer token rows)."""
    outs = []
    for i in range(0, ids.shape[0], batch_size):
        bi, bm = ids[i:i + batch_size], mask[i:i + batch_size]
        if trim_padding and bm.shape[1] > 8:
            cols = (bm != 0).any(dim=0).nonzero()
            last = int(cols.max()) + 1 if cols.numel() else 1
            t_eff = min(bm.shape[1], -(-last // 8) * 8)
            bi, bm = bi[:, :t_eff].contiguous(), bm[:, :t_eff].contiguous()
        outs.append(model.encode_text(bi, bm))
    return torch.cat(outs, dim=0) if len(outs) > 1 else outs[0]


@torch.no_grad()
def embed_waveforms(model: CACO, waves: Sequence[Any], datasetconfig: Optional[DatasetConfig] = None, batch_size: int = 256,
                    trim_padding: bool = True) -> torch.Tensor:
    """L2-normalised audio embeddings of a list of (ragged) 16 kHz clips, `batch_size` clips per library call.  Host clips
    (numpy / CPU tensors): pinned ragged packing -> async H2D.  Clips already on the model's device (what `loader.load_audio`
    returns): packed with device-to-device copies — nothing goes back to the host between the resampler and the tower."""
    cfg = datasetconfig or DatasetConfig()
    dev = model._device()
    outs = []
    for i in range(0, len(waves), batch_size):
        chunk = waves[i:i + batch_size]
        if all(isinstance(w, torch.Tensor) and w.is_cuda for w in chunk):
            w, lens = loader.pad_ragged_device([c.to(dev) for c in chunk])
            longest = int(max(c.shape[0] for c in chunk))
        else:
            buf, lens = loader.pad_ragged(chunk)
            w = buf.to(dev, non_blocking=True)
            longest = int(lens.max())
        outs.append(_encode_ragged(model, w, lens, longest, cfg.patches_seq_len, trim_padding))
    return torch.cat(outs, dim=0) if len(outs) > 1 else outs[0]


def _encode_ragged(model: CACO, w: torch.Tensor, lens: torch.Tensor, longest: int, max_patches: int, trim_padding: bool
                   ) -> torch.Tensor:
    """encode_audio on a packed ragged batch; the trimmed patch count comes from the host-known longest clip, so device
    lengths never have to be read back."""
    P = max_patches
    if trim_padding:
        valid = (((min(longest, w.shape[1]) + 159) // 160) // 16) * 8          # eval_caco_torch.py:67,116-117
        P = max(8, min(P, valid))
    return model.encode_audio(w, max_patches=P, lengths=lens, trim_padding=False)


@torch.no_grad()
def embed_files(model: CACO, filepaths: Sequence[str], sampling_rate: int, datasetconfig: Optional[DatasetConfig] = None,
                batch_size: int = 64, trim_padding: bool = True) -> torch.Tensor:
    """Files -> embeddings without a host round trip: every file's raw samples go through a double-buffered pinned stage
    (async H2D), are resampled on the device and packed device-side; while the tower of batch k runs (enqueued
    asynchronously), the host reads the files of batch k+1."""
    dev = model._device()
    stager = loader.PinnedStager(dev)
    outs = []
    for i in range(0, len(filepaths), batch_size):
        clips = [loader.load_audio(fp, sampling_rate, dev, stager) for fp in filepaths[i:i + batch_size]]
        outs.append(embed_waveforms(model, clips, datasetconfig, batch_size, trim_padding))
    return torch.cat(outs, dim=0) if len(outs) > 1 else outs[0]


@torch.no_grad()
def compute_all_class_embeddings(model: CACO, tokenizer: Any, class_list: List[str], max_text_len: int,
                                 device: Union[str, torch.device], prefix: str = "", batch_size: int = 512) -> torch.Tensor:
    """eval_caco_torch.py:264-286: [n_classes, 768], one tokenizer call, text tower in batches."""
    tb = prepare_text_batch([prefix + c for c in class_list], tokenizer, max_text_len, device)
    return embed_text_ids(model, tb["text_input_ids"], tb["text_mask"], batch_size)


# ------------------------------------------------------------------------------------------------------- captioning
# decode_caption_ids' default: incremental decode against a key / value cache (True) or the reference's literal loop, which
# re-runs text tower + decoder on the whole prefix for every token (False).  The logits are bit-identical either way
# (tests/test_model_gpu.py::test_kv_cached_decode_matches_full_prefix); 1.2 - 1.8 ms per step instead of 1.6 - 5.5.
USE_KV_CACHE = True


@torch.no_grad()
def decode_caption_ids(model: CACO, audio_batch: Dict[str, torch.Tensor], bos_id: int = 0, eos_id: int = 2,
                       max_decode_length: int = 100, temperature: float = 0.0,
                       generator: Optional[torch.Generator] = None, use_cache: Optional[bool] = None,
                       use_graph: Any = False) -> torch.Tensor:
    """The loop of decode_caption (eval_caco_torch.py:411-472) on token ids, for a whole batch: BOS, then one token per step
    from the decoder logits of the sequence so far (the reference's own call passes keyword names ``RobertaDecoder.forward``
    does not have — SURVEY.md 2 row 6 — so the call it means, ``CACO.get_decoder_logits``, is the one made here).
    temperature 0: greedy (device arg-max, ``caco_topk_rows``); > 0: ``softmax(logits / temperature)`` sampled with
    torch.multinomial as the reference does.  Returns [batch, <= L+1] ids; a sequence that has produced EOS keeps emitting EOS.

    use_cache (default ``USE_KV_CACHE``): False re-runs ``get_decoder_logits`` on the full prefix every step, as the reference
    does; True pushes only the newest token through text tower and decoder against cached keys / values
    (``CACO.decode_begin`` / ``decode_step``) — both towers are causal, so the logits are the same.
    use_graph (with the cache): replay each step from one CUDA graph — a step is ~140 small launches, launch-bound at small
    batches.  True captures a ``serving.GraphedDecodeStep`` for this call; pass an instance to reuse one across requests."""
    _, audio_hidden = model.get_audio_embedding(audio_patches=audio_batch["audio_patches"],
                                                audio_time_inds=audio_batch["audio_time_inds"],
                                                audio_freq_inds=audio_batch["audio_freq_inds"],
                                                audio_mask=audio_batch["audio_mask"], deterministic=True,
                                                return_hidden_state=True, normalize=False)
    dev = audio_hidden.device
    B = audio_hidden.shape[0]
    generated = torch.full((B, 1), bos_id, dtype=torch.long, device=dev)
    done = torch.zeros(B, dtype=torch.bool, device=dev)
    if use_cache is None:
        use_cache = USE_KV_CACHE
    stepper = None
    if use_cache:
        capacity = max(1, min(int(max_decode_length), model.text_config.max_position_embeddings))
        if use_graph:
            from .serving import GraphedDecodeStep
            if isinstance(use_graph, GraphedDecodeStep):          # a captured step kept across requests of one shape
                stepper = use_graph
                if (stepper.batch, stepper.seq) != (B, int(audio_hidden.shape[1])) or stepper.capacity < capacity:
                    raise ValueError("use_graph: the captured step was made for another batch / audio length / capacity")
            else:
                stepper = GraphedDecodeStep(model, B, int(audio_hidden.shape[1]), capacity)
            stepper.begin(audio_hidden, audio_batch["audio_mask"])
        else:
            cache = model.decode_begin(audio_hidden, audio_batch["audio_mask"], capacity)
        pos = torch.zeros(B, dtype=torch.long, device=dev)
    for step in range(max_decode_length):
        if use_cache:
            if step >= capacity:                  # the full-prefix call fails at the same length (get_text_embedding)
                raise ValueError(f"text sequence length must be <= {model.text_config.max_position_embeddings}")
            tok = generated[:, -1].contiguous()
            if stepper is not None:
                last, greedy = stepper.step(tok, pos)
            elif temperature > 0:
                last, greedy = model.decode_step(cache, tok, pos), None
            else:
                last, greedy = None, model.decode_step(cache, tok, pos, want_logits=False, want_next=True)
            pos = pos + 1
        else:
            text_mask = torch.ones(generated.shape, dtype=torch.float32, device=dev)
            logits = model.get_decoder_logits(audio_hidden, audio_batch["audio_mask"], generated, text_mask)
            last, greedy = logits[:, -1, :].contiguous(), None
        if temperature > 0:
            nxt = torch.multinomial(torch.softmax(last / temperature, dim=-1), num_samples=1, generator=generator)
        elif greedy is not None:
            nxt = greedy.long()[:, None]
        else:
            nxt = ops.topk_rows(last, 1).long()
        nxt = torch.where(done[:, None], torch.full_like(nxt, eos_id), nxt)
        generated = torch.cat([generated, nxt], dim=1)
        done = done | (nxt[:, 0] == eos_id)
        if bool(done.all()):
            break
    return generated


@torch.no_grad()
def decode_caption(model: CACO, tokenizer: Any, audio_batch: Dict[str, torch.Tensor], max_decode_length: int = 100,
                   temperature: float = 0.1) -> str:
    """eval_caco_torch.py:411-472 (same arguments and return value: the decoded caption of the first clip)."""
    if model.decoder_module is None:
        raise ValueError("Model does not have a decoder module. Load with use_decoder=True.")
    ids = decode_caption_ids(model, audio_batch, tokenizer.bos_token_id, tokenizer.eos_token_id, max_decode_length, temperature)
    return tokenizer.batch_decode(ids, skip_special_tokens=True)[0].strip()


# ------------------------------------------------------------------------------------------------------- zero-shot
@torch.no_grad()
def zero_shot_logits(model: CACO, audio_embeddings: torch.Tensor, all_text_embeddings: torch.Tensor) -> torch.Tensor:
    """eval_caco_torch.py:330: exp(logit_scale) * audio_embedding @ all_text_embeddings.T for all clips at once."""
    at, _ = model.similarity(audio_embeddings, all_text_embeddings, want_ta=False)
    return at


@torch.no_grad()
def zero_shot_topk(model: CACO, audio_embeddings: torch.Tensor, all_text_embeddings: torch.Tensor, k: int = 1) -> torch.Tensor:
    """eval_caco_torch.py:330-331: argsort(-logits)[:, :k] as int32 [n_clips, k] (device)."""
    return ops.topk_rows(zero_shot_logits(model, audio_embeddings, all_text_embeddings), k)


@torch.no_grad()
def zs_classification_arrays(model: CACO, all_text_embeddings: torch.Tensor, waves: Sequence[Any], target_indices: Sequence[int],
                             datasetconfig: Optional[DatasetConfig] = None, ks: Sequence[int] = (1,), batch_size: int = 256
                             ) -> Dict[str, float]:
    """The loop body of zs_classification (eval_caco_torch.py:315-336) for in-memory clips: top-k accuracy per k."""
    a = embed_waveforms(model, waves, datasetconfig, batch_size)
    return zs_accuracy_from_embeddings(model, a, all_text_embeddings, target_indices, ks)


@torch.no_grad()
def zs_accuracy_from_embeddings(model: CACO, audio_embeddings: torch.Tensor, all_text_embeddings: torch.Tensor,
                                target_indices: Sequence[int], ks: Sequence[int] = (1,)) -> Dict[str, float]:
    """eval_caco_torch.py:330-336: top-k accuracy from the logits matrix, ranking on the device."""
    kmax = max(ks)
    top = zero_shot_topk(model, audio_embeddings, all_text_embeddings, kmax)
    tgt = torch.as_tensor(list(target_indices), dtype=torch.int32, device=top.device)[:, None]
    out = {}
    for k in ks:
        out[str(k)] = float((top[:, :k] == tgt).any(dim=1).float().mean().item())
    return out


def zs_classification(model: CACO, tokenizer: Any, dataprocessor: Any, datasetconfig: DatasetConfig,
                      device: Union[str, torch.device], subdir_name: str = "", text_prefix: str = "This is a sound of ") -> float:
    """eval_caco_torch.py:289-340 (same arguments, same printed lines, same return value)."""
    filepaths, descriptions, _ = dataprocessor.get_filepaths_and_descriptions(current_split=subdir_name)
    class_labels = sorted(set(descriptions[a]["description"][0] for a in descriptions))
    class_to_index_map = {v: i for i, v in enumerate(class_labels)}
    all_text_embeddings = compute_all_class_embeddings(model, tokenizer, class_labels, datasetconfig.max_text_len, device,
                                                       prefix=text_prefix)
    targets = []
    for fp in filepaths:
        audio_name = fp.split("/")[-1].split(".wav")[0]
        targets.append(class_to_index_map[descriptions[audio_name]["description"][0]])
    a = embed_files(model, filepaths, dataprocessor.config.sampling_rate, datasetconfig)
    acc = zs_accuracy_from_embeddings(model, a, all_text_embeddings, targets, ks=(1,))
    for k, v in acc.items():
        print(f"top {k} accuracy: {v:.4f}")
    return acc["1"]


# ------------------------------------------------------------------------------------------------------- retrieval
def jackknife_stats_mean(data: np.ndarray, confidence_level: float = 0.95) -> Tuple[float, float, float, np.ndarray]:
    """``astropy.stats.jackknife_stats(data, np.mean, confidence_level)`` (called at eval_utils.py:57-66; astropy is a
    third-party dependency of the reference, not vendored, restated from its documented algorithm): leave-one-out
    resamples, bias = (n-1)(mean(resamples) - stat), std_err = sqrt((n-1) mean((resamples - mean(resamples))^2)),
    estimate = stat - bias, interval = estimate ± z·std_err with z = sqrt(2)·erfinv(confidence_level).  For the mean
    statistic the resamples have the closed form (sum - x_i)/(n-1), so no O(n^2) loop is needed."""
    from scipy.special import erfinv
    x = np.asarray(data, dtype=np.float64)
    n = x.shape[0]
    if n < 2:
        raise ValueError("jackknife_stats_mean: at least two samples required")
    stat = x.mean()
    resamples = (x.sum() - x) / (n - 1)
    mean_jack = resamples.mean()
    bias = (n - 1) * (mean_jack - stat)
    std_err = math.sqrt((n - 1) * np.mean((resamples - mean_jack) * (resamples - mean_jack)))
    estimate = stat - bias
    z = math.sqrt(2.0) * float(erfinv(confidence_level))
    return float(estimate), float(bias), float(std_err), estimate + z * np.array((-std_err, std_err))


def metrics_from_hit_bits(bits: np.ndarray) -> Dict[str, np.ndarray]:
    """eval_utils.py:43-56 from the per-query hit masks (bit j = rank j+1 hit): per-query R1, R5, R10, AP@10 in float64."""
    bits = np.asarray(bits, dtype=np.int64)
    preds = ((bits[:, None] >> np.arange(10)[None, :]) & 1).astype(bool)              # [Q, 10]
    r1 = preds[:, :1].any(axis=1).astype(np.float64)
    r5 = preds[:, :5].any(axis=1).astype(np.float64)
    r10 = preds[:, :10].any(axis=1).astype(np.float64)
    ap = np.zeros(bits.shape[0], dtype=np.float64)
    positions = np.arange(1, 11, dtype=np.float64)
    for q in range(bits.shape[0]):                                                     # O(Q) host work, as in the reference
        pos = positions[preds[q]]
        if len(pos) > 0:
            ap[q] = np.mean(np.arange(1, len(pos) + 1, dtype=np.float64) / pos, dtype=np.float64)
    return {"R1": r1, "R5": r5, "R10": r10, "mAP10": ap}


def _ids(names: Sequence[Any], table: Optional[Dict[Any, int]] = None) -> Tuple[np.ndarray, Dict[Any, int]]:
    table = {} if table is None else table
    out = np.empty(len(names), dtype=np.int32)
    for i, n in enumerate(names):
        out[i] = table.setdefault(n, len(table))
    return out, table


def compute_retrieval_metric(indices: Union[np.ndarray, torch.Tensor], all_querys: Sequence[Any], all_keys: Sequence[Any],
                             gt_query_key: Dict[Any, Any], retrieval_type: str = "at", device: Union[str, torch.device] = "cuda",
                             verbose: bool = True) -> Dict[str, Any]:
    """eval_utils.py:18-66 (same arguments; prints the same four lines and also returns them).  `indices`: [queries, >= 10]
    ranked key indices (numpy or a device tensor straight from ``ops.topk_rows``).  The per-query hit test runs on the
    device (caco_retrieval_hits); names are mapped to integer ids on the host once."""
    dev = torch.device(device)
    top = torch.as_tensor(indices)[:, :10].to(device=dev, dtype=torch.int32).contiguous()
    key_id, table = _ids(all_keys)
    if retrieval_type == "ta":
        gt = np.array([table.get(gt_query_key[q], -1) for q in all_querys], dtype=np.int32)        # unknown audio: never hit
        bits = ops.retrieval_hits(top, torch.from_numpy(key_id).to(dev), torch.from_numpy(gt).to(dev))
    elif retrieval_type == "at":
        qid, qtable = _ids(all_querys)
        n_key_ids = max(1, len(table))
        pairs = set()
        for q, g in qtable.items():
            for key in gt_query_key[q]:
                if key in table:
                    pairs.add(int(g) * n_key_ids + table[key])
        pairs_t = torch.tensor(sorted(pairs) or [-1], dtype=torch.int64, device=dev)
        bits = ops.retrieval_hits(top, torch.from_numpy(key_id).to(dev), torch.from_numpy(qid).to(dev), pairs_t, n_key_ids)
    else:
        raise ValueError("retrieval_type must be 'at' or 'ta'")
    per_query = metrics_from_hit_bits(bits.cpu().numpy())
    out: Dict[str, Any] = {}
    for name in ("R1", "R5", "R10", "mAP10"):
        estimate, _, _, ci = jackknife_stats_mean(per_query[name], 0.95)
        out[name] = (estimate, float(ci[0]), float(ci[1]))
        if verbose:
            print(name, f"{estimate:.3f}", f"[{ci[0]:.3f}, {ci[1]:.3f}]")
    out["per_query"] = per_query
    return out


@torch.no_grad()
def retrieval_topk(text_embeddings: torch.Tensor, audio_embeddings: torch.Tensor, k: int = 10) -> Tuple[torch.Tensor, torch.Tensor]:
    """eval_caco_torch.py:396-406: logits_ar = T·Aᵀ; returns (argsort(-logits_arᵀ)[:, :k] — audio -> text,
    argsort(-logits_ar)[:, :k] — text -> audio), int32 on the device.  Both products are computed directly (fp32 FMA)."""
    ta = ops.sgemm_nt(text_embeddings.contiguous(), audio_embeddings.contiguous())          # [n_text, n_audio]
    at = ops.sgemm_nt(audio_embeddings.contiguous(), text_embeddings.contiguous())          # [n_audio, n_text] = taᵀ
    return ops.topk_rows(at, min(k, at.shape[1])), ops.topk_rows(ta, min(k, ta.shape[1]))


@torch.no_grad()
def audio_retrieval_arrays(model: CACO, waves: Optional[Sequence[Any]], audio_names: Sequence[str],
                           captions: Sequence[Sequence[str]], tokenizer: Any, datasetconfig: Optional[DatasetConfig] = None,
                           verbose: bool = True, audio_embeddings: Optional[torch.Tensor] = None) -> Dict[str, Any]:
    """audio_retrieval (eval_caco_torch.py:343-408) for in-memory clips: captions[i] = the descriptions of clip i
    (audio_embeddings: already-computed [n_clips, 768] embeddings instead of `waves`)."""
    cfg = datasetconfig or DatasetConfig()
    dev = model._device()
    all_text, gt_audio_text, gt_text_audio = [], {}, {}
    for name, caps in zip(audio_names, captions):
        gt_audio_text[name] = []
        for c in caps:
            gt_audio_text[name].append(c)
            gt_text_audio[c] = name
            all_text.append(c)
    tb = prepare_text_batch(all_text, tokenizer, cfg.max_text_len, dev)
    t = embed_text_ids(model, tb["text_input_ids"], tb["text_mask"])
    a = audio_embeddings if audio_embeddings is not None else embed_waveforms(model, waves, cfg)
    at_idx, ta_idx = retrieval_topk(t, a, 10)
    if at_idx.shape[1] < 10 or ta_idx.shape[1] < 10:
        raise ValueError("audio_retrieval needs at least 10 clips and 10 captions (the reference indexes indices[i, :10])")
    if verbose:
        print("audio to text retrieval:")
    res_at = compute_retrieval_metric(at_idx, list(audio_names), all_text, gt_audio_text, "at", dev, verbose)
    if verbose:
        print("text to audio retrieval:")
    res_ta = compute_retrieval_metric(ta_idx, all_text, list(audio_names), gt_text_audio, "ta", dev, verbose)
    return {"at": res_at, "ta": res_ta}


def audio_retrieval(model: CACO, tokenizer: Any, dataprocessor: Any, datasetconfig: DatasetConfig,
                    device: Union[str, torch.device], eval_split: str = "test") -> Dict[str, Any]:
    """eval_caco_torch.py:343-408 (same arguments and printed output; additionally returns the metrics)."""
    filepaths, descriptions, _ = dataprocessor.get_filepaths_and_descriptions(current_split=eval_split)
    names, caps = [], []
    for fp in filepaths:
        name = fp.split("/")[-1].split(".wav")[0]
        names.append(name)
        caps.append(list(descriptions[name]["description"]))
    a = embed_files(model, filepaths, dataprocessor.config.sampling_rate, datasetconfig)
    return audio_retrieval_arrays(model, None, names, caps, tokenizer, datasetconfig, audio_embeddings=a)
