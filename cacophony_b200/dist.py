"""Multi-GPU plumbing for the one exchange step of the path: the batch of audio-text pairs is sharded
across ranks (independent units, weights replicated), every rank encodes its shard, the L2-normalised embeddings
([B_local, 768] fp32 per modality and rank: 0.79 MB at B_local = 256) are all-gathered over NVLink/NVSwitch — the text
side as soon as the text tower is done, hidden under the audio tower — and each rank computes its row block of the global
similarity matrix (caco.py:208-210).  One process per GPU,
``torch.distributed`` (NCCL on GPUs; gloo on CPU for the host-logic tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row range [lo, hi) of rank `rank` when n_total units are split as evenly as possible."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


_RECV: dict = {}


def _recv_buffer(key, shape, like: torch.Tensor) -> torch.Tensor:
    """Persistent receive buffers (one per role and shape): the collective writes straight into memory the similarity
    kernel reads, and no allocation sits between the towers and the exchange."""
    k = (key, tuple(shape), like.device, like.dtype)
    buf = _RECV.get(k)
    if buf is None:
        buf = torch.empty(shape, dtype=like.dtype, device=like.device)
        _RECV[k] = buf
    return buf


def gather_embedding(e_local: torch.Tensor, group: Optional[dist.ProcessGroup] = None, key: str = "emb",
                     async_op: bool = False):
    """All-gather ONE modality's L2-normalised embeddings [B_local, D] (contiguous: the tower's own output buffer is the
    send buffer — K7 of SURVEY.md 2 writes into it directly) into a persistent [world * B_local, D] receive buffer,
    rank-major.  Returns the buffer, or (buffer, work) when async_op."""
    if e_local.dim() != 2 or not e_local.is_contiguous():
        raise ValueError("gather_embedding: contiguous [B_local, D] expected")
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return (e_local, None) if async_op else e_local
    world = dist.get_world_size(group)
    recv = _recv_buffer(key, (world * e_local.shape[0], e_local.shape[1]), e_local)
    work = dist.all_gather_into_tensor(recv, e_local, group=group, async_op=async_op)
    return (recv, work) if async_op else recv


def gather_embeddings(a_local: torch.Tensor, t_local: torch.Tensor, group: Optional[dist.ProcessGroup] = None
                      ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Both modalities: a_local, t_local [B_local, D] (same B_local on every rank) -> (A_all, T_all) [world * B_local, D],
    rank-major (rank r owns rows r*B_local ...).  Two collectives of 0.79 MB each at B_local = 256, text first (the order
    every rank issues them in)."""
    if a_local.shape != t_local.shape or a_local.dim() != 2:
        raise ValueError("gather_embeddings: a_local and t_local must both be [B_local, D]")
    t_all = gather_embedding(t_local.contiguous(), group, "text")
    a_all = gather_embedding(a_local.contiguous(), group, "audio")
    return a_all, t_all


def sharded_contrastive_logits(model, a_local: torch.Tensor, t_local: torch.Tensor,
                               group: Optional[dist.ProcessGroup] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Row blocks of the global logits: at[r-block, :] = s·A_local·T_allᵀ and ta[r-block, :] = s·T_local·A_allᵀ."""
    a_all, t_all = gather_embeddings(a_local, t_local, group)
    at_block, _ = model.similarity(a_local, t_all, want_ta=False)
    ta_block, _ = model.similarity(t_local, a_all, want_ta=False)
    return at_block, ta_block


def sharded_pairs_logits(model, waveform: torch.Tensor, text_input_ids: torch.Tensor, text_mask: torch.Tensor,
                         max_patches: int = 500, group: Optional[dist.ProcessGroup] = None
                         ) -> Tuple[torch.Tensor, torch.Tensor]:
    """The whole sharded step (BASELINE config 4) with the exchange hidden: the text tower runs on the model's side stream
    and its embeddings are all-gathered as soon as it finishes (about 3 ms into a 32 ms step, i.e. entirely under the audio
    tower); only the audio embeddings' gather and the two row-block similarity launches follow the audio tower.  Every rank
    issues the two collectives in the same order (text, audio).  Returns this rank's (at_block, ta_block)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        a, t = model.encode_pairs(waveform, text_input_ids, text_mask, max_patches=max_patches)
        return model.similarity(a, t)
    dev = model._device()
    cur = torch.cuda.current_stream(dev)
    side = model.side_stream()
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        t_local = model.encode_text(text_input_ids.to(dev), text_mask.to(dev))
        t_all = gather_embedding(t_local, group, "text")
    a_local = model.encode_audio(waveform, max_patches=max_patches)
    at_block, _ = None, None
    cur.wait_stream(side)
    t_local.record_stream(cur)
    at_block, _ = model.similarity(a_local, t_all, want_ta=False)          # needs only the (already gathered) text side
    a_all = gather_embedding(a_local, group, "audio")
    ta_block, _ = model.similarity(t_local, a_all, want_ta=False)
    return at_block, ta_block


def gather_rows(x_local: torch.Tensor, n_total: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All-gather row blocks of UNEVEN height (rank r holds rows shard_range(n_total, r, world) of a [n_total, ...] tensor):
    blocks are padded to the tallest shard for the one collective and trimmed afterwards."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x_local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_range(n_total, rank, world)
    if x_local.shape[0] != hi - lo:
        raise ValueError(f"gather_rows: rank {rank} should hold {hi - lo} rows, got {x_local.shape[0]}")
    tallest = -(-n_total // world)
    send = x_local.new_zeros((tallest,) + tuple(x_local.shape[1:]))
    send[: hi - lo] = x_local
    recv = x_local.new_empty((world * tallest,) + tuple(x_local.shape[1:]))
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    parts = []
    for r in range(world):
        rlo, rhi = shard_range(n_total, r, world)
        parts.append(recv[r * tallest: r * tallest + (rhi - rlo)])
    return torch.cat(parts, dim=0)


def sharded_zero_shot_topk(model, waves, all_text_embeddings: torch.Tensor, k: int = 1, datasetconfig=None,
                           group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """BASELINE config #5: clips sharded across ranks (rank r encodes waves[shard_range(...)]), the class
    text embeddings replicated (every rank computed them itself: no collective), each rank ranks its own clips on the
    device, and only the int32 [n_clips, k] predictions are gathered.  `waves` is the FULL list on every rank."""
    from . import eval as ev
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    lo, hi = shard_range(len(waves), rank, world)
    if hi > lo:
        a = ev.embed_waveforms(model, waves[lo:hi], datasetconfig)
        top = ev.zero_shot_topk(model, a, all_text_embeddings, k)
    else:
        top = torch.empty((0, k), dtype=torch.int32, device=all_text_embeddings.device)
    return gather_rows(top, len(waves), group)


def sharded_retrieval_topk(model, a_local: torch.Tensor, t_local: torch.Tensor, n_audio: int, n_text: int, k: int = 10,
                           group: Optional[dist.ProcessGroup] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Retrieval rankings with both embedding sets sharded by rows: two uneven all-gathers of the embeddings, then every
    rank ranks ITS queries against all keys (audio -> text for its clips, text -> audio for its captions) and the int32
    rankings are gathered.  Returns the full (at_idx [n_audio, k], ta_idx [n_text, k]) on every rank."""
    from . import ops
    a_all = gather_rows(a_local, n_audio, group)
    t_all = gather_rows(t_local, n_text, group)
    at = ops.topk_rows(ops.sgemm_nt(a_local.contiguous(), t_all.contiguous()), k) if a_local.shape[0] else \
        torch.empty((0, k), dtype=torch.int32, device=a_local.device)
    ta = ops.topk_rows(ops.sgemm_nt(t_local.contiguous(), a_all.contiguous()), k) if t_local.shape[0] else \
        torch.empty((0, k), dtype=torch.int32, device=t_local.device)
    return gather_rows(at, n_audio, group), gather_rows(ta, n_text, group)
