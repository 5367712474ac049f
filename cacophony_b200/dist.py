"""Multi-GPU plumbing for the one exchange step of the path: the batch of audio-text pairs is sharded
across ranks (independent units, weights replicated), every rank encodes its shard, ONE all-gather moves the
L2-normalised embeddings ([B_local, 2, 768] fp32 per rank: 1.57 MB at B_local = 256) over NVLink/NVSwitch, and each
rank computes its row block of the global similarity matrix (caco.py:208-210).  One process per GPU,
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


def gather_embeddings(a_local: torch.Tensor, t_local: torch.Tensor, group: Optional[dist.ProcessGroup] = None
                      ) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather both modalities with a single collective.  a_local, t_local: [B_local, D] (same B_local on every
    rank).  Returns (A_all, T_all): [world * B_local, D] each, rank-major (rank r owns rows r*B_local ...)."""
    if a_local.shape != t_local.shape or a_local.dim() != 2:
        raise ValueError("gather_embeddings: a_local and t_local must both be [B_local, D]")
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return a_local, t_local
    world = dist.get_world_size(group)
    send = torch.stack([a_local, t_local], dim=1).contiguous()                 # [B, 2, D]
    B, _, D = send.shape
    recv = torch.empty((world * B, 2, D), dtype=send.dtype, device=send.device)      # rank-major concatenation
    dist.all_gather_into_tensor(recv, send, group=group)
    return recv[:, 0, :].contiguous(), recv[:, 1, :].contiguous()


def sharded_contrastive_logits(model, a_local: torch.Tensor, t_local: torch.Tensor,
                               group: Optional[dist.ProcessGroup] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Row blocks of the global logits: at[r-block, :] = s·A_local·T_allᵀ and ta[r-block, :] = s·T_local·A_allᵀ."""
    a_all, t_all = gather_embeddings(a_local, t_local, group)
    at_block, _ = model.similarity(a_local, t_all.contiguous(), want_ta=False)
    ta_block, _ = model.similarity(t_local, a_all.contiguous(), want_ta=False)
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
