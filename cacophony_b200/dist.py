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
