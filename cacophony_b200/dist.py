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
                         max_patches: int = 500, group: Optional[dist.ProcessGroup] = None, use_peer_memory: bool = True
                         ) -> Tuple[torch.Tensor, torch.Tensor]:
    """The whole sharded step (BASELINE config 4) with the exchange hidden.  The text tower runs on the model's side stream
    and its embeddings reach every rank as soon as it finishes (about 3 ms into a 32 ms step, i.e. entirely under the audio
    tower); only the audio embeddings' exchange and the two row-block similarity launches follow the audio tower.
    use_peer_memory (default): the exchange is fused into the towers' final L2-normalisation kernel, which stores the rows
    into every peer's gathered matrix over NVLink (`PeerExchange`); otherwise — or when symmetric memory is unavailable —
    two NCCL all-gathers, issued by every rank in the same order (text, audio).  Same logits bit for bit either way.
    Returns this rank's (at_block, ta_block)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        a, t = model.encode_pairs(waveform, text_input_ids, text_mask, max_patches=max_patches)
        return model.similarity(a, t)
    dev = model._device()
    cur = torch.cuda.current_stream(dev)
    side = model.side_stream()
    ids, msk = text_input_ids.to(dev), text_mask.to(dev)
    ex = peer_exchange(model, int(ids.shape[0]), group) if use_peer_memory else None
    side.wait_stream(cur)
    if ex is not None:
        # exchange fused into the towers' last kernel: normalised rows are stored into every peer's matrix from the kernel
        with torch.cuda.stream(side):
            t_raw = model.encode_text(ids, msk, normalize=False)
            ex.scatter(t_raw, PeerExchange.TEXT)
        a_raw = model.encode_audio(waveform, max_patches=max_patches, normalize=False)
        ex.scatter(a_raw, PeerExchange.AUDIO)
        cur.wait_stream(side)
        for x in (ids, msk, t_raw):
            x.record_stream(cur)
        at_block, _ = model.similarity(ex.local_rows(PeerExchange.AUDIO), ex.gathered(PeerExchange.TEXT), want_ta=False)
        ta_block, _ = model.similarity(ex.local_rows(PeerExchange.TEXT), ex.gathered(PeerExchange.AUDIO), want_ta=False)
        return at_block, ta_block
    with torch.cuda.stream(side):
        t_local = model.encode_text(ids, msk)
        t_all = gather_embedding(t_local, group, "text")
    a_local = model.encode_audio(waveform, max_patches=max_patches)
    cur.wait_stream(side)
    for x in (ids, msk, t_local):
        x.record_stream(cur)
    at_block, _ = model.similarity(a_local, t_all, want_ta=False)          # needs only the (already gathered) text side
    a_all = gather_embedding(a_local, group, "audio")
    ta_block, _ = model.similarity(t_local, a_all, want_ta=False)
    return at_block, ta_block


class PeerExchange:
    """The path's one exchange step without a collective launch (SURVEY.md 2, K7 "writing into the all-gather buffer"): every
    rank owns one symmetric buffer (``torch.distributed._symmetric_memory``: CUDA VMM allocations mapped into every peer over
    NVLink) holding, for `slots` consecutive steps and two modalities, the gathered ``[world * B, D]`` embedding matrix plus
    one flag per (modality, rank).  ``scatter`` L2-normalises a rank's raw embeddings and stores the rows straight into every
    peer's matrix from inside the kernel (``caco_l2norm_scatter``), ``gathered`` orders the caller's stream behind a 1-CTA
    wait on the flags (``caco_wait_flags``) and returns the local matrix.  No NCCL kernel runs next to the towers.

    One instance per (model device, B_local, D, group); construct it on every rank at the same time (rendezvous)."""

    AUDIO, TEXT = 0, 1

    def __init__(self, device: torch.device, b_local: int, dim: int, group: Optional[dist.ProcessGroup] = None,
                 slots: int = 4):
        import torch.distributed._symmetric_memory as symm
        from . import _lib as L
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.B, self.D, self.device = int(b_local), int(dim), torch.device(device)
        self.mat_elems = self.world * self.B * self.D
        self.slots = int(slots)
        self.flag_off_elems = 2 * self.slots * self.mat_elems          # [slot][modality 2] matrices, then the flags
        n = self.flag_off_elems + 64                                   # [modality 2][world <= 16] uint32 flags (as fp32 slots)
        with torch.cuda.device(self.device):
            self.buf = symm.empty(n, dtype=torch.float32, device=self.device)
            self.buf.zero_()
            torch.cuda.synchronize(self.device)
            self.hdl = symm.rendezvous(self.buf, self.group)
            self.ticket = torch.zeros(2, dtype=torch.int32, device=self.device)      # one per modality / stream
            self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        dist.barrier(self.group)                                       # every rank's flags are zero before anyone writes
        self.peer_base_dev = int(self.hdl.buffer_ptrs_dev)
        self.step = [0, 0]                                             # per modality
        self._lib = L

    def _mat_off(self, modality: int, step: int) -> int:
        """Offset of the matrix that scatter number `step` (1-based, per modality) wrote."""
        return (((step - 1) % self.slots) * 2 + modality) * self.mat_elems

    def scatter(self, e_raw: torch.Tensor, modality: int) -> None:
        """e_raw [B_local, D] fp32 (UN-normalised embeddings of this rank) -> normalised rows in every rank's matrix."""
        L = self._lib
        if tuple(e_raw.shape) != (self.B, self.D) or e_raw.dtype != torch.float32 or not e_raw.is_contiguous():
            raise ValueError("PeerExchange.scatter: contiguous fp32 [B_local, D] expected")
        self.step[modality] += 1
        dst = (self._mat_off(modality, self.step[modality]) + self.rank * self.B * self.D) * 4
        with torch.cuda.device(self.device):
            L.check(L.load().caco_l2norm_scatter(L.ptr(e_raw), self.B, self.D, 1e-10, self.peer_base_dev, dst,
                                                 self.flag_off_elems * 4, modality * 16 + self.rank, self.step[modality],
                                                 self.world, self.ticket.data_ptr() + 4 * modality, L.stream_ptr()),
                    "caco_l2norm_scatter")

    def gathered(self, modality: int, step: Optional[int] = None) -> torch.Tensor:
        """The local [world * B, D] matrix of the modality's scatter number `step` (default: the last one), valid for work
        enqueued on the current stream after this call (a 1-CTA kernel waits for every rank's flag first).  Only the last
        `slots - 2` steps are guaranteed not to have been overwritten by a faster peer."""
        L = self._lib
        step = self.step[modality] if step is None else step
        off = self._mat_off(modality, step)
        with torch.cuda.device(self.device):
            flags_ptr = self.buf.data_ptr() + (self.flag_off_elems + modality * 16) * 4
            L.check(L.load().caco_wait_flags(flags_ptr, self.world, step, 10000, L.ptr(self.status), L.stream_ptr()),
                    "caco_wait_flags")
        return self.buf[off: off + self.mat_elems].view(self.world * self.B, self.D)

    def local_rows(self, modality: int, step: Optional[int] = None) -> torch.Tensor:
        """This rank's own normalised rows of scatter number `step` (written by its own kernel: stream order suffices)."""
        step = self.step[modality] if step is None else step
        off = self._mat_off(modality, step) + self.rank * self.B * self.D
        return self.buf[off: off + self.B * self.D].view(self.B, self.D)

    def check(self) -> None:
        """Raises if a wait kernel timed out (a peer never published); synchronises the device."""
        s = int(self.status.item())
        if s:
            raise RuntimeError(f"PeerExchange: rank {s - 1} did not publish its embeddings within the time-out")


class PipelinedPairs:
    """BASELINE config 4 as a software pipeline: step k enqueues both towers and publishes this rank's embeddings, then
    computes the logits of step k - 1 — whose embeddings every peer published a whole step ago, so the flag wait never
    stalls.  Without the lag every step runs at the pace of the slowest of the N power-capped GPUs IN THAT STEP (measured on
    8 B200s: ranks alone 30.4-31.2 ms per step, coupled 31.8-32.6); with it a rank only ever waits for a peer that is a whole
    step behind.  Four buffer slots make the overlap safe: a rank passes the wait for step s only after every peer
    published s, and a peer publishes s + 1 only after it consumed s - 1, so the slot a rank rewrites at s + 3 is free.

    ``step(...)`` returns the (at_block, ta_block) of the PREVIOUS call (None the first time); ``flush()`` returns the last."""

    def __init__(self, model, b_local: int, max_patches: int = 500, group: Optional[dist.ProcessGroup] = None):
        self.model, self.max_patches, self.group = model, max_patches, group
        self.ex = peer_exchange(model, b_local, group)
        if self.ex is None:
            raise RuntimeError("PipelinedPairs needs the peer-memory exchange (symmetric memory unavailable on this box)")
        self.pending = 0                      # exchange step whose logits are still to be computed (0 = none)

    def _logits(self, s: int):
        ex = self.ex
        at_block, _ = self.model.similarity(ex.local_rows(ex.AUDIO, s), ex.gathered(ex.TEXT, s), want_ta=False)
        ta_block, _ = self.model.similarity(ex.local_rows(ex.TEXT, s), ex.gathered(ex.AUDIO, s), want_ta=False)
        return at_block, ta_block

    def step(self, waveform: torch.Tensor, text_input_ids: torch.Tensor, text_mask: torch.Tensor):
        model, ex = self.model, self.ex
        dev = model._device()
        cur, side = torch.cuda.current_stream(dev), model.side_stream()
        ids, msk = text_input_ids.to(dev), text_mask.to(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            t_raw = model.encode_text(ids, msk, normalize=False)
            ex.scatter(t_raw, ex.TEXT)
        a_raw = model.encode_audio(waveform, max_patches=self.max_patches, normalize=False)
        ex.scatter(a_raw, ex.AUDIO)
        cur.wait_stream(side)
        for x in (ids, msk, t_raw):
            x.record_stream(cur)
        out = self._logits(self.pending) if self.pending else None
        self.pending = ex.step[ex.AUDIO]
        return out

    def flush(self):
        if not self.pending:
            return None
        out = self._logits(self.pending)
        self.pending = 0
        return out


_EXCHANGES: dict = {}


def peer_exchange(model, b_local: int, group: Optional[dist.ProcessGroup] = None) -> Optional[PeerExchange]:
    """The cached PeerExchange of (model device, B_local), or None when symmetric memory is unavailable (then the NCCL
    all-gather path is used).  Must be called by every rank of the group with the same b_local."""
    dev = model._device()
    key = (dev, int(b_local), id(group))
    if key not in _EXCHANGES:
        try:
            _EXCHANGES[key] = PeerExchange(dev, b_local, model.audio_config.hidden_size, group)
        except Exception as e:                                  # no P2P / VMM support on this box: NCCL does the exchange
            import warnings
            warnings.warn(f"cacophony_b200.dist: peer-memory exchange unavailable ({type(e).__name__}: {e}); using NCCL all-gather")
            _EXCHANGES[key] = None
    return _EXCHANGES[key]


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
