"""torchrun --nproc-per-node N scripts/check_sharded.py : the sharded path (config 4 shape, scaled down) must reproduce
the single-GPU logits on the same inputs: every rank encodes its shard, all-gathers, computes its row block; rank 0 also
runs the whole batch alone and compares."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import cacophony_b200 as cb
from cacophony_b200 import dist as cdist
from bench import synth_inputs, MAX_PATCHES

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.manual_seed(0)
model = cb.create_caco_model().to("cuda")
B = 8 * world
wave, ids, mask = synth_inputs(B, 7)
lo, hi = cdist.shard_range(B, rank, world)
a, t = model.encode_pairs(wave[lo:hi].cuda(), ids[lo:hi].cuda(), mask[lo:hi].cuda(), max_patches=MAX_PATCHES)
at_blk, ta_blk = cdist.sharded_contrastive_logits(model, a, t)
# the whole sharded step with the text gather hidden under the audio tower: must equal the plain form bit for bit
ok = True
for peer in (False, True, True, True):          # NCCL exchange, then the peer-memory exchange three steps in a row (both parities)
    at_blk2, ta_blk2 = cdist.sharded_pairs_logits(model, wave[lo:hi].cuda(), ids[lo:hi].cuda(), mask[lo:hi].cuda(),
                                                  max_patches=MAX_PATCHES, use_peer_memory=peer)
    torch.cuda.synchronize()
    same = torch.equal(at_blk, at_blk2) and torch.equal(ta_blk, ta_blk2)
    if not same:
        print(f"rank {rank}: sharded_pairs_logits(use_peer_memory={peer}) differs from sharded_contrastive_logits: "
              f"{(at_blk - at_blk2).abs().max().item():.3e} {(ta_blk - ta_blk2).abs().max().item():.3e}")
    ok = ok and same
ex = cdist.peer_exchange(model, hi - lo)
if ex is not None:
    # the software-pipelined form: logits come out one call later and must be the same blocks, bit for bit, over 6 steps
    pipe = cdist.PipelinedPairs(model, hi - lo, MAX_PATCHES)
    outs = [pipe.step(wave[lo:hi].cuda(), ids[lo:hi].cuda(), mask[lo:hi].cuda()) for _ in range(6)] + [pipe.flush()]
    torch.cuda.synchronize()
    same = outs[0] is None and all(torch.equal(o[0], at_blk) and torch.equal(o[1], ta_blk) for o in outs[1:])
    if not same:
        print(f"rank {rank}: PipelinedPairs differs from sharded_contrastive_logits")
    ok = ok and same
if rank == 0:
    print("exchange:", "peer-memory stores fused into the L2-norm kernel" if ex is not None else "NCCL all-gather (peer memory unavailable)")
if ex is not None:
    ex.check()
if rank == 0:
    a_all, t_all = model.encode_pairs(wave.cuda(), ids.cuda(), mask.cuda(), max_patches=MAX_PATCHES)
    at, ta = model.similarity(a_all, t_all)
    d1 = (at[lo:hi] - at_blk).abs().max().item()
    d2 = (ta[lo:hi] - ta_blk).abs().max().item()
    ok = ok and d1 < 1e-4 and d2 < 1e-4 and tuple(at_blk.shape) == (hi - lo, B)
    print(f"sharded vs single-GPU: max|d at| = {d1:.2e}, max|d ta| = {d2:.2e}, block {tuple(at_blk.shape)} -> {'OK' if ok else 'FAIL'}")

# ---- config #5 shape (scaled down): 5 s clips sharded, class prompts (T = 100) replicated, predictions gathered
import numpy as np
from cacophony_b200 import eval as ev
from cacophony_b200 import ops
g = torch.Generator().manual_seed(5)
n_clips, n_cls = 6 * world + 1, 50                     # uneven shard on purpose
clips = [(0.1 * (2 * torch.rand(80000, generator=g) - 1)).numpy() for _ in range(n_clips)]
cls_ids = torch.randint(3, 50265, (n_cls, 100), generator=g)
cls_mask = torch.zeros(n_cls, 100, dtype=torch.int64)
for i in range(n_cls):
    n = 8 + i % 5
    cls_ids[i, 0], cls_ids[i, n - 1], cls_ids[i, n:] = 0, 2, 1
    cls_mask[i, :n] = 1
t_cls = ev.embed_text_ids(model, cls_ids.cuda(), cls_mask.cuda())
top = cdist.sharded_zero_shot_topk(model, clips, t_cls, k=1)
a_full = ev.embed_waveforms(model, clips)
ref_top = ev.zero_shot_topk(model, a_full, t_cls, 1)
ok5 = tuple(top.shape) == (n_clips, 1) and torch.equal(top, ref_top)
# ---- sharded retrieval rankings == single-GPU rankings
lo_a, hi_a = cdist.shard_range(n_clips, rank, world)
lo_t, hi_t = cdist.shard_range(n_cls, rank, world)
at_idx, ta_idx = cdist.sharded_retrieval_topk(model, a_full[lo_a:hi_a].contiguous(), t_cls[lo_t:hi_t].contiguous(), n_clips, n_cls, 10)
ref_at, ref_ta = ev.retrieval_topk(t_cls, a_full, 10)
okr = torch.equal(at_idx, ref_at) and torch.equal(ta_idx, ref_ta)
flag = torch.tensor([int(ok5), int(okr)], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"config-5 sharded zero-shot top-1 ({n_clips} clips x {n_cls} prompts over {world} ranks) == single GPU: {bool(flag[0])}; "
          f"sharded retrieval rankings == single GPU: {bool(flag[1])}")
ok = ok and bool(flag[0]) and bool(flag[1])
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
