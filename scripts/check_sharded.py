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
torch.cuda.synchronize()
ok = True
if rank == 0:
    a_all, t_all = model.encode_pairs(wave.cuda(), ids.cuda(), mask.cuda(), max_patches=MAX_PATCHES)
    at, ta = model.similarity(a_all, t_all)
    d1 = (at[lo:hi] - at_blk).abs().max().item()
    d2 = (ta[lo:hi] - ta_blk).abs().max().item()
    ok = d1 < 1e-4 and d2 < 1e-4 and tuple(at_blk.shape) == (hi - lo, B)
    print(f"sharded vs single-GPU: max|d at| = {d1:.2e}, max|d ta| = {d2:.2e}, block {tuple(at_blk.shape)} -> {'OK' if ok else 'FAIL'}")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
