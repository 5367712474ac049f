"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path (cacophony_b200/dist.py) — shard ranges, the
single all-gather of both modalities, and that the row blocks assemble into the single-process similarity matrix."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cacophony_b200 import dist as cdist


def test_shard_range_covers_everything_once():
    for n, w in [(2048, 8), (400, 8), (7, 3), (5, 8), (256, 1)]:
        got = []
        for r in range(w):
            lo, hi = cdist.shard_range(n, r, w)
            assert 0 <= lo <= hi <= n
            got += list(range(lo, hi))
        assert got == list(range(n))
        sizes = [cdist.shard_range(n, r, w)[1] - cdist.shard_range(n, r, w)[0] for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


class _FakeModel:
    """similarity() with the reference formula (caco.py:208-210) in torch — stands in for the CUDA kernel on CPU."""
    scale = float(np.exp(2.6592))

    def similarity(self, a, t, want_ta=True):
        at = (self.scale * a) @ t.t()
        return at, ((self.scale * t) @ a.t() if want_ta else None)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        B, D = 6, 16
        a_all = torch.nn.functional.normalize(torch.randn(world * B, D, generator=g), dim=-1)
        t_all = torch.nn.functional.normalize(torch.randn(world * B, D, generator=g), dim=-1)
        lo, hi = cdist.shard_range(world * B, rank, world)
        ga, gt = cdist.gather_embeddings(a_all[lo:hi].contiguous(), t_all[lo:hi].contiguous())
        ok = torch.equal(ga, a_all) and torch.equal(gt, t_all)
        at_blk, ta_blk = cdist.sharded_contrastive_logits(_FakeModel(), a_all[lo:hi].contiguous(), t_all[lo:hi].contiguous())
        full_at, full_ta = _FakeModel().similarity(a_all, t_all)
        ok = ok and torch.allclose(at_blk, full_at[lo:hi], atol=1e-6) and torch.allclose(ta_blk, full_ta[lo:hi], atol=1e-6)
        # uneven row blocks (config #5: 400 clips over 8 ranks is even, 50 prompts / 7 clips are not)
        for n_total in (7, 2, 1, 10):
            full = torch.arange(n_total * 3, dtype=torch.int32).reshape(n_total, 3)
            rlo, rhi = cdist.shard_range(n_total, rank, world)
            ok = ok and torch.equal(cdist.gather_rows(full[rlo:rhi].contiguous(), n_total), full)
        q.put((rank, bool(ok)))
    except Exception as e:          # surface the failure instead of letting the parent time out
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_gather_and_row_blocks_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_gather_is_identity_without_process_group():
    a, t = torch.randn(3, 8), torch.randn(3, 8)
    ga, gt = cdist.gather_embeddings(a, t)
    assert ga is a and gt is t
    with pytest.raises(ValueError):
        cdist.gather_embeddings(a, torch.randn(4, 8))
