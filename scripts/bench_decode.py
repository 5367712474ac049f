"""Captioning (row f-4): greedy decode throughput — audio tower once, then per generated token either one get_decoder_logits call
on the whole prefix (text tower + 4 decoder layers + 50 265-way projection, as the reference's decode loop does: "full_prefix"),
or one KV-cached step on the newest token ("kv_cache"), eager or replayed from a CUDA graph ("kv_cache_graph")."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import cacophony_b200 as cb
from cacophony_b200 import eval as ev
from cacophony_b200 import serving

torch.manual_seed(0)
model = cb.create_caco_model().cuda()
ARMS = {"full_prefix": dict(use_cache=False), "kv_cache": dict(use_cache=True), "kv_cache_graph": dict(use_cache=True, use_graph=True)}
for B, steps in ((1, 32), (8, 32), (64, 32)):
    w = (0.1 * (2 * torch.rand(B, 160000, device="cuda") - 1)).float()
    ab = cb.prepare_audio_batch(w, cb.DatasetConfig(patches_seq_len=500), "cuda")
    outs = {}
    stepper = serving.GraphedDecodeStep(model, B, 500, steps)        # captured once per request shape, reused across requests
    for arm, kw in ARMS.items():
        if "use_graph" in kw:
            kw = dict(kw, use_graph=stepper)
        ev.decode_caption_ids(model, ab, eos_id=-1, max_decode_length=4, **kw)          # warm-up (eos never hit: fixed length)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = ev.decode_caption_ids(model, ab, eos_id=-1, max_decode_length=steps, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        outs[arm] = out
        print(json.dumps({"arm": arm, "batch": B, "generated_tokens_per_clip": steps, "s": round(dt, 4),
                          "tokens_per_s": round(B * steps / dt, 1), "ms_per_step": round(1e3 * dt / steps, 3),
                          "same_tokens_as_full_prefix": bool(torch.equal(out, outs["full_prefix"]))}), flush=True)
