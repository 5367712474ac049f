"""Throughput of the BASELINE.json configs that are not the headline bench line, one GPU, CUDA events:
  config 2: 256 x 10 s clips, frontend + audio tower only            -> clips/s
  config 5: 400 x 5 s clips (248 valid tokens) + 50 prompts (T=100)  -> clips/s for the whole zero-shot pass (towers +
            logits + device top-1), with and without padding trim
Prints one JSON line per config (evidence for profiles/, not the driver's bench contract)."""
import json
import sys

import torch

sys.path.insert(0, ".")
import bench
import cacophony_b200 as cb
from cacophony_b200 import eval as ev

torch.manual_seed(0)
model = cb.create_caco_model().cuda()


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


wave, ids, mask = [t.cuda() for t in bench.synth_inputs(256, 0)]
ms = timed(lambda: model.encode_audio(wave, max_patches=500))
print(json.dumps({"config": 2, "what": "256 x 10 s clips, frontend + AudioMAE-ViT tower + pooler", "ms": round(ms, 3),
                  "clips_per_s": round(256 / ms * 1e3, 1), "tflops": round(256 * 95.53e9 / ms / 1e9, 1),
                  "frac_of_sustained_peak": round(256 * 95.53e9 / ms / 1e9 / 1387.6, 4)}), flush=True)
ms = timed(lambda: model.encode_text(ids, mask))
print(json.dumps({"config": "text", "what": "256 x 32-token captions, RoBERTa tower + pooler + projection", "ms": round(ms, 3),
                  "captions_per_s": round(256 / ms * 1e3, 1)}), flush=True)

g = torch.Generator().manual_seed(5)
clips = (0.1 * (2 * torch.rand(400, 80000, generator=g) - 1)).cuda()
cls_ids = torch.randint(3, 50265, (50, 100), generator=g)
cls_mask = torch.zeros(50, 100, dtype=torch.int64)
for i in range(50):
    n = 8 + i % 5
    cls_ids[i, 0], cls_ids[i, n - 1], cls_ids[i, n:] = 0, 2, 1
    cls_mask[i, :n] = 1
cls_ids, cls_mask = cls_ids.cuda(), cls_mask.cuda()
for trim in (False, True):
    def zs():
        t = model.encode_text(cls_ids, cls_mask)
        a = torch.cat([model.encode_audio(clips[i:i + 200], max_patches=500, trim_padding=trim) for i in (0, 200)])
        return ev.zero_shot_topk(model, a, t, 1)
    ms = timed(zs, iters=5)
    print(json.dumps({"config": 5, "what": "400 x 5 s clips + 50 prompts (T=100): towers + logits [400,50] + device top-1",
                      "trim_padding": trim, "tokens_computed_per_clip": 248 if trim else 500, "ms": round(ms, 3),
                      "clips_per_s": round(400 / ms * 1e3, 1)}), flush=True)
