"""Per-request latency of the whole path at small batches: eager launches vs one CUDA-graph replay (CUDA events)."""
import json
import sys

import torch

sys.path.insert(0, ".")
import bench
import cacophony_b200 as cb
from cacophony_b200.serving import GraphedPairs

torch.manual_seed(0)
model = cb.create_caco_model().cuda()


def timed(fn, iters=50, warm=5):
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


for B in (1, 4, 16, 64):
    wave, ids, mask = [t.cuda() for t in bench.synth_inputs(B, 0)]

    def eager():
        a, t = model.encode_pairs(wave, ids, mask, max_patches=500)
        return model.similarity(a, t)
    ref = eager()[0].clone()
    g = GraphedPairs(model, B, wave.shape[1], ids.shape[1])
    out = g(wave, ids, mask)[0]
    torch.cuda.synchronize()
    same = bool(torch.equal(out, ref))
    ms_e, ms_g = timed(eager), timed(lambda: g(wave, ids, mask))
    print(json.dumps({"pairs": B, "eager_ms": round(ms_e, 3), "graph_ms": round(ms_g, 3), "graph_equals_eager": same,
                      "pairs_per_s_graph": round(B / ms_g * 1e3, 1)}), flush=True)
