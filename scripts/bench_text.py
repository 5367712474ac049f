"""Text tower alone at the bench shape (256 captions x 32 tokens) and the full step, for A/B runs of the small-GEMM tile choice."""
import json
import sys

import torch

sys.path.insert(0, ".")
import bench
import cacophony_b200 as cb

torch.manual_seed(0)
model = cb.create_caco_model().cuda()
wave, ids, mask = [t.cuda() for t in bench.synth_inputs(256, 0)]
arms = {"auto": 0, "cg1_n256": 1}
res = {k: {"text": [], "pairs": []} for k in arms}
for rnd in range(3):
    for name, v in arms.items():
        model.set_option("gemm_variant", 0)
        for what in ("text", "pairs"):
            if what == "text":
                model.set_option("gemm_variant", v)            # forced variant applies to every GEMM: text tower only
                fn = lambda: model.encode_text(ids, mask)
            else:
                model.set_option("gemm_variant", 0)
                fn = lambda: model.similarity(*model.encode_pairs(wave, ids, mask, max_patches=500))
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res[name][what].append(round(e0.elapsed_time(e1) / 10, 3))
print(json.dumps(res))
