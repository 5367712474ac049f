"""A/B timing of the full bench step under library switches, interleaved in ONE process on ONE box (box-to-box and
run-to-run spread is ~3 %, larger than most single-kernel effects).
Usage: python scripts/ab_step.py name=option:value[,option:value] ...   e.g.  red=resid_red:1 nored=resid_red:0
(options: the per-model execution options of include/caco_b200.h, set with model.set_option)"""
import json
import sys

import torch

sys.path.insert(0, ".")
import bench
import cacophony_b200 as cb
from cacophony_b200 import _lib as L

lib = L.load()
torch.manual_seed(0)
model = cb.create_caco_model().cuda()
import os
BATCH = int(os.environ.get("AB_BATCH", "256"))
wave, ids, mask = [t.cuda() for t in bench.synth_inputs(BATCH, 0)]
arms = []
for a in sys.argv[1:]:
    name, spec = a.split("=")
    arms.append((name, [(s.split(":")[0], int(s.split(":")[1])) for s in spec.split(",")]))


def step():
    a, t = model.encode_pairs(wave, ids, mask, max_patches=bench.MAX_PATCHES)
    return model.similarity(a, t)


res = {n: [] for n, _ in arms}
for rnd in range(4):
    for name, sets in arms:
        for opt, v in sets:
            model.set_option(opt, v)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            step()
        e1.record()
        torch.cuda.synchronize()
        res[name].append(round(e0.elapsed_time(e1) / 10, 3))
print(json.dumps(res))
