"""L2-blocking sweep: the audio tower of the bench batch (256 x 10 s clips) run in passes of `chunk` clips, so that a pass's
activations (x fp32 3 KB/row, h16 1.5, qkv16 4.5, att16 1.5, mlp16 6 KB/row) stay in the 126 MB L2 between producer and
consumer kernels instead of round-tripping through HBM.  Interleaved rounds on one box; CUDA events."""
import json
import sys

import torch

sys.path.insert(0, ".")
import bench
import cacophony_b200 as cb

torch.manual_seed(0)
model = cb.create_caco_model().cuda()
B = 256
wave, ids, mask = [t.cuda() for t in bench.synth_inputs(B, 0)]
chunks = [int(a) for a in sys.argv[1:]] or [256, 128, 74, 64, 48, 37, 32, 24, 16]
ref = None
res = {c: {"audio": [], "pairs": []} for c in chunks}
for rnd in range(3):
    for c in chunks:
        model.set_option("audio_chunk_rows", c * 500)
        for what in ("audio", "pairs"):
            fn = (lambda: model.encode_audio(wave, max_patches=500)) if what == "audio" else \
                 (lambda: model.similarity(*model.encode_pairs(wave, ids, mask, max_patches=500)))
            for _ in range(2):
                out = fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                out = fn()
            e1.record()
            torch.cuda.synchronize()
            res[c][what].append(round(e0.elapsed_time(e1) / 5, 3))
            if what == "audio":
                if ref is None:
                    ref = out.clone()
                assert torch.equal(out, ref), "chunking changed the embeddings"
    print(json.dumps({"round": rnd, **{str(c): res[c] for c in chunks}}), flush=True)
best = min(chunks, key=lambda c: min(res[c]["pairs"]))
print(json.dumps({"best_chunk_clips": best, "ms_pairs": {str(c): min(res[c]["pairs"]) for c in chunks},
                  "ms_audio": {str(c): min(res[c]["audio"]) for c in chunks}}))
