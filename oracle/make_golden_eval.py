"""Generate tests/golden/eval_retrieval.npz by running the REFERENCE's own compute_retrieval_metric
(src/eval/eval_utils.py:18-66) on the synthetic cases of oracle/eval_oracle.make_retrieval_case (authoring container only).

astropy (jackknife) and soundfile are not installed here, so they are stubbed before the import; the stub jackknife records
the per-query arrays the reference passes to it — those arrays ARE the reference's R@1/5/10 and AP@10 — and returns
placeholders for the printed interval.

Usage:  python oracle/make_golden_eval.py
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import eval_oracle as E  # noqa: E402

CASES = [("plain", 5, 40, 5, 0), ("dups", 6, 33, 4, 7), ("one_cap", 7, 25, 1, 0)]


def main():
    captured = []
    astropy = types.ModuleType("astropy")
    stats = types.ModuleType("astropy.stats")

    class _JK:
        @staticmethod
        def jackknife_stats(data, statistic, conf):
            captured.append(np.asarray(data, dtype=np.float64).copy())
            return 0.0, 0.0, 0.0, (0.0, 0.0)
    stats.jackknife = _JK
    astropy.stats = stats
    sys.modules["astropy"], sys.modules["astropy.stats"] = astropy, stats
    sys.modules["soundfile"] = types.ModuleType("soundfile")
    sys.path.insert(0, "/root/reference")
    from src.eval.eval_utils import compute_retrieval_metric as ref_metric

    out = {}
    for name, seed, n_audio, caps, dup in CASES:
        names, all_text, gt_at, gt_ta, at_idx, ta_idx = E.make_retrieval_case(seed, n_audio, caps, dup)
        for kind, idx, qs, ks, gt in (("at", at_idx, names, all_text, gt_at), ("ta", ta_idx, all_text, names, gt_ta)):
            captured.clear()
            with contextlib.redirect_stdout(io.StringIO()):
                ref_metric(idx, qs, ks, gt, kind)
            assert len(captured) == 4
            for metric, arr in zip(("R1", "R5", "R10", "mAP10"), captured):
                out[f"{name}/{kind}/{metric}"] = arr
    path = os.path.join(ROOT, "tests", "golden", "eval_retrieval.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()
