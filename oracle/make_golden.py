"""Generate tests/golden/*.npz by running the REFERENCE itself (authoring container only).

Usage:  python oracle/make_golden.py            (needs /root/reference; CPU; ~1 min)

The reference is Python and cannot travel to the GPU box, so its outputs on the deterministic
synthetic weights/inputs of ``oracle/weights.py`` are committed as small fixtures.  Inputs are NOT
stored: tests regenerate them from the same seeds.  ``CASES`` below is the single source of truth for
the case definitions; ``tests/`` import it.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")

# frontend cases: (name, seed, kind, n_samples, max_patches)
FRONTEND_CASES = [
    ("fe_10s_noise", 11, "noise", 160000, 500),
    ("fe_5s_noise", 12, "noise", 80000, 500),
    ("fe_odd_noise", 13, "noise", 159999, 500),
    ("fe_short_noise", 14, "noise", 12345, 500),
    ("fe_10s_chirp", 15, "chirp", 160000, 500),
    ("fe_12s_trunc", 16, "noise", 192000, 500),      # 600 patches > 500 -> truncation branch
    ("fe_tiny", 17, "noise", 100, 16),               # < one patch row: 0 valid patches
]

# model cases: name -> dict(seed, sharp, clip lengths, caption lens, T, max_patches)
MODEL_CASES = {
    "model_s0": dict(seed=0, sharp=1.0, clip_lens=[160000, 160000, 80000, 12345], zero_tail=[0, 80000, 0, 0],
                     cap_lens=[32, 20, 9, 4], T=32, max_patches=500),
    "model_s1_sharp": dict(seed=1, sharp=4.0, clip_lens=[160000, 80000], zero_tail=[0, 0],
                           cap_lens=[32, 11], T=32, max_patches=500),
    "model_s2_t100": dict(seed=2, sharp=1.0, clip_lens=[80000, 80000, 80000], zero_tail=[0, 0, 0],
                          cap_lens=[8, 10, 12], T=100, max_patches=500),
    # checkpoint-like outliers (oracle/weights.py::_apply_outliers): LayerNorm gains x 30 in six channels, residual channels at
    # +-300, fc1 pre-activations near 1e3 — the case the fp16-operand scheme has to survive (or the split-weight mode)
    # with the captioning head (4 decoder layers, roberta.py:329-373): get_decoder_logits
    "model_s4_decoder": dict(seed=4, sharp=1.0, decoder_layers=4, clip_lens=[80000, 48000], zero_tail=[0, 0],
                             cap_lens=[24, 9], T=24, max_patches=256),
    "model_s3_outlier": dict(seed=3, sharp=1.0, outlier=True, clip_lens=[160000, 80000, 40000], zero_tail=[0, 0, 0],
                             cap_lens=[32, 13, 6], T=32, max_patches=500),
}


def case_inputs(c):
    """Deterministic inputs of a model case: (list of waveforms, ids, mask)."""
    from oracle import weights as W
    waves = []
    for i, (n, z) in enumerate(zip(c["clip_lens"], c["zero_tail"])):
        # noise only: a pure tone leaves mel bins at the fp32-FFT noise floor, where two correct
        # frontends legitimately differ (tests/util.py) and the embedding inherits ~2e-4 of that.
        w = W.make_waveforms(100 * c["seed"] + i, 1, n, "noise")[0]
        if z:
            w[z:] = 0.0
        waves.append(w)
    ids, mask = W.make_captions(c["seed"], len(c["cap_lens"]), c["T"], lens=c["cap_lens"])
    return waves, ids, mask


def _import_reference():
    sys.path.insert(0, "/root/reference")
    for m in ("astropy", "astropy.stats", "soundfile"):      # off-path imports of eval_utils.py:2-3
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["astropy.stats"].jackknife = None
    from src.caco_torch import create_caco_model
    from src.eval import eval_caco_torch as E
    return create_caco_model, E


def main():
    sys.path.insert(0, ROOT)
    import torch
    from oracle import weights as W
    torch.set_grad_enabled(False)
    create_caco_model, E = _import_reference()
    os.makedirs(GOLDEN, exist_ok=True)

    only = set(sys.argv[1:])         # e.g. `python oracle/make_golden.py model_s3_outlier`: only that fixture
    # ---- frontend -------------------------------------------------------------------------
    fe = {}
    for name, seed, kind, n, mp in FRONTEND_CASES:
        w = W.make_waveforms(seed, 1, n, kind)
        mel = E.compute_mel_spectrogram(torch.from_numpy(w))
        p = E.spectrogram_to_patches(mel, max_patches=mp)
        fe[name + "/mel_sub"] = mel[::3, ::5].copy()
        fe[name + "/mel_sum"] = np.asarray([mel.astype(np.float64).sum(), np.abs(mel).astype(np.float64).sum()])
        fe[name + "/patches_sub"] = p["audio_patches"][::3, ::7].copy()
        for k in ("audio_time_inds", "audio_freq_inds", "audio_mask"):
            fe[name + "/" + k] = p[k]
        if n <= 20000:
            fe[name + "/mel"] = mel
            fe[name + "/patches"] = p["audio_patches"]
    if not only or "frontend" in only:
        np.savez_compressed(os.path.join(GOLDEN, "frontend.npz"), **fe)
        print("frontend.npz", len(fe), "arrays")

    # ---- model ----------------------------------------------------------------------------
    for name, c in MODEL_CASES.items():
        if only and name not in only:
            continue
        sd = W.make_state_dict(c["seed"], c["sharp"], c.get("outlier", False), decoder_layers=c.get("decoder_layers", 0))
        ref = create_caco_model().eval()
        missing, unexpected = ref.load_state_dict(sd, strict=False)
        assert not unexpected and all(k.startswith("decoder_module") for k in missing)
        assert not (c.get("decoder_layers") and missing)
        waves, ids, mask = case_inputs(c)
        cfg = E.DatasetConfig(patches_seq_len=c["max_patches"])
        bs = [E.prepare_audio_batch(torch.from_numpy(w)[None], cfg, "cpu") for w in waves]
        ab = {k: torch.cat([b[k] for b in bs]) for k in bs[0]}
        ids_t, mask_t = torch.from_numpy(ids), torch.from_numpy(mask)
        a_raw, a_hid = ref.get_audio_embedding(**ab)
        a_n = ref.get_audio_embedding(**ab, return_hidden_state=False, normalize=True)
        t_raw, t_hid = ref.get_text_embedding(ids_t, mask_t)
        t_n = ref.get_text_embedding(ids_t, mask_t, return_hidden_state=False, normalize=True)
        out = dict(audio_emb_raw=a_raw, audio_emb=a_n, text_emb_raw=t_raw, text_emb=t_n,
                   audio_hidden_sub=a_hid[:, ::25, ::16], text_hidden_sub=t_hid[:, ::4, ::16],
                   audio_hidden_valid_absmean=(a_hid.abs() * ab["audio_mask"][..., None]).mean((1, 2)))
        if len(waves) == len(c["cap_lens"]):
            at, ta = ref(**ab, text_input_ids=ids_t, text_mask=mask_t)
            out["at_logits"], out["ta_logits"] = at, ta
        if c.get("decoder_layers"):
            dl = ref.get_decoder_logits(a_hid, ab["audio_mask"], ids_t, mask_t)          # [B, T, 50265]
            out["decoder_logits_sub"] = dl[:, :, ::97].contiguous()
            out["decoder_logits_last_valid"] = torch.stack([dl[b, n - 1] for b, n in enumerate(c["cap_lens"])])
            out["decoder_argmax"] = dl.argmax(-1)
        logits = torch.exp(ref.logit_scale) * a_n @ t_n.T                 # eval_caco_torch.py:330
        out["zs_logits"] = logits
        out["zs_top1"] = torch.argsort(-logits, dim=-1)[:, 0]
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"),
                            **{k: v.numpy() for k, v in out.items()})
        print(name, {k: tuple(v.shape) for k, v in out.items()})


if __name__ == "__main__":
    main()
