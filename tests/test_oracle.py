"""CPU: the oracle restatement against the golden vectors produced by the reference itself
(oracle/make_golden.py).  Tolerances: fp32 re-association noise only."""
import os

import numpy as np
import pytest
import torch

from oracle import caco_oracle as O
from oracle import weights as W
from oracle.make_golden import FRONTEND_CASES, MODEL_CASES, case_inputs
from tests.util import assert_logmel_close, to_linear

torch.set_grad_enabled(False)


@pytest.fixture(scope="module")
def fe_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "frontend.npz"))


@pytest.mark.parametrize("case", FRONTEND_CASES, ids=[c[0] for c in FRONTEND_CASES])
def test_frontend_matches_reference(case, fe_golden):
    name, seed, kind, n, mp = case
    w = W.make_waveforms(seed, 1, n, kind)
    mel = O.log_mel(torch.from_numpy(w)).numpy()
    assert mel.shape == (O.num_frames(n), 128)
    p = O.patchify(mel, mp)
    # log-mel: torch.stft vs rfft-of-frames differ by fp32 rounding, amplified by log near the 1e-5 floor
    lin = to_linear(mel)
    rowmax = lin.max(1, keepdims=True) if mel.shape[0] else lin
    assert_logmel_close(mel[::3, ::5], fe_golden[name + "/mel_sub"], rowmax[::3], name)
    clipmax = lin.max() if lin.size else 1.0
    valid = p["audio_mask"][::3].astype(bool)
    assert_logmel_close(p["audio_patches"][::3, ::7][valid], fe_golden[name + "/patches_sub"][valid], clipmax, name)
    assert not p["audio_patches"][~p["audio_mask"].astype(bool)].any()
    for k in ("audio_time_inds", "audio_freq_inds", "audio_mask"):
        np.testing.assert_array_equal(p[k], fe_golden[name + "/" + k])
    assert p["audio_patches"].shape == (mp, 256) and p["audio_patches"].dtype == np.float32
    if name + "/mel" in fe_golden:
        assert_logmel_close(mel, fe_golden[name + "/mel"], rowmax, name)
        v = p["audio_mask"].astype(bool)
        assert_logmel_close(p["audio_patches"][v], fe_golden[name + "/patches"][v], clipmax, name)


def test_mel_filterbank_properties():
    fb = O.mel_filterbank()
    assert fb.shape == (257, 128)
    assert int((fb != 0).sum()) == 505
    assert float(fb[:, 0].abs().sum()) == 0.0   # mel bin 0 is empty -> constant log(1e-5)*0.2+0.9
    try:
        import torchaudio
    except Exception:
        pytest.skip("torchaudio not importable")
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = torchaudio.functional.melscale_fbanks(257, 0.0, 8000.0, 128, 16000, norm=None)
    assert torch.equal(ref, fb)


def _rel_rows(x, y):
    x, y = torch.as_tensor(x), torch.as_tensor(y)
    return float(((x - y).norm(dim=-1) / y.norm(dim=-1)).max())


@pytest.mark.parametrize("name", list(MODEL_CASES))
def test_model_matches_reference(name, golden_dir, synthetic_state_dict):
    c = MODEL_CASES[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    sd = synthetic_state_dict(c["seed"], c["sharp"], c.get("outlier", False), c.get("decoder_layers", 0))
    waves, ids, mask = case_inputs(c)
    ab = O.prepare_audio_batch(waves, c["max_patches"])
    ids, mask = torch.from_numpy(ids), torch.from_numpy(mask)
    a_raw, a_hid = O.get_audio_embedding(sd, ab["audio_patches"], ab["audio_time_inds"],
                                         ab["audio_freq_inds"], ab["audio_mask"])
    t_raw, t_hid = O.get_text_embedding(sd, ids, mask)
    a_n, t_n = O.l2_normalize(a_raw), O.l2_normalize(t_raw)
    tol = 2e-5          # fp32 summation-order noise through 12 layers (observed ~1e-6)
    assert _rel_rows(a_raw, g["audio_emb_raw"]) < tol
    assert _rel_rows(a_n, g["audio_emb"]) < tol
    assert _rel_rows(t_raw, g["text_emb_raw"]) < tol
    assert _rel_rows(t_n, g["text_emb"]) < tol
    valid = ab["audio_mask"][:, ::25].bool()
    np.testing.assert_allclose(a_hid[:, ::25, ::16][valid].numpy(), g["audio_hidden_sub"][valid.numpy()],
                               atol=2e-4, rtol=1e-4)
    tv = mask[:, ::4].bool()
    np.testing.assert_allclose(t_hid[:, ::4, ::16][tv].numpy(), g["text_hidden_sub"][tv.numpy()],
                               atol=2e-4, rtol=1e-4)
    at, ta = O.contrastive_logits(sd, a_n, t_n)
    np.testing.assert_allclose(at.numpy(), g["at_logits"], atol=2e-5 * float(torch.exp(sd["logit_scale"])))
    np.testing.assert_allclose(ta.numpy(), g["ta_logits"], atol=2e-5 * float(torch.exp(sd["logit_scale"])))
    assert np.array_equal(O.zero_shot_top1(sd, a_n, t_n).numpy(), g["zs_top1"])


def test_rounding_emulation_is_within_design_budget(golden_dir, synthetic_state_dict):
    """The device path rounds GEMM operands to fp16 (fp32 accumulate).  Emulated on the CPU this must
    stay inside the 1e-3 embedding budget on the default synthetic model — if this fails the
    precision scheme itself (not a kernel) is wrong."""
    c = MODEL_CASES["model_s0"]
    g = np.load(os.path.join(golden_dir, "model_s0.npz"))
    sd = synthetic_state_dict(c["seed"], c["sharp"], c.get("outlier", False), c.get("decoder_layers", 0))
    waves, ids, mask = case_inputs(c)
    ab = O.prepare_audio_batch(waves[:2], c["max_patches"])
    r = O.Rounding.fp16()
    a, _ = O.get_audio_embedding(sd, ab["audio_patches"], ab["audio_time_inds"], ab["audio_freq_inds"],
                                 ab["audio_mask"], normalize=True, r=r)
    t, _ = O.get_text_embedding(sd, torch.from_numpy(ids[:2]), torch.from_numpy(mask[:2]), normalize=True, r=r)
    assert _rel_rows(a, g["audio_emb"][:2]) < 1e-3
    assert _rel_rows(t, g["text_emb"][:2]) < 1e-3


def test_decoder_logits_match_reference(golden_dir, synthetic_state_dict):
    """Row f-4: the oracle's restatement of CACO.get_decoder_logits (caco.py:214-240, roberta.py:329-373) against the golden
    logits the reference produced on the same synthetic captioning head."""
    c = MODEL_CASES["model_s4_decoder"]
    g = np.load(os.path.join(golden_dir, "model_s4_decoder.npz"))
    sd = synthetic_state_dict(c["seed"], c["sharp"], False, c["decoder_layers"])
    waves, ids, mask = case_inputs(c)
    ab = O.prepare_audio_batch(waves, c["max_patches"])
    _, hid = O.get_audio_embedding(sd, ab["audio_patches"], ab["audio_time_inds"], ab["audio_freq_inds"], ab["audio_mask"])
    dl = O.get_decoder_logits(sd, hid, ab["audio_mask"], torch.from_numpy(ids), torch.from_numpy(mask))
    assert dl.shape == (2, 24, 50265)
    valid = mask.astype(bool)
    sub = dl[:, :, ::97].numpy()
    assert np.abs(sub[valid] - g["decoder_logits_sub"][valid]).max() < 2e-4
    assert np.array_equal(dl.argmax(-1).numpy()[valid], g["decoder_argmax"][valid])


def test_incremental_decoder_equals_full_prefix(golden_dir, synthetic_state_dict):
    """Row f-4, "KV-cached sampling": the incremental restatement (one token per step against cached keys / values: the
    algorithm of caco_model_decode_begin / caco_model_decode_step) gives, at every position, the logits of the full-prefix call
    the reference's loop repeats (eval_caco_torch.py:411-472) — and therefore, on the valid positions, the golden logits the
    reference itself produced.  fp32 re-association only: 1e-5 of the row's spread."""
    c = MODEL_CASES["model_s4_decoder"]
    g = np.load(os.path.join(golden_dir, "model_s4_decoder.npz"))
    sd = synthetic_state_dict(c["seed"], c["sharp"], False, c["decoder_layers"])
    waves, ids, mask = case_inputs(c)
    ab = O.prepare_audio_batch(waves, c["max_patches"])
    _, hid = O.get_audio_embedding(sd, ab["audio_patches"], ab["audio_time_inds"], ab["audio_freq_inds"], ab["audio_mask"])
    T = 10
    ids_t = torch.from_numpy(ids)[:, :T]
    full = O.get_decoder_logits(sd, hid, ab["audio_mask"], ids_t, torch.ones(ids_t.shape))
    dec = O.IncrementalDecoder(sd, hid, ab["audio_mask"])
    for t in range(T):
        lg = dec.step(ids_t[:, t])
        ref = full[:, t]
        err = (lg - ref).norm(dim=-1) / (ref - ref.mean(-1, keepdim=True)).norm(dim=-1)
        assert float(err.max()) < 1e-5, (t, float(err.max()))
        # where the golden caption is still valid (no padding token pushed yet) this is also the reference's own logit row
        for b in range(2):
            if mask[b, : t + 1].all():
                assert np.abs(lg[b, ::97].numpy() - g["decoder_logits_sub"][b, t]).max() < 2e-4
