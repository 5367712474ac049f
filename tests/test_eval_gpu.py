"""GPU: rows (f-1), (f-2), (f-4) and BASELINE config #5 through the C ABI against the CPU oracle and the golden arrays
generated from the reference (tests/golden/eval_retrieval.npz)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import cacophony_b200 as cb
from cacophony_b200 import eval as ev
from cacophony_b200 import hear, loader, ops
from oracle import caco_oracle as O
from oracle import eval_oracle as E
from oracle import weights as W
from oracle.make_golden_eval import CASES
from tests.util import assert_logmel_close, rel_rows, to_linear

torch.set_grad_enabled(False)


class FakeTokenizer:
    """Deterministic stand-in with RobertaTokenizerFast's call interface: <s>=0, body = hash of each word, </s>=2, pad=1."""

    def __call__(self, texts, padding="max_length", truncation=True, max_length=100, return_tensors="pt"):
        ids = torch.full((len(texts), max_length), 1, dtype=torch.int64)
        mask = torch.zeros((len(texts), max_length), dtype=torch.int64)
        for i, t in enumerate(texts):
            body = [3 + (sum(ord(ch) * (j + 1) for j, ch in enumerate(w)) * 7919) % 50000 for w in t.split()][: max_length - 2]
            row = [0] + body + [2]
            ids[i, : len(row)] = torch.tensor(row)
            mask[i, : len(row)] = 1
        return {"input_ids": ids, "attention_mask": mask}


@pytest.fixture(scope="module")
def model(synthetic_state_dict):
    m = cb.create_caco_model()
    m.load_state_dict(synthetic_state_dict(2, 1.0))
    return m.to("cuda")


# ------------------------------------------------------------------------------------------------- top-k / metrics
@pytest.mark.parametrize("rows,cols,k", [(400, 50, 1), (37, 5225, 10), (1045, 209, 10), (5, 33, 32), (3, 10, 10), (1, 1, 1)])
def test_topk_rows_is_bit_exact(rows, cols, k):
    rng = np.random.default_rng(rows * 7 + cols)
    x = rng.standard_normal((rows, cols)).astype(np.float32)
    if cols > 20:
        x[:, 3] = x[:, 17]                  # ties -> lower column first
        x[0, :] = 0.25                      # a fully tied row
        x[-1, 5] = np.nan                   # NaN ranks last
        x[-1, 6] = -np.inf
        x[-1, 7] = np.inf
    idx, val = ops.topk_rows(torch.from_numpy(x).cuda(), k, want_values=True)
    ref = E.topk_indices(x, k)
    assert np.array_equal(idx.cpu().numpy(), ref)
    assert np.array_equal(val.cpu().numpy(), np.take_along_axis(x, ref, axis=1), equal_nan=True)


def test_topk_full_row_with_nans_is_a_permutation():
    x = np.array([[np.nan, 1.0, -np.inf, np.nan, 3.0, np.inf, -2.0, 1.0]], dtype=np.float32)
    idx = ops.topk_rows(torch.from_numpy(x).cuda(), 8).cpu().numpy()
    assert idx.tolist() == [[5, 4, 1, 7, 6, 2, 0, 3]]
    assert np.array_equal(idx, E.topk_indices(x, 8))


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_retrieval_metric_matches_reference_golden(case, golden_dir, capsys):
    name, seed, n_audio, caps, dup = case
    g = np.load(os.path.join(golden_dir, "eval_retrieval.npz"))
    names, all_text, gt_at, gt_ta, at_idx, ta_idx = E.make_retrieval_case(seed, n_audio, caps, dup)
    for kind, idx, qs, ks, gt in (("at", at_idx, names, all_text, gt_at), ("ta", ta_idx, all_text, names, gt_ta)):
        res = ev.compute_retrieval_metric(idx, qs, ks, gt, kind, "cuda")
        for metric in ("R1", "R5", "R10", "mAP10"):
            assert np.array_equal(res["per_query"][metric], g[f"{name}/{kind}/{metric}"]), (name, kind, metric)
        ref = E.compute_retrieval_metric(idx, qs, ks, gt, kind)
        for metric in ("R1", "R5", "R10", "mAP10"):
            np.testing.assert_allclose(res[metric], ref[metric], rtol=1e-10, atol=1e-12)
    out = capsys.readouterr().out.split("\n")
    assert out[0].startswith("R1 ") and out[3].startswith("mAP10 ") and "[" in out[0]       # the reference's printed lines


def test_avg_pool_tokens():
    rng = np.random.default_rng(1)
    h = rng.standard_normal((3, 500, 768)).astype(np.float32)
    got = ops.avg_pool_tokens(torch.from_numpy(h).cuda(), 8).cpu().numpy()
    ref = E.avg_pool_tokens(h, 8)
    assert got.shape == (3, 62, 768)
    np.testing.assert_allclose(got, ref, rtol=1e-6, atol=1e-6)


# ------------------------------------------------------------------------------------------------- ragged loader
def test_ragged_frontend_equals_per_clip_oracle():
    lens = [160000, 80000, 159999, 12345, 100, 192000, 2560, 2559]
    waves = [W.make_waveforms(40 + i, 1, n, "noise" if i % 2 == 0 else "chirp")[0] for i, n in enumerate(lens)]
    out = loader.prepare_audio_batch_ragged(waves, cb.DatasetConfig(patches_seq_len=500), "cuda")
    assert out["audio_patches"].shape == (len(lens), 500, 256)
    valid = loader.valid_patch_counts(lens, 500)
    for i, w in enumerate(waves):
        ref = O.prepare_audio_batch([w], 500)
        for k in ("audio_time_inds", "audio_freq_inds", "audio_mask"):
            assert np.array_equal(out[k][i].cpu().numpy(), ref[k][0].numpy()), (i, k)          # index work: bit-exact
        assert int(out["audio_mask"][i].sum()) == valid[i]
        y, r = out["audio_patches"][i].cpu().numpy(), ref["audio_patches"][0].numpy()
        v = int(valid[i])
        assert not y[v:].any()                                                                 # zero padding rows
        if v:
            assert_logmel_close(y[:v], r[:v], to_linear(r[:v]).max(), f"clip {i}")
        # and identical to the uniform-batch kernel path on that clip alone
        one = cb.prepare_audio_batch(torch.from_numpy(w)[None], cb.DatasetConfig(patches_seq_len=500), "cuda")
        assert torch.equal(one["audio_patches"][0], out["audio_patches"][i])


def test_resample_matches_scipy():
    rng = np.random.default_rng(9)
    for sr, n in ((44100, 44100), (48000, 24001), (8000, 8000), (22050, 11025), (32000, 6400)):
        x = (0.1 * rng.standard_normal(n)).astype(np.float32)
        ref = E.resample(x.astype(np.float64), sr)
        got = loader.resample_to_16k(x, sr, "cuda").cpu().numpy()
        assert got.shape == ref.shape == (round(n * 16000.0 / sr),)
        assert np.abs(got - ref).max() < 2e-6 * max(1.0, np.abs(ref).max() / 0.1)             # fp32 FFT round trip
    assert torch.equal(loader.resample_to_16k(x, 16000, "cuda").cpu(), torch.from_numpy(x))


# ------------------------------------------------------------------------------------------------- model-level rows
def test_encode_audio_ragged_and_trimmed(model):
    lens = [80000, 160000, 51234]
    waves = [W.make_waveforms(60 + i, 1, n, "noise")[0] for i, n in enumerate(lens)]
    buf, ln = loader.pad_ragged(waves)
    e_ragged = model.encode_audio(buf.cuda(), max_patches=500, lengths=ln)
    e_trim, hid, mk = model.encode_audio(buf.cuda(), max_patches=500, lengths=ln, trim_padding=True, return_hidden_state=True)
    assert hid.shape == (3, 496, 768) and mk.shape == (3, 496) and mk.sum(1).tolist() == [248.0, 496.0, 160.0]
    for i, w in enumerate(waves):
        ab = cb.prepare_audio_batch(torch.from_numpy(w)[None], cb.DatasetConfig(patches_seq_len=500), "cuda")
        e_ref = model.get_audio_embedding(**ab, return_hidden_state=False, normalize=True)
        assert rel_rows(e_ragged[i:i + 1], e_ref) < 2e-6       # same kernels, same operands: batch position only
        assert rel_rows(e_trim[i:i + 1], e_ref) < 2e-5         # masked keys contribute exactly 0; tile order may differ
    # against the CPU oracle (reference arithmetic, fp32): the north-star 1e-3 bar
    sd = {k: v.cpu() for k, v in model.state_dict().items()}
    ab = O.prepare_audio_batch([waves[0]], 500)
    a_ref, _ = O.get_audio_embedding(sd, ab["audio_patches"], ab["audio_time_inds"], ab["audio_freq_inds"], ab["audio_mask"],
                                     normalize=True)
    assert rel_rows(e_trim[0:1], a_ref) < 1e-3 and rel_rows(e_ragged[0:1], a_ref) < 1e-3


def test_zero_shot_config5_shape_matches_oracle(model):
    """BASELINE config #5 at a size the CPU oracle finishes in seconds: 5 s clips (248 of 500 tokens valid), prompts
    padded to T = 100 with 8-12 valid tokens; top-1 must agree with the oracle wherever the margin is clear."""
    n_clips, n_cls = 10, 7
    waves = [W.make_waveforms(80 + i, 1, 80000, "noise" if i % 3 else "chirp")[0] for i in range(n_clips)]
    ids, mask = W.make_captions(11, n_cls, 100, lens=[8, 9, 10, 11, 12, 8, 12])
    t = ev.embed_text_ids(model, torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda())
    a = ev.embed_waveforms(model, waves, cb.DatasetConfig(patches_seq_len=500), batch_size=4)
    logits = ev.zero_shot_logits(model, a, t).cpu().numpy()
    top1 = ev.zero_shot_topk(model, a, t, 1).cpu().numpy()[:, 0]
    sd = {k: v.cpu() for k, v in model.state_dict().items()}
    ab = O.prepare_audio_batch(waves, 500)
    a_ref, _ = O.get_audio_embedding(sd, ab["audio_patches"], ab["audio_time_inds"], ab["audio_freq_inds"], ab["audio_mask"],
                                     normalize=True)
    t_ref, _ = O.get_text_embedding(sd, torch.from_numpy(ids), torch.from_numpy(mask), normalize=True)
    ref_logits = (torch.exp(sd["logit_scale"]) * a_ref @ t_ref.T).numpy()
    scale = float(np.exp(W.LOGIT_SCALE_INIT))
    assert logits.shape == (n_clips, n_cls) and np.abs(logits - ref_logits).max() < 1e-3 * scale
    srt = np.sort(ref_logits, -1)
    clear = (srt[:, -1] - srt[:, -2]) > 2e-3 * scale
    assert clear.sum() >= n_clips // 2
    assert np.array_equal(top1[clear], ref_logits.argmax(-1)[clear])
    assert np.array_equal(top1, logits.argmax(-1))                     # device top-1 == argmax of the device logits, always
    acc = ev.zs_classification_arrays(model, t, waves, list(logits.argmax(-1)), cb.DatasetConfig(patches_seq_len=500))
    assert acc == {"1": 1.0}


def test_audio_retrieval_driver_end_to_end(model, capsys):
    """Driver output == reference metric code (oracle restatement, pinned by the golden test above) applied to a full
    argsort of the driver's own similarity matrix."""
    n = 12
    waves = [W.make_waveforms(120 + i, 1, 40000 + 8000 * (i % 4), "noise")[0] for i in range(n)]
    names = [f"a{i}" for i in range(n)]
    caps = [[f"sound number {i} take {c}" for c in range(2)] for i in range(n)]
    caps[3][1] = caps[7][0]                                              # a caption string shared by two clips
    tok = FakeTokenizer()
    cfg = cb.DatasetConfig(patches_seq_len=500, max_text_len=32)
    res = ev.audio_retrieval_arrays(model, waves, names, caps, tok, cfg)
    printed = capsys.readouterr().out
    assert "audio to text retrieval:" in printed and "text to audio retrieval:" in printed and printed.count("mAP10") == 2
    all_text = [c for cs in caps for c in cs]
    tb = ev.prepare_text_batch(all_text, tok, 32, "cuda")
    t = ev.embed_text_ids(model, tb["text_input_ids"], tb["text_mask"]).cpu().numpy()
    a = ev.embed_waveforms(model, waves, cfg).cpu().numpy()
    logits_ar = t @ a.T
    gt_at = {nm: list(cs) for nm, cs in zip(names, caps)}
    gt_ta = {c: nm for nm, cs in zip(names, caps) for c in cs}
    ref_at = E.compute_retrieval_metric(E.topk_indices(logits_ar.T, 10), names, all_text, gt_at, "at")
    ref_ta = E.compute_retrieval_metric(E.topk_indices(logits_ar, 10), all_text, names, gt_ta, "ta")
    for m in ("R1", "R5", "R10", "mAP10"):
        np.testing.assert_allclose(res["at"][m], ref_at[m], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(res["ta"][m], ref_ta[m], rtol=1e-9, atol=1e-12)


def test_hear_embeddings(model):
    emb = hear.Embedding(model, audio_max_len=10)
    assert emb.max_patches == 496                                         # caco_embeddings.py:72-73
    w = W.make_waveforms(150, 1, 160000, "noise")[0]
    scene = emb.get_embedding_as_numpy(w)
    ev_emb, ts = emb.get_embedding_as_numpy(w, "event")
    assert scene.shape == (768,) and abs(np.linalg.norm(scene) - 1.0) < 1e-5
    assert ev_emb.shape == (62, 768) and ts[0].shape == (62,) and ts[0][0] == 0 and ts[0][-1] == 10000
    ab = cb.prepare_audio_batch(torch.from_numpy(w)[None], cb.DatasetConfig(patches_seq_len=496), "cuda")
    e_ref, hid = model.get_audio_embedding(**ab, normalize=True)
    assert rel_rows(scene[None], e_ref) < 2e-6
    np.testing.assert_allclose(ev_emb, E.avg_pool_tokens(hid.cpu().numpy(), 8)[0], rtol=1e-5, atol=1e-5)


def test_text_padding_trim_is_exact(model):
    """Prompts padded to T = 100 (config 5): running the tower on the first 16 columns gives the same embeddings."""
    ids, mask = W.make_captions(21, 9, 100, lens=[8, 9, 10, 11, 12, 8, 12, 3, 2])
    ids, mask = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
    full = ev.embed_text_ids(model, ids, mask, trim_padding=False)
    trim = ev.embed_text_ids(model, ids, mask, trim_padding=True)
    assert rel_rows(trim, full) < 2e-6
    holes = mask.clone()
    holes[0, 3] = 0                                       # a hole in a mask must not move the trim point
    assert rel_rows(ev.embed_text_ids(model, ids, holes), ev.embed_text_ids(model, ids, holes, trim_padding=False)) < 2e-6


def test_load_audio_wav_file_matches_reference_pipeline(tmp_path):
    """eval_utils.py:6-16 end to end for a 44.1 kHz stereo int16 WAV: read, channel mean, Fourier resample to 16 kHz."""
    from scipy.io import wavfile
    rng = np.random.default_rng(4)
    pcm = (rng.standard_normal((22050, 2)) * 3000).astype(np.int16)
    path = str(tmp_path / "clip.wav")
    wavfile.write(path, 44100, pcm)
    got = loader.load_audio(path, 44100, "cuda").cpu().numpy()
    ref = (pcm.astype(np.float32) / 32768.0).mean(axis=-1)              # what soundfile.read returns for int16 PCM, then :9-10
    ref = E.resample(ref.astype(np.float64), 44100)
    assert got.shape == ref.shape == (8000,)
    assert np.abs(got - ref).max() < 5e-6


def test_device_resident_clips_skip_the_host(tmp_path, synthetic_state_dict):
    """Row f-1: clips that are already on the GPU (what load_audio returns) are packed with device-to-device copies and give
    bit-identical embeddings to the pinned-host path; embed_files (double-buffered pinned staging -> GPU resample -> device
    packing) equals loading every file by hand."""
    from scipy.io import wavfile
    from cacophony_b200 import eval as ev
    from cacophony_b200 import loader
    m = cb.create_caco_model()
    m.load_state_dict(synthetic_state_dict(0, 1.0))
    m = m.to("cuda")
    rng = np.random.default_rng(3)
    clips = [(0.1 * rng.standard_normal(n)).astype(np.float32) for n in (80000, 52345, 160000, 16000, 99999)]
    e_host = ev.embed_waveforms(m, clips)
    e_dev = ev.embed_waveforms(m, [torch.from_numpy(c).cuda() for c in clips])
    assert torch.equal(e_host, e_dev)
    buf, lens = loader.pad_ragged_device([torch.from_numpy(c).cuda() for c in clips])
    hbuf, hlens = loader.pad_ragged(clips)
    assert torch.equal(buf.cpu(), hbuf) and torch.equal(lens.cpu(), hlens)
    paths = []
    for i, sr in enumerate((44100, 16000, 22050)):
        x = (0.2 * rng.standard_normal(sr * 2)).astype(np.float32)
        p = str(tmp_path / f"clip{i}.wav")
        wavfile.write(p, sr, x)
        paths.append((p, sr))
    for sr in (44100, 16000, 22050):
        group = [p for p, s in paths if s == sr]
        e_files = ev.embed_files(m, group, sr, batch_size=2)
        by_hand = ev.embed_waveforms(m, [loader.load_audio(p, sr, "cuda") for p in group])
        assert torch.equal(e_files, by_hand) and torch.isfinite(e_files).all()
