"""CPU: the evaluation-side oracle (oracle/eval_oracle.py) against the golden arrays produced by the reference's own
compute_retrieval_metric (tests/golden/eval_retrieval.npz, generator oracle/make_golden_eval.py), and the host-side logic
of the batched drivers (packing, metric arithmetic, jackknife closed form)."""
import os

import numpy as np
import pytest
import torch

from cacophony_b200 import eval as ev
from cacophony_b200 import loader
from oracle import eval_oracle as E
from oracle.make_golden_eval import CASES


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_retrieval_matches_reference_golden(case, golden_dir):
    name, seed, n_audio, caps, dup = case
    g = np.load(os.path.join(golden_dir, "eval_retrieval.npz"))
    names, all_text, gt_at, gt_ta, at_idx, ta_idx = E.make_retrieval_case(seed, n_audio, caps, dup)
    for kind, idx, qs, ks, gt in (("at", at_idx, names, all_text, gt_at), ("ta", ta_idx, all_text, names, gt_ta)):
        per_query = E.retrieval_per_query(E.retrieval_preds(idx, qs, ks, gt, kind))
        for metric in ("R1", "R5", "R10", "mAP10"):
            ref = g[f"{name}/{kind}/{metric}"]
            assert np.array_equal(per_query[metric], ref), (name, kind, metric)      # float64, bit-exact
        assert per_query["R10"].sum() > 0                                             # the cases do contain hits


def test_metrics_from_hit_bits_equals_oracle():
    rng = np.random.default_rng(3)
    preds = rng.random((500, 10)) < 0.2
    preds[0] = False
    preds[1] = True
    bits = (preds.astype(np.int64) << np.arange(10)[None, :]).sum(axis=1)
    got = ev.metrics_from_hit_bits(bits)
    ref = E.retrieval_per_query(preds)
    for k in ref:
        assert np.array_equal(got[k], ref[k]), k


def test_jackknife_closed_form_equals_leave_one_out():
    rng = np.random.default_rng(5)
    for data in (rng.random(57), (rng.random(200) < 0.3).astype(float), np.array([0.0, 1.0])):
        est, bias, se, ci = ev.jackknife_stats_mean(data, 0.95)
        r_est, r_bias, r_se, r_ci = E.jackknife_stats(data, np.mean, 0.95)
        np.testing.assert_allclose([est, se, ci[0], ci[1]], [r_est, r_se, r_ci[0], r_ci[1]], rtol=1e-10, atol=1e-12)
        assert abs(bias - r_bias) < 1e-12
    # known answer: for the mean, std_err is the usual standard error of the mean and z(0.95) = 1.959964
    x = np.arange(10, dtype=float)
    est, _, se, ci = ev.jackknife_stats_mean(x)
    assert abs(est - 4.5) < 1e-12 and abs(se - x.std(ddof=1) / np.sqrt(10)) < 1e-12
    assert abs((ci[1] - est) / se - 1.959963984540054) < 1e-9


def test_oracle_topk_is_argsort_of_negated_scores():
    rng = np.random.default_rng(7)
    x = rng.standard_normal((40, 300)).astype(np.float32)
    ref = torch.argsort(-torch.from_numpy(x), dim=-1, stable=True)[:, :10].numpy()
    assert np.array_equal(E.topk_indices(x, 10), ref)
    x[:, 5] = x[:, 200]                                  # ties resolve to the lower column
    top = E.topk_indices(x, 300)
    pos5 = np.argmax(top == 5, axis=1)
    pos200 = np.argmax(top == 200, axis=1)
    assert (pos5 + 1 == pos200).all()
    x[3, 7] = np.nan
    assert E.topk_indices(x, 300)[3, -1] == 7            # NaN ranks last


def test_pad_ragged_and_valid_counts():
    waves = [np.arange(5, dtype=np.float32), torch.ones(330), np.ones((200, 2), dtype=np.float64) * [1.0, 3.0]]
    buf, lens = loader.pad_ragged(waves, pin=False)
    assert buf.shape == (3, 480) and buf.dtype == torch.float32 and lens.tolist() == [5, 330, 200]
    assert buf[0, :5].tolist() == [0, 1, 2, 3, 4] and float(buf[0, 5:].abs().sum()) == 0
    assert float(buf[2, :200].mean()) == 2.0 and float(buf[2, 200:].abs().sum()) == 0      # channel mean
    buf2, lens2 = loader.pad_ragged(waves, stride=100, pin=False)
    assert buf2.shape == (3, 100) and lens2.tolist() == [5, 100, 100]
    # eval_caco_torch.py:67,116-117,132-138
    assert loader.valid_patch_counts([160000, 80000, 159999, 12345, 100, 192000], 500).tolist() == [496, 248, 496, 32, 0, 500]
    with pytest.raises(ValueError):
        loader.pad_ragged([])


def test_drivers_refuse_cpu_and_missing_tokenizer():
    with pytest.raises(RuntimeError, match="CUDA"):
        loader.prepare_audio_batch_ragged([np.zeros(16000, np.float32)], device="cpu")
    with pytest.raises(RuntimeError, match="CUDA"):
        loader.resample_to_16k(np.zeros(100, np.float32), 44100, device="cpu")
    with pytest.raises(ValueError, match="tokenizer"):
        ev.prepare_text_batch("a dog barks", None, 100, "cpu")


def test_load_caco_torch_accepts_the_reference_checkpoint_layouts():
    from oracle import weights as W
    sd = {n: torch.zeros(s) for n, s, _ in W.param_spec(decoder_layers=4)}      # a full checkpoint incl. the captioning head
    out = ev.load_caco_torch(None, "cpu", tokenizer="tok",
                             state_dict={k: v for k, v in sd.items() if not k.startswith("decoder_module.")})    # encoder-only
    assert out["model"].decoder_module is not None
    for wrap in (lambda d: d, lambda d: {"state_dict": d}, lambda d: {"model_state_dict": d, "epoch": 3}):
        out = ev.load_caco_torch(None, "cpu", tokenizer="tok", state_dict=wrap(sd))
        assert set(out) == {"model", "tokenizer", "device"} and out["tokenizer"] == "tok"
        assert not out["model"].training


class _ScriptedCaptioner:
    """Host-logic stand-in for CACO in decode_caption_ids' cached loop: next-token logits are scripted per (sequence, position),
    everything stays on the CPU (the arithmetic is not under test here, the loop of eval_caco_torch.py:411-472 is)."""

    def __init__(self, script, vocab=11, max_pos=6):
        self.script, self.vocab = script, vocab                       # script[b][t] = token emitted after position t
        self.text_config = type("C", (), {"max_position_embeddings": max_pos})()
        self.steps, self.begun = [], None

    def get_audio_embedding(self, **kw):
        B = kw["audio_patches"].shape[0]
        return torch.zeros(B, 4), torch.zeros(B, 3, 4)

    def decode_begin(self, audio_hidden, audio_mask, capacity, cache=None):
        self.begun = (tuple(audio_hidden.shape), capacity)
        return "cache"

    def decode_step(self, cache, ids, pos, want_logits=True, want_next=False, **kw):
        assert cache == "cache" and ids.dtype == torch.long and pos.dtype == torch.long
        self.steps.append((ids.tolist(), pos.tolist()))
        nxt = torch.tensor([self.script[b][int(p)] for b, p in enumerate(pos)], dtype=torch.int32)
        if want_next and not want_logits:
            return nxt
        logits = torch.full((len(nxt), self.vocab), -30.0)
        logits[torch.arange(len(nxt)), nxt.long()] = 30.0
        return logits


def test_cached_decode_loop_host_logic():
    """BOS first, one token per step fed back with its position, a finished sequence keeps emitting EOS, the loop stops when every
    sequence has finished, and the position limit raises like the full-prefix call does."""
    ab = {"audio_patches": torch.zeros(2, 3, 256), "audio_time_inds": torch.zeros(2, 3), "audio_freq_inds": torch.zeros(2, 3),
          "audio_mask": torch.ones(2, 3)}
    m = _ScriptedCaptioner([[5, 2, 9, 9, 9, 9], [7, 8, 2, 9, 9, 9]])
    out = ev.decode_caption_ids(m, ab, bos_id=0, eos_id=2, max_decode_length=5, temperature=0.0, use_cache=True)
    assert out.tolist() == [[0, 5, 2, 2], [0, 7, 8, 2]]              # stopped after step 3: both have produced EOS
    assert m.begun == ((2, 3, 4), 5)
    assert m.steps == [([0, 0], [0, 0]), ([5, 7], [1, 1]), ([2, 8], [2, 2])]
    # temperature sampling takes the logits path: with one-hot logits the sample is the scripted token
    m = _ScriptedCaptioner([[5, 2, 9, 9, 9, 9], [7, 8, 2, 9, 9, 9]])
    out = ev.decode_caption_ids(m, ab, bos_id=0, eos_id=2, max_decode_length=5, temperature=0.5, use_cache=True,
                                generator=torch.Generator().manual_seed(0))
    assert out.tolist() == [[0, 5, 2, 2], [0, 7, 8, 2]]
    # fixed length when EOS never comes; capacity = min(max_decode_length, max positions)
    m = _ScriptedCaptioner([[4, 4, 4, 4, 4, 4], [3, 3, 3, 3, 3, 3]])
    out = ev.decode_caption_ids(m, ab, eos_id=2, max_decode_length=4, use_cache=True)
    assert out.shape == (2, 5) and m.begun[1] == 4
    with pytest.raises(ValueError, match="sequence length"):
        ev.decode_caption_ids(m, ab, eos_id=2, max_decode_length=9, use_cache=True)      # > max_position_embeddings (6)
