"""GPU: every non-GEMM kernel through the C ABI against the CPU oracle (oracle/caco_oracle.py) and the golden
vectors the reference produced (tests/golden/frontend.npz)."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from cacophony_b200 import ops
from oracle import caco_oracle as O
from oracle import weights as W
from oracle.make_golden import FRONTEND_CASES
from tests.util import assert_logmel_close, to_linear

torch.set_grad_enabled(False)


@pytest.fixture(scope="module")
def fe_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "frontend.npz"))


@pytest.mark.parametrize("case", FRONTEND_CASES, ids=[c[0] for c in FRONTEND_CASES])
def test_frontend_matches_reference_golden(case, fe_golden):
    name, seed, kind, n, mp = case
    w = W.make_waveforms(seed, 1, n, kind)
    out = ops.frontend(torch.from_numpy(w).cuda(), mp, want_log_mel=True)
    mel = out["log_mel"][0].cpu().numpy()
    patches = out["audio_patches"][0].cpu().numpy()
    assert mel.shape == (O.num_frames(n), 128)
    lin = to_linear(mel)
    rowmax = lin.max(1, keepdims=True) if mel.shape[0] else lin
    assert_logmel_close(mel[::3, ::5], fe_golden[name + "/mel_sub"], rowmax[::3], name)
    for k in ("audio_time_inds", "audio_freq_inds", "audio_mask"):
        np.testing.assert_array_equal(out[k][0].cpu().numpy(), fe_golden[name + "/" + k])
    mask = fe_golden[name + "/audio_mask"].astype(bool)
    clipmax = lin.max() if lin.size else 1.0
    assert_logmel_close(patches[::3, ::7][mask[::3]], fe_golden[name + "/patches_sub"][mask[::3]], clipmax, name)
    assert not patches[~mask].any()
    if name + "/mel" in fe_golden:
        assert_logmel_close(mel, fe_golden[name + "/mel"], rowmax, name)
        assert_logmel_close(patches[mask], fe_golden[name + "/patches"][mask], clipmax, name)


def test_frontend_batched_equals_oracle_and_f16_copy():
    waves = W.make_waveforms(5, 3, 48000, "noise")
    waves[1, 30000:] = 0.0                     # silence tail -> log(1e-5) floor
    out = ops.frontend(torch.from_numpy(waves).cuda(), 500, want_f16=True)
    ref = O.prepare_audio_batch(list(waves), 500)
    for b in range(3):
        mel = O.log_mel(torch.from_numpy(waves[b])).numpy()
        clipmax = to_linear(mel).max()
        m = ref["audio_mask"][b].bool().numpy()
        assert_logmel_close(out["audio_patches"][b].cpu().numpy()[m], ref["audio_patches"][b].numpy()[m], clipmax, f"clip{b}")
    for k in ("audio_time_inds", "audio_freq_inds", "audio_mask"):
        assert torch.equal(out[k].cpu(), ref[k])
    assert torch.equal(out["audio_patches_f16"].cpu(), out["audio_patches"].cpu().half())


def test_layernorm_matches_oracle():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1003, 768, generator=g) * 3 + 0.5
    gamma, beta = torch.randn(768, generator=g), torch.randn(768, generator=g)
    o32, o16 = ops.layernorm(x.cuda(), gamma.cuda(), beta.cuda(), 1e-5, want_f32=True, want_f16=True)
    ref = O.layer_norm(x, gamma, beta)
    np.testing.assert_allclose(o32.cpu().numpy(), ref.numpy(), atol=2e-5, rtol=1e-5)
    assert torch.equal(o16.cpu(), o32.cpu().half())


def test_audio_add_pos_matches_oracle():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 500, 768, generator=g)
    t = torch.arange(500).float().div(8, rounding_mode="floor")[None].repeat(2, 1)
    f = (torch.arange(500) % 8).float()[None].repeat(2, 1)
    fe = torch.randn(8, 768, generator=g) * 0.02
    ref = x + O.sincos_time_embed(t, 768) + fe[f.long()]
    y = ops.audio_add_pos(x.clone().cuda(), t.cuda(), f.cuda(), fe.cuda())
    np.testing.assert_allclose(y.cpu().numpy(), ref.numpy(), atol=3e-5)      # sin/cos of angles up to 62 rad in fp32


def _attn_ref(qkv, mask, heads, causal=False):
    B, S, D3 = qkv.shape
    D = D3 // 3
    dh = D // heads
    q, k, v = [z.float().reshape(B, S, heads, dh).transpose(1, 2) for z in qkv.split(D, -1)]
    s = (q / math.sqrt(dh)) @ k.transpose(-1, -2)
    allow = mask.bool()[:, None, None, :]
    if causal:
        allow = allow & torch.tril(torch.ones(S, S, dtype=torch.bool))[None, None]
    s = s.masked_fill(~allow, float("-inf"))
    return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, S, D)


@pytest.mark.parametrize("S,valid", [(500, 496), (500, 248), (77, 32), (64, 64), (1500, 1500)])
def test_attention_audio_matches_reference_math(S, valid):
    g = torch.Generator().manual_seed(S)
    B, H, dh = 2, 8, 96
    qkv = (torch.randn(B, S, 3 * H * dh, generator=g) * 1.5).half()
    mask = torch.zeros(B, S)
    mask[0, :valid] = 1
    mask[1, : max(1, valid // 2)] = 1
    out = ops.attention_audio(qkv.cuda(), mask.cuda(), H).cpu().float()
    ref = _attn_ref(qkv, mask, H)
    assert torch.isfinite(out).all()
    # P is rounded to fp16 before P·V (2^-11 relative per weight), output rounded to fp16
    np.testing.assert_allclose(out.numpy(), ref.numpy(), atol=4e-3, rtol=2e-3)
    assert float((out - ref).norm() / ref.norm()) < 1e-3


def test_attention_audio_sharp_scores():
    g = torch.Generator().manual_seed(7)
    B, S, H, dh = 1, 500, 8, 96
    qkv = torch.randn(B, S, 3 * H * dh, generator=g)
    qkv[..., : 2 * H * dh] *= 4.0              # peaky softmax rows
    qkv = qkv.half()
    mask = torch.ones(B, S)
    mask[:, 496:] = 0
    out = ops.attention_audio(qkv.cuda(), mask.cuda(), H).cpu().float()
    ref = _attn_ref(qkv, mask, H)
    assert float((out - ref).norm() / ref.norm()) < 1e-3


@pytest.mark.parametrize("T,lens", [(32, [32, 20, 9, 4]), (100, [8, 10, 12, 100]), (1, [1]), (33, [33, 2]), (200, [200, 130])])
def test_attention_text_matches_reference_math(T, lens):
    g = torch.Generator().manual_seed(T)
    B, H = len(lens), 12
    qkv = (torch.randn(B, T, 3 * 768, generator=g) * 1.5).half()
    mask = torch.zeros(B, T)
    for b, n in enumerate(lens):
        mask[b, :n] = 1
    out = ops.attention_text(qkv.cuda(), mask.cuda(), H).cpu().float()
    ref = _attn_ref(qkv, mask, H, causal=True)
    np.testing.assert_allclose(out.numpy(), ref.numpy(), atol=3e-3, rtol=2e-3)


def test_text_embed_ln_matches_oracle():
    g = torch.Generator().manual_seed(2)
    V, P, D = 1000, 514, 768
    word, pos, ty = torch.randn(V, D, generator=g), torch.randn(P, D, generator=g), torch.randn(1, D, generator=g)
    gamma, beta = torch.randn(D, generator=g), torch.randn(D, generator=g)
    ids = torch.randint(0, V, (3, 32), generator=g)
    x = word[ids] + pos[torch.arange(32)][None] + ty[0]
    ref = O.layer_norm(x, gamma, beta)
    o32, o16 = ops.text_embed_ln(ids.cuda(), None, word.cuda(), pos.cuda(), ty.cuda(), gamma.cuda(), beta.cuda())
    np.testing.assert_allclose(o32.cpu().numpy(), ref.numpy(), atol=2e-5, rtol=1e-5)
    assert torch.equal(o16.cpu(), o32.cpu().half())
    pids = torch.randint(0, P, (3, 32), generator=g)
    ref2 = O.layer_norm(word[ids] + pos[pids] + ty[0], gamma, beta)
    o32b, _ = ops.text_embed_ln(ids.cuda(), pids.cuda(), word.cuda(), pos.cuda(), ty.cuda(), gamma.cuda(), beta.cuda())
    np.testing.assert_allclose(o32b.cpu().numpy(), ref2.numpy(), atol=2e-5, rtol=1e-5)


@pytest.mark.parametrize("with_ln", [True, False])
def test_attn_pool_matches_direct_formula(with_ln):
    g = torch.Generator().manual_seed(4)
    B, S, D, H = 3, 500, 768, 2
    hid = torch.randn(B, S, D, generator=g) * 2
    mask = torch.ones(B, S)
    mask[0, 496:] = 0
    mask[1, 248:] = 0
    mask[2, 1:] = 0
    u, c = torch.randn(H, D, generator=g) * 0.05, torch.randn(H, generator=g)
    gamma, beta = torch.randn(D, generator=g), torch.randn(D, generator=g)
    hn = O.layer_norm(hid, gamma, beta) if with_ln else hid
    s = torch.einsum("hd,bjd->bhj", u, hn) + c[None, :, None]
    s = s.masked_fill((mask == 0)[:, None, :], float("-inf"))
    ref = torch.einsum("bhj,bjd->bhd", torch.softmax(s, -1), hn)
    pooled, hid_out = ops.attn_pool(hid.cuda(), mask.cuda(), u.cuda(), c.cuda(), gamma.cuda() if with_ln else None,
                                    beta.cuda() if with_ln else None, 1e-5, want_hidden=True)
    np.testing.assert_allclose(pooled.cpu().numpy(), ref.numpy(), atol=3e-5, rtol=1e-4)
    if with_ln:
        np.testing.assert_allclose(hid_out.cpu().numpy(), hn.numpy(), atol=3e-5, rtol=1e-5)


@pytest.mark.parametrize("H", [1, 3, 5, 8])
def test_attn_pool_head_counts(H):
    """1..8 pooler heads (the JAX configuration uses 8, src/caco/load_model.py:47; more than four take a second pass)."""
    g = torch.Generator().manual_seed(40 + H)
    B, S, D = 2, 300, 768
    hid = torch.randn(B, S, D, generator=g)
    mask = torch.ones(B, S)
    mask[1, 200:] = 0
    u, c = torch.randn(H, D, generator=g) * 0.05, torch.randn(H, generator=g)
    gamma, beta = torch.randn(D, generator=g), torch.randn(D, generator=g)
    hn = O.layer_norm(hid, gamma, beta)
    s = torch.einsum("hd,bjd->bhj", u, hn) + c[None, :, None]
    s = s.masked_fill((mask == 0)[:, None, :], float("-inf"))
    ref = torch.einsum("bhj,bjd->bhd", torch.softmax(s, -1), hn)
    pooled, hid_out = ops.attn_pool(hid.cuda(), mask.cuda(), u.cuda(), c.cuda(), gamma.cuda(), beta.cuda(), 1e-5, want_hidden=True)
    assert pooled.shape == (B, H, D)
    np.testing.assert_allclose(pooled.cpu().numpy(), ref.numpy(), atol=3e-5, rtol=1e-4)
    np.testing.assert_allclose(hid_out.cpu().numpy(), hn.numpy(), atol=3e-5, rtol=1e-5)


def test_sgemm_l2norm_sim():
    g = torch.Generator().manual_seed(6)
    a, w, b = torch.randn(70, 768, generator=g), torch.randn(130, 768, generator=g), torch.randn(130, generator=g)
    y = ops.sgemm_nt(a.cuda(), w.cuda(), b.cuda()).cpu()
    np.testing.assert_allclose(y.numpy(), (a.double() @ w.double().t() + b).float().numpy(), atol=2e-4, rtol=1e-5)
    e = ops.l2norm(a.cuda()).cpu()
    np.testing.assert_allclose(e.numpy(), O.l2_normalize(a).numpy(), atol=1e-7, rtol=1e-6)
    t = O.l2_normalize(w)
    ls = torch.tensor([2.6592])
    at, ta = ops.sim_logits(e.cuda(), t.cuda(), ls.cuda())
    ref_at, ref_ta = O.contrastive_logits({"logit_scale": ls[0]}, e, t)
    np.testing.assert_allclose(at.cpu().numpy(), ref_at.numpy(), atol=2e-5)
    np.testing.assert_allclose(ta.cpu().numpy(), ref_ta.numpy(), atol=2e-5)
