"""GPU: the audio-attention kernel in every exp2 split it ships (attn_poly 0..3: none, 1/4, 1/2, 3/8 of the scores through the
FMA-pipe polynomial instead of MUFU.EX2) against the same fp32 reference math, including ragged masks, holes in the mask,
one live key, and the sharp-softmax (lazy rescale) path."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from cacophony_b200 import _lib as L
from cacophony_b200 import ops
from tests.test_ops_gpu import _attn_ref

IMPLS = {"mufu_only": 0, "poly_1_of_4": 1, "poly_1_of_2": 2, "poly_3_of_8": 3}
DEFAULT_POLY = 0


@pytest.fixture(autouse=True)
def _restore_impl():
    yield
    assert L.load().caco_set_default_option(b"attn_poly", DEFAULT_POLY) == 0


@pytest.mark.parametrize("impl", list(IMPLS))
@pytest.mark.parametrize("S,valid,sharp", [(500, 496, 1.0), (500, 248, 1.0), (77, 32, 1.0), (129, 129, 1.0), (500, 496, 4.0),
                                           (1500, 1500, 1.0), (300, 1, 1.0)])
def test_attention_impl(impl, S, valid, sharp):
    assert L.load().caco_set_default_option(b"attn_poly", IMPLS[impl]) == 0
    g = torch.Generator().manual_seed(S + valid)
    B, H, dh = 3, 8, 96
    qkv = torch.randn(B, S, 3 * H * dh, generator=g) * 1.2
    qkv[..., : 2 * H * dh] *= sharp
    qkv = qkv.half()
    mask = torch.zeros(B, S)
    mask[0, :valid] = 1
    mask[1, : max(1, valid // 2)] = 1
    mask[2, :valid] = 1
    mask[2, 5:9] = 0            # holes in the mask, not just a prefix
    out = ops.attention_audio(qkv.cuda(), mask.cuda(), H).cpu().float()
    ref = _attn_ref(qkv, mask, H)
    assert torch.isfinite(out).all()
    assert float((out - ref).norm() / ref.norm()) < 1e-3
    np.testing.assert_allclose(out.numpy(), ref.numpy(), atol=6e-3, rtol=3e-3)
