"""GPU: every audio-attention implementation (persistent ping-pong tcgen05, persistent, one-tile tcgen05, mma.sync) against
the same fp32 reference math, including ragged masks and the sharp-softmax (rescale) path."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from cacophony_b200 import _lib as L
from cacophony_b200 import ops
from tests.test_ops_gpu import _attn_ref

IMPLS = {"mma_sync": 1, "tc_one_tile": 2, "tc_persistent": 3, "tc_pingpong": 4}


@pytest.fixture(autouse=True)
def _restore_impl():
    yield
    L.load().caco_set_attention_impl(0)


@pytest.mark.parametrize("impl", list(IMPLS))
@pytest.mark.parametrize("S,valid,sharp", [(500, 496, 1.0), (500, 248, 1.0), (77, 32, 1.0), (129, 129, 1.0), (500, 496, 4.0),
                                           (1500, 1500, 1.0), (300, 1, 1.0)])
def test_attention_impl(impl, S, valid, sharp):
    L.load().caco_set_attention_impl(IMPLS[impl])
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
