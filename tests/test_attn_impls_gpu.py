"""GPU: the audio-attention kernel (persistent ping-pong tcgen05, csrc/attention_pp.cu) and the warp-level fallback it shares
the entry point with, against the same fp32 reference math: ragged masks, holes in the mask, one live key, the sharp-softmax
(lazy rescale) path, 1500 tokens, and a 4200-token sequence that takes the fallback."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from cacophony_b200 import ops
from tests.test_ops_gpu import _attn_ref


@pytest.mark.parametrize("S,valid,sharp", [(500, 496, 1.0), (500, 248, 1.0), (77, 32, 1.0), (129, 129, 1.0), (500, 496, 4.0),
                                           (1500, 1500, 1.0), (300, 1, 1.0), (256, 256, 1.0), (4200, 4100, 1.0)])
def test_attention_audio_head_dim_96(S, valid, sharp):
    g = torch.Generator().manual_seed(S + valid)
    B, H, dh = (3, 8, 96) if S <= 1500 else (2, 2, 96)
    qkv = torch.randn(B, S, 3 * H * dh, generator=g) * 1.2
    qkv[..., : 2 * H * dh] *= sharp
    qkv = qkv.half()
    mask = torch.zeros(B, S)
    mask[0, :valid] = 1
    mask[1, : max(1, valid // 2)] = 1
    if B > 2:
        mask[2, :valid] = 1
        mask[2, 5:9] = 0            # holes in the mask, not just a prefix
    out = ops.attention_audio(qkv.cuda(), mask.cuda(), H).cpu().float()
    ref = _attn_ref(qkv, mask, H)
    assert torch.isfinite(out).all()
    assert float((out - ref).norm() / ref.norm()) < 1e-3
    np.testing.assert_allclose(out.numpy(), ref.numpy(), atol=6e-3, rtol=3e-3)


def test_attention_audio_is_deterministic_and_batch_invariant():
    g = torch.Generator().manual_seed(5)
    B, S, H, dh = 37, 500, 8, 96                    # 37 * 8 * 2 = 592 work items over 148 persistent CTAs
    qkv = (torch.randn(B, S, 3 * H * dh, generator=g)).half().cuda()
    mask = torch.ones(B, S).cuda()
    mask[:, 496:] = 0
    o1 = ops.attention_audio(qkv, mask, H)
    o2 = ops.attention_audio(qkv, mask, H)
    assert torch.equal(o1, o2)
    o3 = ops.attention_audio(qkv[11:12].contiguous(), mask[11:12].contiguous(), H)
    assert torch.equal(o1[11:12], o3)
