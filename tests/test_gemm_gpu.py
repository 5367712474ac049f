"""GPU: the tcgen05 GEMM (csrc/gemm.cu) through the C ABI against a plain PyTorch fp32 matmul of the same
fp16-rounded operands (TF32 off).  Tolerances: fp32 outputs differ only by accumulation order (1e-5 relative
Frobenius); fp16 outputs add one half-precision rounding (2^-11 relative per element)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from cacophony_b200 import _lib as L
from cacophony_b200 import ops

VARIANTS = {"cg1_n256": L.GEMM_CG1_N256, "cg1_n128": L.GEMM_CG1_N128, "cg2_n256": L.GEMM_CG2_N256,
            "cg2_e16": L.GEMM_CG2_N256_E16}


def _ref(a, w, bias, epi, resid):
    torch.backends.cuda.matmul.allow_tf32 = False
    y = a.float() @ w.float().t() + bias
    if epi == L.EPI_BIAS_RESID_F32:
        y = y + resid
    if epi == L.EPI_BIAS_SILU_F16:
        y = torch.nn.functional.silu(y)
    if epi == L.EPI_BIAS_GELU_F16:
        y = torch.nn.functional.gelu(y)
    return y


def _run(M, N, K, epi, variant, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = (torch.randn(M, K, device="cuda", generator=g)).half()
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).half()
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g) if epi == L.EPI_BIAS_RESID_F32 else None
    out = ops.gemm_f16(a, w, bias, epi, resid=resid, variant=variant)
    torch.cuda.synchronize()
    ref = _ref(a, w, bias, epi, resid)
    assert out.shape == ref.shape
    assert torch.isfinite(out.float()).all()
    err = (out.float() - ref).norm() / ref.norm()
    mx = (out.float() - ref).abs().max()
    if out.dtype == torch.float32:
        assert err < 1e-5 and mx < 1e-3, (float(err), float(mx))
    else:
        assert err < 6e-4 and mx < 2e-2, (float(err), float(mx))


@pytest.mark.parametrize("vname", list(VARIANTS))
@pytest.mark.parametrize("shape", [(128, 256, 64), (256, 768, 768), (1000, 2304, 768), (4096, 768, 3072), (77, 768, 256),
                                   (500 * 3, 3072, 768)])
def test_gemm_bias_f32(vname, shape):
    _run(*shape, L.EPI_BIAS_F32, VARIANTS[vname])


@pytest.mark.parametrize("vname", list(VARIANTS))
@pytest.mark.parametrize("epi", [L.EPI_BIAS_F16, L.EPI_BIAS_SILU_F16, L.EPI_BIAS_GELU_F16, L.EPI_BIAS_RESID_F32])
def test_gemm_epilogues(vname, epi):
    _run(1300, 768, 768, epi, VARIANTS[vname], seed=3)


@pytest.mark.parametrize("vname", list(VARIANTS))
def test_gemm_many_tiles_persistent(vname):
    """More tiles than CTAs: exercises the persistent loop, both accumulator stages and barrier phase wrap."""
    _run(128 * 70, 256 * 5, 64 * 9, L.EPI_BIAS_F32, VARIANTS[vname], seed=5)


@pytest.mark.parametrize("vname", list(VARIANTS))
def test_gemm_resid_in_place(vname):
    """out aliases resid (how the towers update the fp32 residual stream)."""
    M, N, K = 2048, 768, 768
    a = torch.randn(M, K, device="cuda").half()
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).half()
    bias = torch.randn(N, device="cuda")
    x = torch.randn(M, N, device="cuda")
    ref = _ref(a, w, bias, L.EPI_BIAS_RESID_F32, x.clone())
    ops.gemm_f16(a, w, bias, L.EPI_BIAS_RESID_F32, resid=x, variant=VARIANTS[vname], out=x)
    torch.cuda.synchronize()
    assert (x - ref).norm() / ref.norm() < 1e-5


def test_gemm_rejects_bad_arguments():
    a = torch.randn(128, 70, device="cuda").half()     # K % 8 != 0
    w = torch.randn(256, 70, device="cuda").half()
    with pytest.raises(L.CacoError):
        ops.gemm_f16(a, w, torch.zeros(256, device="cuda"), L.EPI_BIAS_F32)
    with pytest.raises(ValueError):
        ops.gemm_f16(a.float(), w, None, L.EPI_BIAS_F32)
    with pytest.raises(ValueError):
        ops.gemm_f16(a.cpu(), w, None, L.EPI_BIAS_F32)


@pytest.mark.parametrize("vname", list(VARIANTS))
@pytest.mark.parametrize("shape", [(256, 768, 768), (1000, 2304, 768), (333, 768, 3072), (512, 768, 256)])
def test_gemm_split_weights(vname, shape):
    """Split-weight form (precision mode): W enters as fp16 hi + lo along a doubled reduction, A's k-blocks are re-read.  The
    result must match the product of the fp16 activations with the EXACT fp32 weights up to the tensor core's fp32
    accumulation over the (doubled) reduction — under 1e-5 relative even at K = 3072 — and be an order of magnitude closer
    to it than the plain fp16-weight GEMM."""
    M, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).half()
    w = torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)
    bias = torch.randn(N, device="cuda", generator=g)
    w2 = ops.cast_f16_split(w.contiguous())
    assert w2.shape == (N, 2 * K)
    hi, lo = w2[:, :K].float(), w2[:, K:].float()
    assert float((hi + lo - w).abs().max()) <= float(w.abs().max()) * 2.0 ** -21
    out = ops.gemm_f16_wsplit(a, w2, bias, L.EPI_BIAS_F32, variant=VARIANTS[vname])
    plain = ops.gemm_f16(a, w.half(), bias, L.EPI_BIAS_F32, variant=VARIANTS[vname])
    ref = (a.double() @ w.double().t() + bias.double()).float()
    e_split = float((out - ref).norm() / ref.norm())
    e_plain = float((plain - ref).norm() / ref.norm())
    assert e_split < 1e-5 and e_split < e_plain / 10, (e_split, e_plain)


def test_per_handle_options_do_not_leak():
    """Execution options are per model handle (caco_model_set_option); the library defaults used by op-level calls stay
    untouched, unknown names are rejected."""
    lib = L.load()
    assert lib.caco_set_default_option(b"no_such_option", 1) == -1
    assert lib.caco_set_default_option(b"gemm_variant", 99) == -1
    assert lib.caco_set_default_option(b"pdl", 1) == 0
