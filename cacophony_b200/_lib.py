"""ctypes binding of libcaco_b200.so (C ABI declared in include/caco_b200.h).

The library is the product: if it is missing, cannot be loaded, or a call fails, this module raises — there
is no CPU or PyTorch fallback anywhere in the package.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcaco_b200.so")

# mirrors of the header's constants
EPI_BIAS_F16, EPI_BIAS_SILU_F16, EPI_BIAS_GELU_F16, EPI_BIAS_F32, EPI_BIAS_RESID_F32 = range(5)
GEMM_AUTO, GEMM_CG1_N256, GEMM_CG1_N128, GEMM_CG2_N256, GEMM_CG2_N256_E16 = range(5)

_ERR = {-1: "CACO_ERR_ARG (bad shape / null pointer / unsupported size)",
        -2: "CACO_ERR_ALIGN (pointer or leading dimension not 16-byte aligned)",
        -3: "CACO_ERR_DRIVER (cuTensorMapEncodeTiled unavailable or failed)",
        -4: "CACO_ERR_STATE (model not packed / missing tensor)"}


class CacoConfig(C.Structure):
    _fields_ = [("hidden", C.c_int), ("ffn", C.c_int), ("patch_dim", C.c_int), ("audio_layers", C.c_int),
                ("audio_heads", C.c_int), ("n_freq", C.c_int), ("pool_heads", C.c_int), ("text_layers", C.c_int),
                ("text_heads", C.c_int), ("vocab", C.c_int), ("max_pos", C.c_int), ("ln_eps", C.c_float),
                ("audio_ln_eps", C.c_float)]


_P, _I, _F, _L = C.c_void_p, C.c_int, C.c_float, C.c_int64

# name -> (restype, argtypes); every symbol include/caco_b200.h declares
SIGNATURES = {
    "caco_version": (_I, []),
    "caco_built_arch": (_I, []),
    "caco_frontend": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "caco_frontend_ragged": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "caco_gemm_f16": (_I, [_P, _I, _P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _P]),
    "caco_gemm_f16_wsplit": (_I, [_P, _I, _P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _P]),
    "caco_cast_f32_f16_split": (_I, [_P, _P, _L, _L, _P]),
    "caco_set_default_option": (_I, [C.c_char_p, _I]),
    "caco_saturation_count": (C.c_uint, [_I]),
    "caco_gemm_profile": (None, [_I]),
    "caco_gemm_profile_read": (_I, [C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "caco_cast_f32_f16": (_I, [_P, _P, _L, _P]),
    "caco_layernorm": (_I, [_P, _P, _P, _F, _P, _P, _I, _I, _P]),
    "caco_audio_add_pos": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "caco_attention_audio": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "caco_attention_text": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "caco_text_embed_ln": (_I, [_P, _P, _P, _P, _P, _P, _P, _F, _P, _P, _I, _I, _I, _I, _I, _P]),
    "caco_attn_pool": (_I, [_P, _P, _P, _P, _P, _P, _F, _P, _P, _I, _I, _I, _I, _P]),
    "caco_sgemm_nt": (_I, [_P, _I, _P, _I, _P, _F, _P, _I, _I, _I, _I, _P]),
    "caco_l2norm": (_I, [_P, _P, _I, _I, _F, _P]),
    "caco_sim_logits": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "caco_l2norm_scatter": (_I, [_P, _I, _I, _F, _P, C.c_longlong, C.c_longlong, _I, C.c_uint, _I, _P, _P]),
    "caco_wait_flags": (_I, [_P, _I, C.c_uint, _I, _P, _P]),
    "caco_model_create": (_I, [C.POINTER(CacoConfig), C.POINTER(_P)]),
    "caco_model_destroy": (None, [_P]),
    "caco_model_set_tensor": (_I, [_P, C.c_char_p, _P, _L]),
    "caco_model_pack": (_I, [_P, _P]),
    "caco_model_set_option": (_I, [_P, C.c_char_p, _I]),
    "caco_model_generation": (C.c_uint64, [_P]),
    "caco_model_audio_embedding": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P]),
    "caco_model_text_embedding": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P]),
    "caco_model_decoder_logits": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P, _P]),
    "caco_model_decoder_vocab": (_I, [_P]),
    "caco_model_decode_cache_bytes": (C.c_size_t, [_P, _I, _I, _I]),
    "caco_model_decode_begin": (_I, [_P, _P, C.c_size_t, _P, _P, _I, _I, _I, _P]),
    "caco_model_decode_step": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P]),
    "caco_attention_cross": (_I, [_P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "caco_model_encode_audio": (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "caco_model_encode_audio_ex": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "caco_model_logit_scale": (_P, [_P]),
    "caco_topk_rows": (_I, [_P, _I, _I, _I, _I, _P, _P, _P]),
    "caco_retrieval_hits": (_I, [_P, _I, _I, _P, _P, _P, _I, C.c_longlong, _I, _P, _P]),
    "caco_avg_pool_tokens": (_I, [_P, _I, _I, _I, _I, _P, _P]),
    "caco_launch_count": (_L, []),
    "caco_last_error": (C.c_char_p, []),
    "caco_mel_filterbank": (_I, [_P]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (once) and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m cacophony_b200.build` (needs nvcc, targets sm_100a). "
            "cacophony_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class CacoError(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    lib = load()
    detail = lib.caco_last_error().decode() if rc == -4 else ""
    if rc < 0:
        msg = _ERR.get(rc, f"error {rc}")
    else:
        msg = f"cudaError_t {rc}"
        try:
            import torch
            msg += f" ({torch.cuda.cudart().cudaGetErrorString(rc)})" if hasattr(torch.cuda.cudart(), "cudaGetErrorString") else ""
        except Exception:
            pass
    raise CacoError(f"{what}: {msg}{' — ' + detail if detail else ''}")


def ptr(t) -> Optional[int]:
    """Raw device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
