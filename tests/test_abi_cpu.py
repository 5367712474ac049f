"""CPU (no GPU calls): the C-ABI library is built, loads, exports every symbol include/caco_b200.h declares, and the
host-side mirrors behave like the reference's (state_dict keys, defaults, error behaviour)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import cacophony_b200 as cb
from cacophony_b200 import _lib as L
from cacophony_b200 import build as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    B.build()
    return L.load()


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "caco_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(caco_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(lib):
    syms = _declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in caco_b200.h but not exported"
        assert s in L.SIGNATURES, f"{s} has no ctypes prototype in cacophony_b200/_lib.py"
    assert lib.caco_version() == 100 and lib.caco_built_arch() == 100


def test_library_is_sm100a_tcgen05():
    """The shipped .so must contain sm_100a SASS with tensor-memory MMAs and TMA (B200_PROFILING.md mnemonics)."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    B.build()
    sass = subprocess.run(["cuobjdump", "-sass", L.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UTCHMMA" in sass, "no tcgen05.mma in the library"
    assert "UTMALDG" in sass, "no TMA loads in the library"
    assert "LDTM" in sass, "no tcgen05.ld in the library"


def test_mel_filterbank_matches_torchaudio_formula(lib):
    from oracle import caco_oracle as O
    fb = np.zeros((257, 128), np.float32)
    assert lib.caco_mel_filterbank(fb.ctypes.data_as(ctypes.c_void_p)) == 0
    ref = O.mel_filterbank().numpy()
    assert (fb != 0).sum() == 505 and not fb[:, 0].any()
    assert np.array_equal(fb != 0, ref != 0)
    assert np.abs(fb - ref).max() < 2e-5          # fp32 pow/linspace ulp differences only


def test_state_dict_keys_match_reference_layout():
    from oracle import weights as W
    m = cb.create_caco_model()
    want = {n: tuple(s) for n, s, _ in W.param_spec(decoder_layers=4)}      # the reference's 465 tensors
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == want and len(got) == 465
    if os.path.isdir("/root/reference/src/caco_torch"):                       # and literally the reference's own state_dict
        from oracle import ref_loader
        ref = ref_loader.load()[0]()
        assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == got
    enc = {k: torch.zeros(s) for k, s in want.items() if not k.startswith("decoder_module.")}
    m.load_state_dict(enc)                        # an encoder-only checkpoint leaves the captioning head as it is
    m.load_state_dict({k: torch.zeros(s) for k, s in want.items()})
    with pytest.raises(RuntimeError):
        m.load_state_dict({k: v for k, v in enc.items() if k != "text_proj.bias"})
    a_cfg = cb.AudioTransformerConfig(768, 12, 8, 3072, 256, 512, 8, 0.0, 0.0)
    headless = cb.CACO(a_cfg, cb.RobertaConfig(), cb.CACOConfig())            # decoder_config=None: caco.py:118-121
    assert headless.decoder_module is None and len(headless.state_dict()) == 465 - 106
    headless.load_state_dict({k: torch.zeros(s) for k, s in want.items()})    # decoder tensors of a checkpoint are dropped


def test_reference_signatures_and_defaults():
    import inspect
    m = cb.create_caco_model()
    a = inspect.signature(m.get_audio_embedding).parameters
    assert list(a) == ["audio_patches", "audio_time_inds", "audio_freq_inds", "audio_mask", "deterministic",
                       "return_hidden_state", "normalize"]
    assert a["return_hidden_state"].default is True and a["normalize"].default is False
    t = inspect.signature(m.get_text_embedding).parameters
    assert list(t) == ["text_input_ids", "text_mask", "position_ids", "deterministic", "return_hidden_state", "normalize"]
    f = inspect.signature(m.forward).parameters
    assert list(f) == ["audio_patches", "audio_time_inds", "audio_freq_inds", "audio_mask", "text_input_ids", "text_mask",
                       "deterministic"]
    assert cb.DatasetConfig().patches_seq_len == 512 and cb.CACOConfig().num_attention_pool_heads == 2
    assert cb.NORM_EPS == 1e-10


def test_no_cpu_fallback():
    m = cb.create_caco_model()
    with pytest.raises(RuntimeError, match="CUDA"):
        m.get_text_embedding(torch.zeros(1, 4, dtype=torch.long), torch.ones(1, 4))
    with pytest.raises(RuntimeError, match="CUDA"):
        cb.prepare_audio_batch(torch.zeros(1, 16000), cb.DatasetConfig(), "cpu")
    a_cfg = cb.AudioTransformerConfig(768, 1, 8, 3072, 256, 512, 8, 0.0, 0.0)
    headless = cb.CACO(a_cfg, cb.RobertaConfig(vocab_size=10, num_hidden_layers=1), cb.CACOConfig())
    with pytest.raises(ValueError, match="Decoder module not initialized"):
        headless.get_decoder_logits(None, None, None, None)
    with pytest.raises(ValueError, match="Decoder module not initialized"):
        headless.decode_begin(torch.zeros(1, 8, 768), torch.ones(1, 8), capacity=4)
    with pytest.raises(RuntimeError, match="CUDA"):
        m.decode_begin(torch.zeros(1, 8, 768), torch.ones(1, 8), capacity=4)
    with pytest.raises(RuntimeError, match="CUDA"):
        m.get_decoder_logits(torch.zeros(1, 8, 768), torch.ones(1, 8), torch.zeros(1, 4, dtype=torch.long), torch.ones(1, 4))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "cacophony_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("no CPU", ""), f"{f} mentions the oracle"


def test_header_is_plain_c_and_links_without_python(tmp_path):
    """include/caco_b200.h compiles as C99 and a plain-C program links libcaco_b200.so and gets answers from the
    host-side entry points (tests/c/abi_smoke.c) — the drop-in boundary has no torch / C++ types in it."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("gcc not on PATH")
    B.build()
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.dirname(L.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_smoke.c"),
           "-L" + libdir, "-lcaco_b200", "-Wl,-rpath," + libdir, "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "abi ok" in r.stdout, r.stdout + r.stderr
