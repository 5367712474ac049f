"""CPU: row (f-3) — the Flax -> torch checkpoint converter (cacophony_b200/checkpoint.py) and its msgpack container.

No flax checkpoint is available offline, so the converter is pinned by (1) a closed round trip: a torch ``state_dict`` with
the reference's exact keys/shapes -> the Flax parameter tree in the layout of ``src/caco/load_model.py:12-63`` (scan-stacked
RoBERTa layers, DenseGeneral attention kernels, ``nn.compact`` auto-names) -> flax's msgpack container on disk ->
``convert_caco_checkpoint`` -> identical tensors under identical keys; and (2) a direct check against the REFERENCE's own
torch modules: the converted tensors load with ``strict=True`` into ``src.caco_torch.create_caco_model()`` (when
/root/reference is mounted) and reproduce its forward bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch

from cacophony_b200 import checkpoint as ck
from oracle import weights as W

SPEC = dict(audio_layers=2, text_layers=3, vocab=500)


def _state_dict_with_decoder(seed=5):
    sd = dict(W.make_state_dict(seed, **SPEC))
    g = torch.Generator().manual_seed(seed)
    D, F = 768, 3072
    for i in range(2):
        p = f"decoder_module.encoder.layers.{i}."
        for blk in ("attention", "crossattention"):
            for nm in ("query", "key", "value"):
                sd[p + f"{blk}.self.{nm}.weight"] = torch.randn(D, D, generator=g) * 0.02
                sd[p + f"{blk}.self.{nm}.bias"] = torch.randn(D, generator=g) * 0.02
            sd[p + f"{blk}.output.dense.weight"] = torch.randn(D, D, generator=g) * 0.02
            sd[p + f"{blk}.output.dense.bias"] = torch.randn(D, generator=g) * 0.02
            sd[p + f"{blk}.output.LayerNorm.weight"] = torch.randn(D, generator=g)
            sd[p + f"{blk}.output.LayerNorm.bias"] = torch.randn(D, generator=g)
        sd[p + "intermediate.dense.weight"] = torch.randn(F, D, generator=g) * 0.02
        sd[p + "intermediate.dense.bias"] = torch.randn(F, generator=g) * 0.02
        sd[p + "output.dense.weight"] = torch.randn(D, F, generator=g) * 0.02
        sd[p + "output.dense.bias"] = torch.randn(D, generator=g) * 0.02
        sd[p + "output.LayerNorm.weight"] = torch.randn(D, generator=g)
        sd[p + "output.LayerNorm.bias"] = torch.randn(D, generator=g)
    sd["decoder_module.decoder_proj.weight"] = torch.randn(500, D, generator=g) * 0.02
    sd["decoder_module.decoder_proj.bias"] = torch.randn(500, generator=g) * 0.02
    return sd


@pytest.mark.parametrize("scan", [True, False])
def test_round_trip_through_flax_layout_and_msgpack(tmp_path, scan):
    sd = _state_dict_with_decoder()
    tree = ck.flax_tree_from_state_dict(sd, audio_heads=8, scan=scan)
    # the Flax layout really is what src/caco says it is
    lay = tree["text_module"]["encoder"]["layer"]
    if scan:
        assert list(lay) == ["ScanFlaxRobertaLayer_0"]
        assert lay["ScanFlaxRobertaLayer_0"]["attention"]["self"]["query"]["kernel"].shape == (3, 768, 768)   # layer axis first
        assert "crossattention" not in lay["ScanFlaxRobertaLayer_0"]
        assert tree["decoder_module"]["encoder"]["layer"]["ScanFlaxRobertaLayer_0"]["crossattention"]["self"]["key"]["kernel"].shape == (2, 768, 768)
    else:
        assert sorted(lay) == ["0", "1", "2"]
    att = tree["audio_module"]["AudioEncoderLayer_1"]["MultiHeadDotProductAttention_0"]
    assert att["query"]["kernel"].shape == (768, 8, 96) and att["query"]["bias"].shape == (8, 96)
    assert att["out"]["kernel"].shape == (8, 96, 768)
    assert tree["audio_attention_pool"]["Dense_0"]["kernel"].shape == (768, 1536)
    path = str(tmp_path / "Cacophony.ckpt")
    ck.write_flax_msgpack({"0": {"params": tree, "step": np.int32(7)}, "1": {"count": np.int64(3)}}, path, chunk_bytes=1 << 20)
    restored = ck.read_flax_msgpack(path)
    assert int(restored["0"]["step"]) == 7
    out = ck.convert_caco_checkpoint(path)
    assert set(out) == set(sd)
    for k in sd:
        assert out[k].dtype == torch.float32 and out[k].shape == sd[k].shape, k
        assert torch.equal(out[k], sd[k].float()), k
    # the three accepted container shapes
    for src in (restored, restored["0"], restored["0"]["params"]):
        again = ck.convert_caco_checkpoint(src, include_decoder=False)
        assert not any(k.startswith("decoder_module") for k in again) and torch.equal(again["text_proj.weight"], sd["text_proj.weight"])
    with pytest.raises(ValueError):
        ck.convert_caco_checkpoint({"params": {"foo": np.zeros(3)}})


def test_msgpack_container_edge_cases():
    tree = {"a": {"w": np.arange(12, dtype=np.float32).reshape(3, 4), "s": np.float32(2.5), "i": np.arange(5, dtype=np.int32)},
            "big": np.arange(70000, dtype=np.float32).reshape(700, 100), "n": 3, "name": "x"}
    blob = ck.write_flax_msgpack(tree, chunk_bytes=65536)          # forces the chunked-leaf form for 'big'
    back = ck.read_flax_msgpack(blob)
    assert np.array_equal(back["a"]["w"], tree["a"]["w"]) and back["a"]["w"].dtype == np.float32
    assert float(back["a"]["s"]) == 2.5 and np.array_equal(back["a"]["i"], tree["a"]["i"])
    assert np.array_equal(back["big"], tree["big"]) and back["big"].shape == (700, 100)
    assert back["n"] == 3 and back["name"] == "x"


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/caco_torch"), reason="reference not mounted (GPU box)")
def test_converted_checkpoint_loads_into_the_reference_model():
    """The converted state_dict is what the reference's own torch model expects: strict load + identical forward."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import ref_loader
    create_ref, _ = ref_loader.load()
    torch.manual_seed(0)
    ref = create_ref().eval()
    sd = ref.state_dict()
    tree = ck.flax_tree_from_state_dict(sd, audio_heads=8, scan=True)
    out = ck.convert_caco_checkpoint(ck.write_flax_msgpack({"0": {"params": tree}}))
    assert set(out) == set(sd)
    ref2 = create_ref().eval()
    ref2.load_state_dict(out, strict=True)
    g = torch.Generator().manual_seed(1)
    patches = torch.randn(2, 24, 256, generator=g)
    ti = torch.arange(24).float()[None].repeat(2, 1) // 8
    fi = torch.arange(24).float()[None].repeat(2, 1) % 8
    mask = torch.ones(2, 24)
    ids = torch.randint(3, 1000, (2, 9), generator=g)
    with torch.no_grad():
        a = ref(patches, ti, fi, mask, ids, torch.ones(2, 9, dtype=torch.long))
        b = ref2(patches, ti, fi, mask, ids, torch.ones(2, 9, dtype=torch.long))
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
