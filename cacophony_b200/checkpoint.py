"""Row (f-3): checkpoint compatibility — the JAX/Flax -> torch converter the reference exports but never ships.

``src/caco_torch/__init__.py:10`` lists ``convert_caco_checkpoint`` in ``__all__`` without defining it (the converter lives in a
git-ignored ``conversion_toolkit``).  This module restates it from the two parameter layouts themselves:

* Flax side (the released ``Cacophony.ckpt``): ``flax.training.checkpoints.restore_checkpoint(path, target=None)`` gives
  ``state['0']['params']`` (``src/caco/load_model.py:15-16``) with the module tree of ``src/caco/caco.py:56-70``,
  ``src/caco/audio_models/mae.py:55-143`` (``nn.compact`` auto-names: ``Dense_0``, ``AudioEncoderLayer_{i}``,
  ``MultiHeadDotProductAttention_0/{query,key,value,out}``, ``MLP_0/Dense_{0,1}``, ``LayerNorm_{0,1}``) and
  ``src/caco/text_models/roberta_text_model.py:92-603`` (``encoder/layer/ScanFlaxRobertaLayer_0/...`` with every leaf stacked
  along a leading layer axis by ``nn.scan``, ``:448-462``; or ``encoder/layer/{i}/...`` when built with ``scan=False``).
* torch side: the ``state_dict`` keys of ``src/caco_torch``, which ``cacophony_b200.CACO`` shares.

Layout rules: ``nn.Dense.kernel [in, out]`` -> ``Linear.weight = kernel.T``; ``LayerNorm.scale`` -> ``weight``;
``Embed.embedding`` -> ``weight``; Flax attention projections are ``DenseGeneral`` with ``kernel [hidden, heads, head_dim]``
(q/k/v, bias ``[heads, head_dim]``) and ``[heads, head_dim, hidden]`` (out): flattened, transposed and — for the audio tower's
``nn.MultiheadAttention`` — concatenated q|k|v into ``in_proj_weight / in_proj_bias``.

No flax / jax import is needed: the checkpoint container is msgpack with flax's three extension types
(``flax/serialization.py``: 1 = ndarray as (shape, dtype name, bytes), 2 = native complex, 3 = numpy scalar) plus its chunked
form for arrays above the msgpack size limit; ``read_flax_msgpack`` / ``write_flax_msgpack`` restate that container.
NOTE the two reference implementations disagree on one hyper-parameter the tensors cannot reveal: the JAX loader builds the
audio pooler with 8 heads (``load_model.py:47``), the torch port with 2 (``caco.py:292``); the parameter shapes are the same,
so pass the head count you want to ``CACOConfig(num_attention_pool_heads=...)`` (1..8 are supported).
"""
from __future__ import annotations

from typing import Any, Dict, Mapping, Optional, Union

import numpy as np
import torch

_EXT_NDARRAY, _EXT_COMPLEX, _EXT_NPSCALAR = 1, 2, 3
_CHUNK_KEY = "__msgpack_chunked_array__"


# ------------------------------------------------------------------------------------------------ msgpack container
def _ext_unpack(code: int, data: bytes):
    import msgpack
    if code == _EXT_NDARRAY:
        shape, dtype_name, buf = msgpack.unpackb(data, raw=True)
        name = dtype_name.decode() if isinstance(dtype_name, bytes) else dtype_name
        if name == "bfloat16":                                   # numpy has no bfloat16: widen through the bit pattern
            u16 = np.frombuffer(buf, dtype=np.uint16)
            return (u16.astype(np.uint32) << 16).view(np.float32).reshape(shape)
        return np.frombuffer(buf, dtype=np.dtype(name)).reshape(shape)
    if code == _EXT_COMPLEX:
        re, im = msgpack.unpackb(data)
        return complex(re, im)
    if code == _EXT_NPSCALAR:
        shape, dtype_name, buf = msgpack.unpackb(data, raw=True)
        name = dtype_name.decode() if isinstance(dtype_name, bytes) else dtype_name
        return np.frombuffer(buf, dtype=np.dtype(name)).reshape(shape)[()]
    import msgpack as _m
    return _m.ExtType(code, data)


def _unchunk(tree):
    if isinstance(tree, dict):
        if tree.get(_CHUNK_KEY):
            def seq(x):                                  # flax stores tuples as {'0': .., '1': ..} dicts
                return [x[str(i)] if str(i) in x else x[i] for i in range(len(x))] if isinstance(x, dict) else list(x)
            flat = np.concatenate([np.asarray(c).reshape(-1) for c in seq(tree["chunks"])])
            return flat.reshape(tuple(int(d) for d in seq(tree["shape"])))
        return {k: _unchunk(v) for k, v in tree.items()}
    return tree


def read_flax_msgpack(source: Union[str, bytes]) -> Dict[str, Any]:
    """A flax msgpack checkpoint (file path or bytes) -> nested dict of numpy arrays (what ``restore_checkpoint(path,
    target=None)`` returns, minus jax)."""
    import msgpack
    data = open(source, "rb").read() if isinstance(source, str) else source
    tree = msgpack.unpackb(data, ext_hook=_ext_unpack, raw=False, strict_map_key=False)
    return _unchunk(tree)


def write_flax_msgpack(tree: Mapping[str, Any], path: Optional[str] = None, chunk_bytes: int = 2 ** 30) -> bytes:
    """Inverse of `read_flax_msgpack` (flax's ``msgpack_serialize``): used by the round-trip tests and to hand converted or
    synthetic parameter trees back to JAX tooling."""
    import msgpack

    def enc_array(a: np.ndarray):
        a = np.ascontiguousarray(a)
        return msgpack.ExtType(_EXT_NDARRAY, msgpack.packb((list(a.shape), a.dtype.name, a.tobytes("C")), use_bin_type=True))

    def prep(node):
        if isinstance(node, Mapping):
            return {str(k): prep(v) for k, v in node.items()}
        if isinstance(node, torch.Tensor):
            node = node.detach().cpu().numpy()
        if isinstance(node, np.ndarray):
            if node.nbytes > chunk_bytes:                         # flax chunks leaves beyond msgpack's 2 GiB limit
                flat = node.reshape(-1)
                per = max(1, chunk_bytes // node.dtype.itemsize)
                chunks = {str(i): enc_array(flat[o:o + per]) for i, o in enumerate(range(0, flat.size, per))}
                return {_CHUNK_KEY: True, "shape": {str(i): int(d) for i, d in enumerate(node.shape)}, "chunks": chunks}
            return enc_array(node)
        if isinstance(node, np.generic):
            return msgpack.ExtType(_EXT_NPSCALAR, msgpack.packb(([], node.dtype.name, node.tobytes()), use_bin_type=True))
        return node
    blob = msgpack.packb(prep(tree), use_bin_type=True)
    if path is not None:
        with open(path, "wb") as f:
            f.write(blob)
    return blob


# ------------------------------------------------------------------------------------------------ layout conversion
def _t(x) -> torch.Tensor:
    return torch.from_numpy(np.array(x, dtype=np.float32, copy=True, order="C"))


def _dense(node, out: Dict[str, torch.Tensor], prefix: str) -> None:
    out[prefix + ".weight"] = _t(np.asarray(node["kernel"]).T)
    out[prefix + ".bias"] = _t(node["bias"])


def _layernorm(node, out: Dict[str, torch.Tensor], prefix: str) -> None:
    out[prefix + ".weight"] = _t(node["scale"])
    out[prefix + ".bias"] = _t(node["bias"])


def _general_in(node) -> tuple:
    """DenseGeneral projection into heads: kernel [hidden, heads, dh] -> Linear weight [heads*dh, hidden], bias [heads*dh]."""
    k = np.asarray(node["kernel"])
    return k.reshape(k.shape[0], -1).T, np.asarray(node["bias"]).reshape(-1)


def _roberta_layer(node, out: Dict[str, torch.Tensor], prefix: str, cross: bool) -> None:
    """One FlaxRobertaLayer (roberta_text_model.py:383-428) -> RobertaLayer keys (roberta.py:181-215)."""
    for blk in (("attention",) + (("crossattention",) if cross and "crossattention" in node else ())):
        a = node[blk]
        for nm in ("query", "key", "value"):
            _dense(a["self"][nm], out, f"{prefix}.{blk}.self.{nm}")
        _dense(a["output"]["dense"], out, f"{prefix}.{blk}.output.dense")
        _layernorm(a["output"]["LayerNorm"], out, f"{prefix}.{blk}.output.LayerNorm")
    _dense(node["intermediate"]["dense"], out, f"{prefix}.intermediate.dense")
    _dense(node["output"]["dense"], out, f"{prefix}.output.dense")
    _layernorm(node["output"]["LayerNorm"], out, f"{prefix}.output.LayerNorm")


def _index_tree(node, i: int):
    """Layer i of an nn.scan-stacked subtree (every leaf has the layer axis first, roberta_text_model.py:448-455)."""
    if isinstance(node, Mapping):
        return {k: _index_tree(v, i) for k, v in node.items()}
    return np.asarray(node)[i]


def _roberta_layers(layer_node, out: Dict[str, torch.Tensor], prefix: str, cross: bool) -> int:
    """``encoder/layer``: either {'ScanFlaxRobertaLayer_0': stacked} or {'0': ..., '1': ...}.  Returns the layer count."""
    scan_keys = [k for k in layer_node if str(k).startswith("Scan")]
    if scan_keys:
        stacked = layer_node[scan_keys[0]]
        leaf = stacked["intermediate"]["dense"]["bias"]
        n = int(np.asarray(leaf).shape[0])
        for i in range(n):
            _roberta_layer(_index_tree(stacked, i), out, f"{prefix}.{i}", cross)
        return n
    idx = sorted(int(k) for k in layer_node)
    for i in idx:
        _roberta_layer(layer_node[str(i)] if str(i) in layer_node else layer_node[i], out, f"{prefix}.{i}", cross)
    return len(idx)


def convert_caco_checkpoint(source: Union[str, bytes, Mapping[str, Any]], include_decoder: bool = True
                            ) -> Dict[str, torch.Tensor]:
    """Flax CACO parameters -> the torch ``state_dict`` of ``src/caco_torch`` / ``cacophony_b200.CACO``.

    source: a flax msgpack checkpoint (path or bytes), the restored state (``{'0': {'params': ...}}``,
    ``{'params': ...}``) or the parameter tree itself (keys ``audio_module``, ``text_module``, ``audio_attention_pool``,
    ``text_proj``, ``logit_scale`` [, ``decoder_module``]).  Returns fp32 CPU tensors; ``decoder_module.*`` keys are
    included when present (``CACO.load_state_dict`` ignores them: captioning is off this path)."""
    tree: Any = read_flax_msgpack(source) if isinstance(source, (str, bytes)) else source
    for key in ("0", 0, "params", "target", "model"):          # unwrap TrainState containers (load_model.py:16)
        while isinstance(tree, Mapping) and "audio_module" not in tree and key in tree:
            tree = tree[key]
    if not isinstance(tree, Mapping) or "audio_module" not in tree or "text_module" not in tree:
        raise ValueError("convert_caco_checkpoint: no CACO parameter tree found (expected audio_module / text_module)")
    sd: Dict[str, torch.Tensor] = {}
    sd["logit_scale"] = _t(np.asarray(tree["logit_scale"]).reshape(()))

    # ---- audio tower (mae.py:112-143 <- src/caco/audio_models/mae.py:105-143)
    am = tree["audio_module"]
    _dense(am["Dense_0"], sd, "audio_module.input_proj")
    sd["audio_module.freq_positional_embedding"] = _t(am["freq_positional_embedding"])
    n_layers = len([k for k in am if str(k).startswith("AudioEncoderLayer_")])
    for i in range(n_layers):
        ly = am[f"AudioEncoderLayer_{i}"]
        p = f"audio_module.layers.{i}"
        _layernorm(ly["LayerNorm_0"], sd, p + ".norm1")
        _layernorm(ly["LayerNorm_1"], sd, p + ".norm2")
        att = ly["MultiHeadDotProductAttention_0"]
        ws, bs = zip(*(_general_in(att[nm]) for nm in ("query", "key", "value")))
        sd[p + ".attn.in_proj_weight"] = _t(np.concatenate(ws, axis=0))          # q | k | v rows (nn.MultiheadAttention)
        sd[p + ".attn.in_proj_bias"] = _t(np.concatenate(bs, axis=0))
        ko = np.asarray(att["out"]["kernel"])                                      # [heads, dh, hidden]
        sd[p + ".attn.out_proj.weight"] = _t(ko.reshape(-1, ko.shape[-1]).T)
        sd[p + ".attn.out_proj.bias"] = _t(att["out"]["bias"])
        _dense(ly["MLP_0"]["Dense_0"], sd, p + ".mlp.fc1")
        _dense(ly["MLP_0"]["Dense_1"], sd, p + ".mlp.fc2")
    _layernorm(am["LayerNorm_0"], sd, "audio_module.norm")

    # ---- audio pooler (caco.py:24-79 <- src/caco/caco.py:19-54)
    ap = tree["audio_attention_pool"]
    sd["audio_attention_pool.query"] = _t(ap["query"])
    _dense(ap["Dense_0"], sd, "audio_attention_pool.kv_proj")
    _dense(ap["Dense_1"], sd, "audio_attention_pool.out_proj")

    # ---- text tower (roberta.py:26-326 <- roberta_text_model.py:92-583)
    tm = tree["text_module"]
    emb = tm["embeddings"]
    for nm in ("word_embeddings", "position_embeddings", "token_type_embeddings"):
        sd[f"text_module.embeddings.{nm}.weight"] = _t(emb[nm]["embedding"])
    _layernorm(emb["LayerNorm"], sd, "text_module.embeddings.LayerNorm")
    _roberta_layers(tm["encoder"]["layer"], sd, "text_module.encoder.layers", cross=False)
    sd["text_module.pooler.attention_pool_query"] = _t(tm["pooler"]["attention_pool_query"])
    _dense(tm["pooler"]["key_proj"], sd, "text_module.pooler.key_proj")
    _dense(tm["pooler"]["value_proj"], sd, "text_module.pooler.value_proj")
    _dense(tree["text_proj"], sd, "text_proj")

    # ---- captioning head (roberta.py:329-373 <- roberta_text_model.py:585-627): carried along, never executed here
    if include_decoder and "decoder_module" in tree:
        dm = tree["decoder_module"]
        _roberta_layers(dm["encoder"]["layer"], sd, "decoder_module.encoder.layers", cross=True)
        _dense(dm["decoder_proj"], sd, "decoder_module.decoder_proj")
    return sd


def flax_tree_from_state_dict(sd: Mapping[str, torch.Tensor], audio_heads: int = 8, scan: bool = True,
                              include_decoder: bool = True) -> Dict[str, Any]:
    """Inverse mapping (torch ``state_dict`` -> Flax parameter tree in the layout `convert_caco_checkpoint` reads): lets a model
    trained or edited on the torch side go back to the JAX tooling, and gives the converter a closed round trip to test."""
    def a(k):
        return sd[k].detach().cpu().numpy().astype(np.float32)

    def dense(prefix):
        return {"kernel": a(prefix + ".weight").T.copy(), "bias": a(prefix + ".bias")}

    def ln(prefix):
        return {"scale": a(prefix + ".weight"), "bias": a(prefix + ".bias")}
    D = a("audio_module.input_proj.weight").shape[0]
    dh = D // audio_heads
    am: Dict[str, Any] = {"Dense_0": dense("audio_module.input_proj"),
                          "freq_positional_embedding": a("audio_module.freq_positional_embedding"),
                          "LayerNorm_0": ln("audio_module.norm")}
    i = 0
    while f"audio_module.layers.{i}.norm1.weight" in sd:
        p = f"audio_module.layers.{i}"
        w, b = a(p + ".attn.in_proj_weight"), a(p + ".attn.in_proj_bias")
        att = {}
        for j, nm in enumerate(("query", "key", "value")):
            att[nm] = {"kernel": w[j * D:(j + 1) * D].T.reshape(D, audio_heads, dh).copy(),
                       "bias": b[j * D:(j + 1) * D].reshape(audio_heads, dh).copy()}
        att["out"] = {"kernel": a(p + ".attn.out_proj.weight").T.reshape(audio_heads, dh, D).copy(),
                      "bias": a(p + ".attn.out_proj.bias")}
        am[f"AudioEncoderLayer_{i}"] = {"LayerNorm_0": ln(p + ".norm1"), "LayerNorm_1": ln(p + ".norm2"),
                                        "MultiHeadDotProductAttention_0": att,
                                        "MLP_0": {"Dense_0": dense(p + ".mlp.fc1"), "Dense_1": dense(p + ".mlp.fc2")}}
        i += 1

    def roberta_layer(p, cross):
        node: Dict[str, Any] = {}
        for blk in ("attention",) + (("crossattention",) if cross else ()):
            node[blk] = {"self": {nm: dense(f"{p}.{blk}.self.{nm}") for nm in ("query", "key", "value")},
                         "output": {"dense": dense(f"{p}.{blk}.output.dense"), "LayerNorm": ln(f"{p}.{blk}.output.LayerNorm")}}
        node["intermediate"] = {"dense": dense(p + ".intermediate.dense")}
        node["output"] = {"dense": dense(p + ".output.dense"), "LayerNorm": ln(p + ".output.LayerNorm")}
        return node

    def stack(nodes):
        if isinstance(nodes[0], dict):
            return {k: stack([n[k] for n in nodes]) for k in nodes[0]}
        return np.stack(nodes, axis=0)

    def roberta_layers(prefix, cross):
        layers, j = [], 0
        while f"{prefix}.{j}.intermediate.dense.weight" in sd:
            layers.append(roberta_layer(f"{prefix}.{j}", cross))
            j += 1
        return {"ScanFlaxRobertaLayer_0": stack(layers)} if scan else {str(j): l for j, l in enumerate(layers)}
    tm = {"embeddings": {nm: {"embedding": a(f"text_module.embeddings.{nm}.weight")}
                         for nm in ("word_embeddings", "position_embeddings", "token_type_embeddings")},
          "encoder": {"layer": roberta_layers("text_module.encoder.layers", False)},
          "pooler": {"attention_pool_query": a("text_module.pooler.attention_pool_query"),
                     "key_proj": dense("text_module.pooler.key_proj"), "value_proj": dense("text_module.pooler.value_proj")}}
    tm["embeddings"]["LayerNorm"] = ln("text_module.embeddings.LayerNorm")
    tree: Dict[str, Any] = {"logit_scale": a("logit_scale"), "audio_module": am, "text_module": tm,
                            "audio_attention_pool": {"query": a("audio_attention_pool.query"),
                                                     "Dense_0": dense("audio_attention_pool.kv_proj"),
                                                     "Dense_1": dense("audio_attention_pool.out_proj")},
                            "text_proj": dense("text_proj")}
    if include_decoder and "decoder_module.decoder_proj.weight" in sd:
        tree["decoder_module"] = {"encoder": {"layer": roberta_layers("decoder_module.encoder.layers", True)},
                                  "decoder_proj": dense("decoder_module.decoder_proj")}
    return tree
