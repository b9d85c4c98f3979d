"""Deterministic synthetic weights and inputs (test infrastructure).

A counter-based generator (splitmix64 over the flat element index, keyed by a stable hash of the
tensor name) so that the golden-vector script, the CPU oracle and the GPU tests all build bit-identical
tensors without shipping weight files.  The key sets are the reference's state-dict ABI (SURVEY.md
§8b; reference recipes/vanilla_vit.py:140-155, recipes/vanilla_bert.py:169-184) and are themselves
pinned by tests/golden/state_dict_keys.json, which was dumped from the reference's own modules.
"""
from __future__ import annotations

import zlib
from typing import Any, Dict, List, Tuple

import numpy as np

from .configs import is_vit, n_players

_U64 = np.uint64


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + _U64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> _U64(30))) * _U64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> _U64(27))) * _U64(0x94D049BB133111EB)
    return z ^ (z >> _U64(31))


def uniform(name: str, shape: Tuple[int, ...], seed: int = 0) -> np.ndarray:
    """float32 uniforms in [-1, 1), a pure function of (name, seed, flat index)."""
    n = int(np.prod(shape)) if len(shape) else 1
    with np.errstate(over="ignore"):
        key = _U64(zlib.crc32(name.encode("utf-8"))) * _U64(0x100000001B3) + _U64(seed) * _U64(0x9E3779B1)
        idx = np.arange(n, dtype=np.uint64) * _U64(0xD1342543DE82EF95) + key
        h = _splitmix64(idx)
    u = (h >> _U64(40)).astype(np.float64) / float(1 << 24)  # [0, 1)
    return (2.0 * u - 1.0).astype(np.float32).reshape(shape)


def randint(name: str, shape: Tuple[int, ...], lo: int, hi: int, seed: int = 0) -> np.ndarray:
    u = (uniform(name, shape, seed).astype(np.float64) + 1.0) * 0.5
    return (lo + np.floor(u * (hi - lo))).astype(np.int64).clip(lo, hi - 1)


# ------------------------------------------------------------------------------------------------
# state-dict shapes
# ------------------------------------------------------------------------------------------------
def _vit_layer_shapes(prefix: str, H: int, I: int, ln1: bool = True, ln2: bool = True) -> List[Tuple[str, Tuple[int, ...]]]:
    out = []
    for nm in ("query", "key", "value"):
        out += [(f"{prefix}.attention.self.{nm}.weight", (H, H)), (f"{prefix}.attention.self.{nm}.bias", (H,))]
    out += [(f"{prefix}.attention.output.dense.weight", (H, H)), (f"{prefix}.attention.output.dense.bias", (H,))]
    out += [(f"{prefix}.intermediate.dense.weight", (I, H)), (f"{prefix}.intermediate.dense.bias", (I,))]
    out += [(f"{prefix}.output.dense.weight", (H, I)), (f"{prefix}.output.dense.bias", (H,))]
    if ln1:
        out += [(f"{prefix}.layernorm_before.weight", (H,)), (f"{prefix}.layernorm_before.bias", (H,))]
    if ln2:
        out += [(f"{prefix}.layernorm_after.weight", (H,)), (f"{prefix}.layernorm_after.bias", (H,))]
    return out


def _bert_layer_shapes(prefix: str, H: int, I: int, ln1: bool = True, ln2: bool = True) -> List[Tuple[str, Tuple[int, ...]]]:
    out = []
    for nm in ("query", "key", "value"):
        out += [(f"{prefix}.attention.self.{nm}.weight", (H, H)), (f"{prefix}.attention.self.{nm}.bias", (H,))]
    out += [(f"{prefix}.attention.output.dense.weight", (H, H)), (f"{prefix}.attention.output.dense.bias", (H,))]
    if ln1:
        out += [(f"{prefix}.attention.output.LayerNorm.weight", (H,)), (f"{prefix}.attention.output.LayerNorm.bias", (H,))]
    out += [(f"{prefix}.intermediate.dense.weight", (I, H)), (f"{prefix}.intermediate.dense.bias", (I,))]
    out += [(f"{prefix}.output.dense.weight", (H, I)), (f"{prefix}.output.dense.bias", (H,))]
    if ln2:
        out += [(f"{prefix}.output.LayerNorm.weight", (H,)), (f"{prefix}.output.LayerNorm.bias", (H,))]
    return out


def backbone_shapes(cfg: Dict[str, Any]) -> List[Tuple[str, Tuple[int, ...]]]:
    H, I, L = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_hidden_layers"]
    out: List[Tuple[str, Tuple[int, ...]]] = []
    if is_vit(cfg):
        T = n_players(cfg) + 1
        P, Cin = cfg["img_patch_size"], cfg["img_channels"]
        out += [("vit.embeddings.cls_token", (1, 1, H)), ("vit.embeddings.position_embeddings", (1, T, H))]
        out += [("vit.embeddings.patch_embeddings.projection.weight", (H, Cin, P, P)),
                ("vit.embeddings.patch_embeddings.projection.bias", (H,))]
        for i in range(L):
            out += _vit_layer_shapes(f"vit.encoder.layers.{i}", H, I)
        out += [("vit.layernorm.weight", (H,)), ("vit.layernorm.bias", (H,))]
    else:
        out += [("bert.embeddings.word_embeddings.weight", (cfg["vocab_size"], H)),
                ("bert.embeddings.position_embeddings.weight", (cfg["max_position_embeddings"], H)),
                ("bert.embeddings.token_type_embeddings.weight", (cfg["type_vocab_size"], H)),
                ("bert.embeddings.LayerNorm.weight", (H,)), ("bert.embeddings.LayerNorm.bias", (H,))]
        for i in range(L):
            out += _bert_layer_shapes(f"bert.encoder.layers.{i}", H, I)
    return out


def surrogate_shapes(cfg: Dict[str, Any]) -> List[Tuple[str, Tuple[int, ...]]]:
    """Keys of VanillaViTClassifier/Surrogate (reference models/vanilla_vit.py:35-66) and
    VanillaBertClassifier/Surrogate (reference models/vanilla_bert.py:42-87)."""
    H, C = cfg["hidden_size"], cfg["num_labels"]
    out = backbone_shapes(cfg)
    if not is_vit(cfg):
        out += [("bert_pooler.dense.weight", (H, H)), ("bert_pooler.dense.bias", (H,))]
    out += [("classifier.weight", (C, H)), ("classifier.bias", (C,))]
    return out


def explainer_shapes(cfg: Dict[str, Any]) -> List[Tuple[str, Tuple[int, ...]]]:
    """Keys of VanillaViTExplainer (reference models/vanilla_vit.py:69-100) and VanillaBertExplainer
    (reference models/vanilla_bert.py:90-121)."""
    H, I, E, C = cfg["hidden_size"], cfg["intermediate_size"], cfg["explainer_head_hidden_size"], cfg["num_labels"]
    out = backbone_shapes(cfg)
    for i in range(cfg["explainer_attn_num_layers"]):
        if is_vit(cfg):
            out += _vit_layer_shapes(f"explainer_attn.{i}", H, I, ln1=(i != 0), ln2=True)
        else:
            out += _bert_layer_shapes(f"explainer_attn.{i}", H, I, ln1=(i != 0), ln2=True)
    if is_vit(cfg):
        out += [("explainer_mlp.0.weight", (H,)), ("explainer_mlp.0.bias", (H,)),
                ("explainer_mlp.1.weight", (E, H)), ("explainer_mlp.1.bias", (E,)),
                ("explainer_mlp.3.weight", (E, E)), ("explainer_mlp.3.bias", (E,)),
                ("explainer_mlp.5.weight", (C, E)), ("explainer_mlp.5.bias", (C,))]
    else:
        out += [("explainer_mlp.0.weight", (E, H)), ("explainer_mlp.0.bias", (E,)),
                ("explainer_mlp.2.weight", (E, E)), ("explainer_mlp.2.bias", (E,)),
                ("explainer_mlp.4.weight", (C, E)), ("explainer_mlp.4.bias", (C,))]
    return out


def _init_one(name: str, shape: Tuple[int, ...], seed: int) -> np.ndarray:
    u = uniform(name, shape, seed)
    leaf = name.rsplit(".", 1)[-1]
    if "LayerNorm" in name or "layernorm" in name or (name.startswith(("explainer_mlp.0.", "s_explainer_mlp.0.")) and len(shape) == 1):
        # LayerNorm affine: weight near 1, bias small (explainer_mlp.0 is a LayerNorm for ViT only,
        # where its parameters are 1-D; for BERT explainer_mlp.0.bias is a Linear bias — same scale)
        return (1.0 + 0.2 * u).astype(np.float32) if leaf == "weight" else (0.1 * u).astype(np.float32)
    if leaf == "bias":
        return (0.1 * u).astype(np.float32)
    if "embeddings" in name and len(shape) in (2, 3) and "projection" not in name:
        # cls/pos/word/type tables: unit variance like the reference's randn / nn.Embedding initialisers
        return (np.float32(np.sqrt(3.0)) * u).astype(np.float32)
    # nn.Linear / nn.Conv2d default initialiser of the reference's constructors: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
    fan_in = int(np.prod(shape[1:]))
    return (u * (1.0 / np.sqrt(fan_in))).astype(np.float32)


def state_like(shapes: Dict[str, Any], seed: int = 0) -> Dict[str, np.ndarray]:
    """Synthetic weights for an arbitrary key -> shape table (e.g. `{k: v.shape for k, v in module.state_dict().items()}`):
    used for the Froyo / LTT classes, whose tables are pinned against the reference by tests/golden/*_keys.json."""
    out = {name: _init_one(name, tuple(int(d) for d in shape), seed) for name, shape in shapes.items()}
    if "surrogate_null" in out:
        C = out["surrogate_null"].shape[-1]
        out["surrogate_null"] = ((uniform("surrogate_null", out["surrogate_null"].shape, seed) * np.float32(0.5) + np.float32(0.5))
                                 / np.float32(C)).astype(np.float32)
    return out


def make_state_dict(shapes: List[Tuple[str, Tuple[int, ...]]], seed: int = 0) -> Dict[str, np.ndarray]:
    return {name: _init_one(name, shape, seed) for name, shape in shapes}


def surrogate_state(cfg: Dict[str, Any], seed: int = 0) -> Dict[str, np.ndarray]:
    return make_state_dict(surrogate_shapes(cfg), seed)


def explainer_state(cfg: Dict[str, Any], seed: int = 1) -> Dict[str, np.ndarray]:
    return make_state_dict(explainer_shapes(cfg), seed)


def froyo_final_state(cfg: Dict[str, Any], seed: int = 1) -> Dict[str, np.ndarray]:
    """State dict of the Froyo bundle (reference models/froyo_vit.py:100-138, froyo_bert.py:105-158): backbone and explainer
    tail from explainer_state(seed), `classifier.*` (+ `bert_pooler.*`) from surrogate_state(seed + 10), the surrogate head
    `srg_*` from surrogate_state(seed + 20), and a non-trivial `surrogate_null`."""
    out = dict(explainer_state(cfg, seed))
    cls, srg = surrogate_state(cfg, seed + 10), surrogate_state(cfg, seed + 20)
    for k, v in cls.items():
        if k.startswith(("classifier.", "bert_pooler.")):
            out[k] = v
    for k, v in srg.items():
        if k.startswith(("classifier.", "bert_pooler.")):
            out["srg_" + k] = v
    C = cfg["num_labels"]
    out["surrogate_null"] = (uniform("froyo.surrogate_null", (1, C), seed) * np.float32(0.5) + np.float32(0.5)) / np.float32(C)
    return out


def inputs(cfg: Dict[str, Any], batch: int, seed: int = 0) -> np.ndarray:
    """ViT: (B,3,px,px) float32 images; BERT: (B,T) int64 ids with ids[:,0]=101."""
    if is_vit(cfg):
        px = cfg["img_px_size"]
        return uniform("inputs.images", (batch, cfg["img_channels"], px, px), seed) * np.float32(1.5)
    T = cfg["max_position_embeddings"]
    lo = min(1000, cfg["vocab_size"] // 2)
    ids = randint("inputs.ids", (batch, T), lo, cfg["vocab_size"], seed)
    ids[:, 0] = min(101, cfg["vocab_size"] - 1)
    return ids
