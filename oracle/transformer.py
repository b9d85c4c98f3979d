"""numpy restatement of the reference's masked ViT / BERT surrogate and explainer forward passes
(test infrastructure; see oracle/__init__.py).  Weights come in as a dict keyed by the reference's
state-dict names.  `dtype` selects the arithmetic (float32 mirrors the reference, float64 is used as
a high-precision yardstick in tolerance tests).
"""
from __future__ import annotations

import math
from typing import Any, Dict, Optional, Tuple

import numpy as np
from scipy.special import erf

from . import shapley as osh
from .configs import is_vit, n_players

Array = np.ndarray
State = Dict[str, Array]


# ------------------------------------------------------------------------------------------------
# primitives
# ------------------------------------------------------------------------------------------------
def linear(x: Array, sd: State, prefix: str) -> Array:
    return x @ sd[prefix + ".weight"].T + sd[prefix + ".bias"]


def layernorm(x: Array, sd: State, prefix: str, eps: float) -> Array:
    mean = x.mean(axis=-1, keepdims=True)
    xc = x - mean
    var = (xc * xc).mean(axis=-1, keepdims=True)
    return xc / np.sqrt(var + x.dtype.type(eps)) * sd[prefix + ".weight"] + sd[prefix + ".bias"]


def gelu(x: Array) -> Array:
    """nn.GELU() default = exact erf form (reference models/vanilla_vit.py:488, vanilla_bert.py:573)"""
    return (x * x.dtype.type(0.5)) * (x.dtype.type(1.0) + erf(x * x.dtype.type(1.0 / math.sqrt(2.0)))).astype(x.dtype)


def softmax(x: Array) -> Array:
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=-1, keepdims=True)


def _split_heads(x: Array, heads: int) -> Array:
    N, T, H = x.shape
    return x.reshape(N, T, heads, H // heads).transpose(0, 2, 1, 3)


def masked_attention(q: Array, k: Array, v: Array, token_mask: Array, heads: int, mode: str) -> Array:
    """q,k,v (N,T,H); token_mask (N,T) {0,1}.
    mode "mul0": ViT — scaled scores are MULTIPLIED by the key mask before softmax, so a masked key
      keeps logit 0 and a share exp(0)/Z of the weight (reference models/vanilla_vit.py:444-454).
    mode "neginf": BERT — additive (1-m)*finfo(dtype).min (HF get_extended_attention_mask; reference
      models/vanilla_bert.py:264-266, 520-523)."""
    N, T, H = q.shape
    d = H // heads
    dt = q.dtype
    qh, kh, vh = _split_heads(q, heads), _split_heads(k, heads), _split_heads(v, heads)
    scores = (qh @ kh.transpose(0, 1, 3, 2)) / dt.type(math.sqrt(d))
    m = token_mask.astype(dt)[:, None, None, :]
    if mode == "mul0":
        scores = scores * m
    elif mode == "neginf":
        scores = scores + (dt.type(1.0) - m) * np.finfo(dt).min
    else:
        raise ValueError(mode)
    probs = softmax(scores)
    ctx = probs @ vh
    return ctx.transpose(0, 2, 1, 3).reshape(N, T, H)


# ------------------------------------------------------------------------------------------------
# ViT (reference models/vanilla_vit.py)
# ------------------------------------------------------------------------------------------------
def vit_patchify(images: Array, patch: int) -> Array:
    """(N,C,px,px) -> (N, n_patches, C*patch*patch) in Conv2d weight order (c, py, px); patches are
    row-major over the image (reference models/vanilla_vit.py:279-284: conv -> flatten(2) -> transpose)."""
    N, C, Hh, Ww = images.shape
    gh, gw = Hh // patch, Ww // patch
    x = images.reshape(N, C, gh, patch, gw, patch).transpose(0, 2, 4, 1, 3, 5)
    return x.reshape(N, gh * gw, C * patch * patch)


def vit_embeddings(sd: State, cfg: Dict[str, Any], images: Array) -> Array:
    """reference models/vanilla_vit.py:242-253 (dropout is identity in eval())"""
    H = cfg["hidden_size"]
    w = sd["vit.embeddings.patch_embeddings.projection.weight"].reshape(H, -1)
    b = sd["vit.embeddings.patch_embeddings.projection.bias"]
    patches = vit_patchify(images, cfg["img_patch_size"]) @ w.T + b
    cls = np.broadcast_to(sd["vit.embeddings.cls_token"], (images.shape[0], 1, H))
    return np.concatenate([cls, patches], axis=1) + sd["vit.embeddings.position_embeddings"]


def vit_layer(sd: State, prefix: str, x: Array, token_mask: Array, heads: int, eps: float, ln1: bool = True) -> Array:
    """pre-LN block, reference models/vanilla_vit.py:364-377"""
    h = layernorm(x, sd, prefix + ".layernorm_before", eps) if ln1 else x
    q = linear(h, sd, prefix + ".attention.self.query")
    k = linear(h, sd, prefix + ".attention.self.key")
    v = linear(h, sd, prefix + ".attention.self.value")
    ctx = masked_attention(q, k, v, token_mask, heads, "mul0")
    x = x + linear(ctx, sd, prefix + ".attention.output.dense")
    h = layernorm(x, sd, prefix + ".layernorm_after", eps)
    h = gelu(linear(h, sd, prefix + ".intermediate.dense"))
    return linear(h, sd, prefix + ".output.dense") + x


def vit_backbone(sd: State, cfg: Dict[str, Any], images: Array, token_mask: Array) -> Array:
    """reference models/vanilla_vit.py:207-214"""
    x = vit_embeddings(sd, cfg, images)
    for i in range(cfg["num_hidden_layers"]):
        x = vit_layer(sd, f"vit.encoder.layers.{i}", x, token_mask, cfg["num_attention_heads"], cfg["layer_norm_eps"])
    return layernorm(x, sd, "vit.layernorm", cfg["layer_norm_eps"])


def vit_surrogate(sd: State, cfg: Dict[str, Any], images: Array, token_mask: Array) -> Array:
    """reference models/vanilla_vit.py:51-56 — class PROBABILITIES (Softmax head)."""
    h = vit_backbone(sd, cfg, images, token_mask)
    return softmax(linear(h[:, 0, :], sd, "classifier"))


def vit_explainer_pred(sd: State, cfg: Dict[str, Any], images: Array, token_mask: Array) -> Array:
    """reference models/vanilla_vit.py:119-123 — per-token head output (B,T,C) before normalisation."""
    x = vit_backbone(sd, cfg, images, token_mask)
    for i in range(cfg["explainer_attn_num_layers"]):
        x = vit_layer(sd, f"explainer_attn.{i}", x, token_mask, cfg["num_attention_heads"], cfg["layer_norm_eps"], ln1=(i != 0))
    x = layernorm(x, sd, "explainer_mlp.0", 1e-5)  # nn.LayerNorm default eps (reference l.94)
    x = gelu(linear(x, sd, "explainer_mlp.1"))
    x = gelu(linear(x, sd, "explainer_mlp.3"))
    return linear(x, sd, "explainer_mlp.5")


# ------------------------------------------------------------------------------------------------
# BERT (reference models/vanilla_bert.py)
# ------------------------------------------------------------------------------------------------
def bert_embeddings(sd: State, cfg: Dict[str, Any], ids: Array, token_type_ids: Optional[Array] = None) -> Array:
    """reference models/vanilla_bert.py:307-325"""
    N, T = ids.shape
    tt = np.zeros_like(ids) if token_type_ids is None else token_type_ids
    x = sd["bert.embeddings.word_embeddings.weight"][ids] + sd["bert.embeddings.token_type_embeddings.weight"][tt]
    x = x + sd["bert.embeddings.position_embeddings.weight"][:T][None, :, :]
    return layernorm(x, sd, "bert.embeddings.LayerNorm", cfg["layer_norm_eps"])


def bert_layer(sd: State, prefix: str, x: Array, token_mask: Array, heads: int, eps: float, ln1: bool = True) -> Array:
    """post-LN block, reference models/vanilla_bert.py:396-427, 556-560, 600-604"""
    q = linear(x, sd, prefix + ".attention.self.query")
    k = linear(x, sd, prefix + ".attention.self.key")
    v = linear(x, sd, prefix + ".attention.self.value")
    ctx = masked_attention(q, k, v, token_mask, heads, "neginf")
    a = linear(ctx, sd, prefix + ".attention.output.dense") + x
    if ln1:
        a = layernorm(a, sd, prefix + ".attention.output.LayerNorm", eps)
    h = gelu(linear(a, sd, prefix + ".intermediate.dense"))
    return layernorm(linear(h, sd, prefix + ".output.dense") + a, sd, prefix + ".output.LayerNorm", eps)


def bert_backbone(sd: State, cfg: Dict[str, Any], ids: Array, token_mask: Array) -> Array:
    """reference models/vanilla_bert.py:253-271"""
    x = bert_embeddings(sd, cfg, ids)
    for i in range(cfg["num_hidden_layers"]):
        x = bert_layer(sd, f"bert.encoder.layers.{i}", x, token_mask, cfg["num_attention_heads"], cfg["layer_norm_eps"])
    return x


def bert_surrogate(sd: State, cfg: Dict[str, Any], ids: Array, token_mask: Array) -> Array:
    """reference models/vanilla_bert.py:61-77 + pooler 615-619"""
    h = bert_backbone(sd, cfg, ids, token_mask)
    pooled = np.tanh(linear(h[:, 0, :], sd, "bert_pooler.dense"))
    return softmax(linear(pooled, sd, "classifier"))


def bert_explainer_pred(sd: State, cfg: Dict[str, Any], ids: Array, token_mask: Array) -> Array:
    """reference models/vanilla_bert.py:139-154"""
    x = bert_backbone(sd, cfg, ids, token_mask)
    for i in range(cfg["explainer_attn_num_layers"]):
        x = bert_layer(sd, f"explainer_attn.{i}", x, token_mask, cfg["num_attention_heads"], cfg["layer_norm_eps"], ln1=(i != 0))
    x = gelu(linear(x, sd, "explainer_mlp.0"))
    x = gelu(linear(x, sd, "explainer_mlp.2"))
    return linear(x, sd, "explainer_mlp.4")


# ------------------------------------------------------------------------------------------------
# recipe-shaped entry points (reference recipes/vanilla_vit.py:227-261, recipes/vanilla_bert.py:293-329)
# ------------------------------------------------------------------------------------------------
def _cast_state(sd: State, dtype) -> State:
    return {k: (v.astype(dtype) if v.dtype.kind == "f" else v) for k, v in sd.items()}


def fw_surrogate(sd: State, cfg: Dict[str, Any], xs: Array, player_mask: Array, dtype=np.float32, chunk: int = 16) -> Array:
    """(N,...) inputs + (N,n) player masks -> (N,C) probabilities; one mask row per input row."""
    sd = _cast_state(sd, dtype)
    token_mask = osh.prepend_cls(np.asarray(player_mask))
    outs = []
    for i in range(0, xs.shape[0], chunk):
        x = xs[i:i + chunk]
        tm = token_mask[i:i + chunk]
        if is_vit(cfg):
            outs.append(vit_surrogate(sd, cfg, x.astype(dtype), tm))
        else:
            outs.append(bert_surrogate(sd, cfg, x, tm))
    return np.concatenate(outs, axis=0)


def fw_explainer(sd: State, cfg: Dict[str, Any], xs: Array, player_mask: Array, grand: Array, null: Array,
                 dtype=np.float32, chunk: int = 16) -> Tuple[Array, Array]:
    """-> (phi (B,C,n), pred (B,T,C))"""
    sd = _cast_state(sd, dtype)
    token_mask = osh.prepend_cls(np.asarray(player_mask))
    preds = []
    for i in range(0, xs.shape[0], chunk):
        x = xs[i:i + chunk]
        tm = token_mask[i:i + chunk]
        if is_vit(cfg):
            preds.append(vit_explainer_pred(sd, cfg, x.astype(dtype), tm))
        else:
            preds.append(bert_explainer_pred(sd, cfg, x, tm))
    pred = np.concatenate(preds, axis=0)
    phi = osh.explainer_output(pred, grand.astype(dtype), null.astype(dtype), cfg["explainer_normalize"])
    return phi, pred


def null_input(cfg: Dict[str, Any]) -> Array:
    """ViT: zero image (reference recipes/vanilla_vit.py:213-216).  BERT: tokenised "" padded to T =
    [CLS]=101, [SEP]=102, [PAD]=0... (reference recipes/vanilla_bert.py:265-278; ids clipped to the
    vocabulary for the reduced test configs)."""
    if is_vit(cfg):
        px = cfg["img_px_size"]
        return np.zeros((1, cfg["img_channels"], px, px), dtype=np.float32)
    T = cfg["max_position_embeddings"]
    ids = np.zeros((1, T), dtype=np.int64)
    ids[0, 0] = min(101, cfg["vocab_size"] - 1)
    ids[0, 1] = min(102, cfg["vocab_size"] - 1)
    return ids


def flops_per_eval(cfg: Dict[str, Any]) -> float:
    """Dense forward FLOPs of one masked surrogate evaluation (SURVEY.md §8d / BASELINE.md §4)."""
    H, I, L, C = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_hidden_layers"], cfg["num_labels"]
    T = n_players(cfg) + 1
    per_layer = 2 * T * H * (3 * H) + 2 * T * H * H + 2 * 2 * T * H * I + 4 * T * T * H
    total = L * per_layer + 2 * H * C
    if is_vit(cfg):
        total += 2 * (T - 1) * H * cfg["img_channels"] * cfg["img_patch_size"] ** 2
    else:
        total += 2 * H * H
    return float(total)
