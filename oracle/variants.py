"""CPU restatement (numpy) of the reference's model VARIANTS over the vanilla blocks of oracle/transformer.py — test
infrastructure, like the rest of oracle/ (only tests/, smoke() and bench.py's CPU legs may import it).

  * Froyo  — reference models/froyo_vit.py:100-171, models/froyo_bert.py:105-204: one frozen backbone pass feeds the
             classifier head, the surrogate head (srg_*) and the explainer tail;
  * LTT    — reference models/ltt_vit.py:290-440, models/ltt_bert.py:352-499: every backbone block's output h_i feeds a narrow
             side ladder, s <- block_i^side(s + GELU(W_i h_i)); surrogate = ladder 0 + side classifier, explainer = ladder + side
             explainer; the bundle carries two ladders (0 = surrogate, 1 = explainer);
  * Duo    — reference models/duo_vanilla_vit.py:75-137, models/duo_vanilla_bert.py:79-150: the explainer also carries a
             classification head on its own backbone.

Pinned by tests/test_oracle_variants.py against outputs of the reference's own classes (tests/golden/make_golden.py
froyo | ltt | duo)."""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from . import shapley as osh
from . import transformer as otr
from .configs import is_vit

Array = np.ndarray
State = Dict[str, Array]


def _backbone_states(sd: State, cfg: Dict[str, Any], xs: Array, token_mask: Array) -> List[Array]:
    """outputs of every backbone block (ViT: BEFORE vit.layernorm)"""
    vit = is_vit(cfg)
    x = otr.vit_embeddings(sd, cfg, xs) if vit else otr.bert_embeddings(sd, cfg, xs)
    root = "vit" if vit else "bert"
    out = []
    for i in range(cfg["num_hidden_layers"]):
        layer = otr.vit_layer if vit else otr.bert_layer
        x = layer(sd, f"{root}.encoder.layers.{i}", x, token_mask, cfg["num_attention_heads"], cfg["layer_norm_eps"])
        out.append(x)
    return out


def _main_probs(sd: State, cfg: Dict[str, Any], h_last: Array, head: str = "classifier", pooler: str = "bert_pooler") -> Array:
    if is_vit(cfg):
        h = otr.layernorm(h_last, sd, "vit.layernorm", cfg["layer_norm_eps"])
        return otr.softmax(otr.linear(h[:, 0, :], sd, head))
    return otr.softmax(otr.linear(np.tanh(otr.linear(h_last[:, 0, :], sd, pooler + ".dense")), sd, head))


def _tail_pred(sd: State, cfg: Dict[str, Any], x: Array, token_mask: Array, attn: str, mlp: str, n_attn: int) -> Array:
    """explainer tail on a (B, T, width) state: n_attn blocks (first one without its first LayerNorm) + the 3-layer head"""
    vit = is_vit(cfg)
    layer = otr.vit_layer if vit else otr.bert_layer
    for i in range(n_attn):
        x = layer(sd, f"{attn}.{i}", x, token_mask, cfg["num_attention_heads"], cfg["layer_norm_eps"], ln1=(i != 0))
    if vit:
        x = otr.layernorm(x, sd, mlp + ".0", 1e-5)
        names = (mlp + ".1", mlp + ".3", mlp + ".5")
    else:
        names = (mlp + ".0", mlp + ".2", mlp + ".4")
    x = otr.gelu(otr.linear(x, sd, names[0]))
    x = otr.gelu(otr.linear(x, sd, names[1]))
    return otr.linear(x, sd, names[2])


# ------------------------------------------------------------------------------------------------
# Froyo
# ------------------------------------------------------------------------------------------------
def froyo_final(sd: State, cfg: Dict[str, Any], xs: Array, player_mask: Array) -> Tuple[Array, Array]:
    """-> (class probabilities (B, C), phi (B, C, n))"""
    tm = osh.prepend_cls(np.asarray(player_mask))
    h = _backbone_states(sd, cfg, xs.astype(np.float32) if is_vit(cfg) else xs, tm)[-1]
    cls = _main_probs(sd, cfg, h)
    x = otr.layernorm(h, sd, "vit.layernorm", cfg["layer_norm_eps"]) if is_vit(cfg) else h
    pred = _tail_pred(sd, cfg, x, tm, "explainer_attn", "explainer_mlp", cfg["explainer_attn_num_layers"])
    if cfg["explainer_normalize"]:
        grand = _main_probs(sd, cfg, h, head="srg_classifier", pooler="srg_bert_pooler")
        return cls, osh.explainer_output(pred, grand, sd["surrogate_null"], True)
    return cls, osh.explainer_output(pred, None, None, False)


# ------------------------------------------------------------------------------------------------
# LTT
# ------------------------------------------------------------------------------------------------
def _ladder(sd: State, cfg: Dict[str, Any], states: List[Array], token_mask: Array, b: int, freeze_layer: Optional[int] = None) -> Array:
    vit = is_vit(cfg)
    root = "vit" if vit else "bert"
    layer = otr.vit_layer if vit else otr.bert_layer
    stop = len(states) if freeze_layer is None else max(1, min(len(states), int(freeze_layer)))
    s: Any = 0.0
    for i, h in enumerate(states[:stop]):
        s = s + otr.gelu(otr.linear(h, sd, f"{root}.encoder.s_attn_maps.{b}_{i}"))
        s = layer(sd, f"{root}.encoder.s_attn_layers.{b}_{i}", s, token_mask, cfg["num_attention_heads"], cfg["layer_norm_eps"])
    return otr.layernorm(s, sd, f"vit.s_attn_layernorm.{b}", cfg["layer_norm_eps"]) if vit else s


def _side_probs(sd: State, cfg: Dict[str, Any], s: Array) -> Array:
    if is_vit(cfg):
        return otr.softmax(otr.linear(s[:, 0, :], sd, "s_attn_classifier"))
    return otr.softmax(otr.linear(np.tanh(otr.linear(s[:, 0, :], sd, "bert_s_attn_pooler.dense")), sd, "s_attn_classifier"))


def _ltt_names(cfg: Dict[str, Any]) -> Tuple[str, str]:
    return ("s_explainer_attn", "s_explainer_mlp") if is_vit(cfg) else ("s_attn_attention_layers", "s_attn_explainer")


def ltt_surrogate(sd: State, cfg: Dict[str, Any], xs: Array, player_mask: Array) -> Tuple[Array, Array]:
    """-> (side-ladder probabilities, backbone probabilities)"""
    tm = osh.prepend_cls(np.asarray(player_mask))
    states = _backbone_states(sd, cfg, xs.astype(np.float32) if is_vit(cfg) else xs, tm)
    return _side_probs(sd, cfg, _ladder(sd, cfg, states, tm, 0)), _main_probs(sd, cfg, states[-1])


def ltt_explainer(sd: State, cfg: Dict[str, Any], xs: Array, player_mask: Array, grand: Array, null: Array) -> Tuple[Array, Array]:
    """-> (phi (B, C, n), backbone probabilities)"""
    tm = osh.prepend_cls(np.asarray(player_mask))
    states = _backbone_states(sd, cfg, xs.astype(np.float32) if is_vit(cfg) else xs, tm)
    attn, mlp = _ltt_names(cfg)
    pred = _tail_pred(sd, cfg, _ladder(sd, cfg, states, tm, 0), tm, attn, mlp, cfg["explainer_s_attn_num_layers"])
    return osh.explainer_output(pred, grand, null, cfg["explainer_normalize"]), _main_probs(sd, cfg, states[-1])


def ltt_final(sd: State, cfg: Dict[str, Any], xs: Array, player_mask: Array) -> Tuple[Array, Array]:
    """-> (backbone probabilities, phi): ladder 0 supplies `grand`, ladder 1 the attributions"""
    tm = osh.prepend_cls(np.asarray(player_mask))
    states = _backbone_states(sd, cfg, xs.astype(np.float32) if is_vit(cfg) else xs, tm)
    attn, mlp = _ltt_names(cfg)
    pred = _tail_pred(sd, cfg, _ladder(sd, cfg, states, tm, 1), tm, attn, mlp, cfg["explainer_s_attn_num_layers"])
    cls = _main_probs(sd, cfg, states[-1])
    if cfg["explainer_normalize"]:
        grand = _side_probs(sd, cfg, _ladder(sd, cfg, states, tm, 0))
        return cls, osh.explainer_output(pred, grand, sd["surrogate_null"], True)
    return cls, osh.explainer_output(pred, None, None, False)


# ------------------------------------------------------------------------------------------------
# Duo
# ------------------------------------------------------------------------------------------------
def duo_explainer(sd: State, cfg: Dict[str, Any], xs: Array, player_mask: Array, grand: Array, null: Array) -> Tuple[Array, Array]:
    """-> (phi, class output): ViT softmax probabilities, BERT raw logits (as the reference returns them)"""
    tm = osh.prepend_cls(np.asarray(player_mask))
    h = _backbone_states(sd, cfg, xs.astype(np.float32) if is_vit(cfg) else xs, tm)[-1]
    if is_vit(cfg):
        x = otr.layernorm(h, sd, "vit.layernorm", cfg["layer_norm_eps"])
        cls = otr.softmax(otr.linear(x[:, 0, :], sd, "classifier"))
    else:
        x = h
        cls = otr.linear(np.tanh(otr.linear(h[:, 0, :], sd, "bert_pooler.dense")), sd, "classifier")
    pred = _tail_pred(sd, cfg, x, tm, "explainer_attn", "explainer_mlp", cfg["explainer_attn_num_layers"])
    return osh.explainer_output(pred, grand, null, cfg["explainer_normalize"]), cls
