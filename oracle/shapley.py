"""numpy restatement of reference models/shapley.py (test infrastructure; see oracle/__init__.py).

Integer outputs (masks, packed words) are compared bit-exactly with the reference and with the CUDA
kernels; floating-point outputs with the tolerance written in the tests.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


# ------------------------------------------------------------------------------------------------
# a1: paired Shapley-kernel coalition sampler
# ------------------------------------------------------------------------------------------------
def shapley_size_probs(n_players: int) -> np.ndarray:
    """reference models/shapley.py:65-67 — p[k'] ∝ 1 / ((k'+1)(n-k'-1)), k' = 0..n-2, float32."""
    a = np.arange(1, n_players, dtype=np.int64)
    probs = (a * (n_players - a)).astype(np.float32)
    probs = np.float32(1.0) / probs
    return (probs / probs.sum(dtype=np.float32)).astype(np.float32)


def shapley_size_prefix(n_players: int) -> np.ndarray:
    """reference models/shapley.py:132 — exclusive prefix `cumsum(p) - p` (sequential float32)."""
    p = shapley_size_probs(n_players)
    return (np.cumsum(p, dtype=np.float32) - p).astype(np.float32)


def masks_from_uniforms(u_players: np.ndarray, u_size: np.ndarray, prefix: np.ndarray, n_players: int) -> np.ndarray:
    """reference models/shapley.py:56-79 + 131-135 with the uniforms made explicit.

    u_players: (P, n) float32 = `masks_1` (l.69); u_size: (P,) float32 = the `torch.rand` drawn
    inside `_torch_choice` (l.133); prefix: (n-1,) float32 (l.132).  Returns (2P, n) int64 with the
    complement of row 2i in row 2i+1 (l.75-78).
    """
    u_players = np.asarray(u_players, dtype=np.float32)
    u_size = np.asarray(u_size, dtype=np.float32).reshape(-1, 1)
    prefix = np.asarray(prefix, dtype=np.float32).reshape(1, -1)
    count = (u_size >= prefix).sum(axis=1)                       # l.134
    position = np.maximum(count - 1, 0).astype(np.int64)         # l.134: max(.., 0)
    # l.70: (1 / n_players) * a[position]; python float scalar times int64 tensor -> float32 product
    thresh = (np.float32(1.0 / n_players) * position.astype(np.float32)).astype(np.float32).reshape(-1, 1)
    m = (u_players > thresh).astype(np.int64)                    # l.73
    out = np.stack([m, 1 - m], axis=1).reshape(2 * m.shape[0], n_players)  # l.75-78
    return out


def mask_purely_uniform_from_uniforms(u_players: np.ndarray, u_row: np.ndarray) -> np.ndarray:
    """reference models/shapley.py:109-115"""
    return (np.asarray(u_players, np.float32) > np.asarray(u_row, np.float32).reshape(-1, 1)).astype(np.int64)


# ------------------------------------------------------------------------------------------------
# a2: CLS column + bit packing (layout contract of include/autognothi_b200.h)
# ------------------------------------------------------------------------------------------------
def prepend_cls(mask: np.ndarray) -> np.ndarray:
    """reference recipes/vanilla_vit.py:219-224 / recipes/vanilla_bert.py:281-290"""
    ones = np.ones((mask.shape[0], 1), dtype=mask.dtype)
    return np.concatenate([ones, mask], axis=1)


def packed_words(n_tokens: int) -> int:
    return (n_tokens + 31) // 32


def pack_token_mask(token_mask: np.ndarray) -> np.ndarray:
    """(rows, T) {0,1} -> (rows, ceil(T/32)) uint32, bit t%32 of word t//32 = token t."""
    rows, T = token_mask.shape
    W = packed_words(T)
    padded = np.zeros((rows, W * 32), dtype=np.uint64)
    padded[:, :T] = token_mask != 0
    weights = (np.uint64(1) << np.arange(32, dtype=np.uint64))
    return (padded.reshape(rows, W, 32) * weights).sum(axis=2).astype(np.uint32)


def pack_player_mask(player_mask: np.ndarray) -> np.ndarray:
    return pack_token_mask(prepend_cls(np.asarray(player_mask)))


def unpack_token_mask(packed: np.ndarray, n_tokens: int) -> np.ndarray:
    rows, W = packed.shape
    bits = (packed[:, :, None].astype(np.uint64) >> np.arange(32, dtype=np.uint64)) & np.uint64(1)
    return bits.reshape(rows, W * 32)[:, :n_tokens].astype(np.int64)


# ------------------------------------------------------------------------------------------------
# a11: additive efficiency normalisation
# ------------------------------------------------------------------------------------------------
def normalize_shapley_explanation(pred: np.ndarray, grand: np.ndarray, null: np.ndarray) -> np.ndarray:
    """reference models/shapley.py:82-93.  pred (B,T,C) — callers pass the un-sliced tensor, so the
    divisor is T = n_players + 1 (reference models/vanilla_vit.py:125-128)."""
    B, T, _ = pred.shape
    grand = grand[:, None, :]
    null = np.broadcast_to(null.reshape(1, 1, -1), (B, 1, null.size))
    diff = (grand - null) - pred.sum(axis=1, keepdims=True)
    return pred + diff / pred.dtype.type(T)


def explainer_output(pred: np.ndarray, grand: np.ndarray, null: np.ndarray, normalize: bool = True) -> np.ndarray:
    """normalise -> drop CLS -> (B,C,n)  (reference models/vanilla_vit.py:123-130)"""
    out = normalize_shapley_explanation(pred, grand, null) if normalize else pred
    return np.ascontiguousarray(out[:, 1:, :].transpose(0, 2, 1))


def explainer_output_grad(dphi: np.ndarray, n_tokens: int, normalize: bool = True) -> np.ndarray:
    """Adjoint of explainer_output w.r.t. pred: dpred (B,T,C) from dphi (B,C,n)."""
    B, C, n = dphi.shape
    d = np.zeros((B, n_tokens, C), dtype=dphi.dtype)
    d[:, 1:, :] = dphi.transpose(0, 2, 1)
    if normalize:
        d = d - d.sum(axis=1, keepdims=True) / dphi.dtype.type(n_tokens)
    return d


# ------------------------------------------------------------------------------------------------
# a12: Shapley loss
# ------------------------------------------------------------------------------------------------
def loss_shapley_new(mask: np.ndarray, v_0: np.ndarray, v_s: np.ndarray, phi: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """reference models/shapley.py:9-53.  mask (B,S,n) {0,1}; v_0 (1,C); v_s (B*S,C); phi (B,C,n).
    Returns (loss, dloss/dphi)."""
    B, S, n = mask.shape
    C = phi.shape[1]
    dt = phi.dtype
    m = mask.astype(dt)
    approx = v_0.reshape(1, 1, -1).astype(dt) + m @ phi.transpose(0, 2, 1)        # l.40-44
    resid = approx.reshape(B * S, C) - v_s.astype(dt)
    loss = dt.type(n) * np.mean(resid * resid, dtype=dt)                           # l.48-50
    dapprox = (dt.type(2.0 * n / (B * S * C)) * resid).reshape(B, S, C)
    dphi = np.einsum("bsc,bsn->bcn", dapprox, m).astype(dt)
    return loss, dphi


# ------------------------------------------------------------------------------------------------
# 8f-1: rank masks for the faithfulness / masked-accuracy evaluators
# ------------------------------------------------------------------------------------------------
def descending_ranking(a: np.ndarray) -> np.ndarray:
    """np.argsort(a)[::-1] as in reference scripts/measure_faithfulness.py:239, with the sort made stable so ties are
    defined (equal scores: the larger index ranks first)."""
    return np.argsort(np.asarray(a).reshape(-1), kind="stable")[::-1]


def perturbed_samples(explanations: np.ndarray, n_players: int, steps: int, mask_base: int) -> Tuple[np.ndarray, np.ndarray]:
    """reference scripts/measure_faithfulness.py:225-251 (_get_perturbed_samples): stops = linspace(0, n, steps) as
    int64; mask_i = mask_base everywhere, XOR 1 on the stops[i] top-ranked players.  -> (stops (steps,), masks (steps, n))"""
    steps = min(n_players, steps)
    ranking = descending_ranking(explanations)
    stops = np.linspace(0, n_players, steps, dtype=np.int64)
    masks = np.ones((steps, n_players), dtype=np.int64) * mask_base
    for r, i in enumerate(stops):
        masks[r, ranking[:i]] ^= 1
    return stops, masks


def selective_masks_from_keys(keys: np.ndarray, n_masked: int) -> np.ndarray:
    """Fixed-count masks (reference models/shapley.py:118-128 draws the masked subset with random.shuffle): here the
    subset is the n_masked top-ranked random keys, which is the same uniform distribution over subsets.
    keys (rows, n) float32 -> (rows, n) int64 with exactly n_masked zeros per row."""
    keys = np.asarray(keys, dtype=np.float32)
    out = np.ones(keys.shape, dtype=np.int64)
    for r in range(keys.shape[0]):
        out[r, descending_ranking(keys[r])[:n_masked]] = 0
    return out
