"""KernelSHAP weighted-least-squares restatement (test infrastructure; float64 numpy).

PARITY UNPINNED: in the reference this arithmetic lives entirely inside the third-party package
`shap ~= 0.44.1` (reference requirements.txt:10; call sites models/kernel_shap_bert.py:170-181 and
scripts/train_kernel_shap_explainer.py:50), which is neither vendored under /root/reference nor installed
here, and no reference test pins a value at that boundary.  What follows restates the published KernelSHAP
estimator (Lundberg & Lee 2017) the way `shap.KernelExplainer.solve` implements it — eliminate one feature
with the efficiency constraint, solve the remaining weighted least squares — WITHOUT shap's optional
`l1_reg="auto"` LassoLarsIC feature pre-selection; coalitions and kernel weights are inputs.

OMITTED ON PURPOSE, stated loudly: shap's default l1_reg="auto" runs LassoLarsIC(criterion="aic") on the
sqrt-weighted, constraint-eliminated design whenever the sampled coalitions cover less than 20 % of the 2^M subset
space (always for M >= 14 at the reference's n_samples = 512) and keeps only the features with a non-zero Lasso
coefficient before the final least squares.  Neither this restatement nor the CUDA path applies that pre-selection:
both solve the un-regularised weighted least squares over all varying features (`explain_varying` below restates
shap's restriction to the features that differ between the explained row and the background, and the minimum-norm
solution numpy.linalg.lstsq returns when the system is under-determined).
"""
from __future__ import annotations

from math import comb
from typing import Tuple

import numpy as np


def logit(p: np.ndarray) -> np.ndarray:
    """shap link="logit" (reference models/kernel_shap_bert.py:173)"""
    p = np.asarray(p, dtype=np.float64)
    return np.log(p / (1.0 - p))


def shapley_kernel_weight(d: int, k: int) -> float:
    """pi(k) = (d-1) / (C(d,k) k (d-k)) for 0 < k < d"""
    return (d - 1.0) / (comb(d, k) * k * (d - k))


def sample_coalitions(d: int, n_samples: int, seed: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """A KernelSHAP-style coalition set (an INPUT generator for tests/bench, not a claim about shap's RNG):
    all subsets of size 1 and d-1 with their exact kernel weights (2d rows), then paired random subsets whose
    sizes follow the Shapley-kernel size distribution, the sampled rows sharing the remaining weight equally.
    Returns Z (S, d) uint8 and w (S,) float64 with sum(w) = 1."""
    rng = np.random.default_rng(seed)
    rows, weights = [], []
    size_w = np.array([(d - 1.0) / (k * (d - k)) for k in range(1, d)])
    size_w /= size_w.sum()                       # mass of each subset size 1..d-1
    enumerated_mass = 0.0
    if n_samples >= 2 * d + 2:
        for k in (1, d - 1):
            for j in range(d):
                z = np.zeros(d, np.uint8) if k == 1 else np.ones(d, np.uint8)
                z[j] = 1 if k == 1 else 0
                rows.append(z)
                weights.append(size_w[k - 1] / d)
            enumerated_mass += size_w[k - 1]
        sizes = np.arange(2, d - 1)
    else:
        sizes = np.arange(1, d)
    n_left = n_samples - len(rows)
    p = size_w[sizes - 1] / size_w[sizes - 1].sum()
    n_pairs = n_left // 2
    for _ in range(n_pairs):
        k = int(rng.choice(sizes, p=p))
        z = np.zeros(d, np.uint8)
        z[rng.permutation(d)[:k]] = 1
        rows.append(z)
        rows.append(1 - z)
    if n_left % 2:
        k = int(rng.choice(sizes, p=p))
        z = np.zeros(d, np.uint8)
        z[rng.permutation(d)[:k]] = 1
        rows.append(z)
    n_sampled = len(rows) - len(weights)
    weights += [(1.0 - enumerated_mass) / max(n_sampled, 1)] * n_sampled
    return np.stack(rows), np.asarray(weights, dtype=np.float64)


def pack_features(Z: np.ndarray) -> np.ndarray:
    """(S, d) {0,1} -> (S, ceil(d/32)) uint32, bit j%32 of word j//32 = feature j (no CLS offset here:
    KernelSHAP's features are the T token positions themselves, reference models/kernel_shap_bert.py:183-185)."""
    S, d = Z.shape
    W = (d + 31) // 32
    padded = np.zeros((S, W * 32), dtype=np.uint64)
    padded[:, :d] = Z != 0
    return (padded.reshape(S, W, 32) * (np.uint64(1) << np.arange(32, dtype=np.uint64))).sum(axis=2).astype(np.uint32)


def wls_solve(Z: np.ndarray, w: np.ndarray, y: np.ndarray, delta: np.ndarray) -> np.ndarray:
    """Efficiency-constrained weighted least squares.
    Z (S,d) coalitions, w (S,) kernel weights, y (S,C) = link(E_bg f(h_x(z))) - link(f_null),
    delta (C,) = link(f(x)) - link(f_null).  Eliminate the last feature:
        E = Z[:, :-1] - Z[:, -1:],  y~ = y - Z[:, -1:] * delta
        (E^T W E) phi[:-1] = E^T W y~   (Gram + Cholesky),   phi[-1] = delta - sum(phi[:-1])
    Returns phi (C, d)."""
    Z = np.asarray(Z, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    delta = np.asarray(delta, dtype=np.float64).reshape(-1)
    E = Z[:, :-1] - Z[:, -1:]
    yt = y - Z[:, -1:] * delta[None, :]
    A = E.T @ (w[:, None] * E)
    rhs = E.T @ (w[:, None] * yt)
    L = np.linalg.cholesky(A)
    sol = np.linalg.solve(L.T, np.linalg.solve(L, rhs))      # (d-1, C)
    last = delta - sol.sum(axis=0)
    return np.concatenate([sol, last[None, :]], axis=0).T     # (C, d)


def wls_solve_lstsq(Z, w, y, delta) -> np.ndarray:
    """Same estimator through the sqrt-weighted lstsq form (what shap 0.44 calls) — a self-check of wls_solve."""
    Z = np.asarray(Z, dtype=np.float64)
    sw = np.sqrt(np.asarray(w, dtype=np.float64))[:, None]
    delta = np.asarray(delta, dtype=np.float64).reshape(-1)
    E = Z[:, :-1] - Z[:, -1:]
    yt = np.asarray(y, dtype=np.float64) - Z[:, -1:] * delta[None, :]
    sol = np.linalg.lstsq(sw * E, sw * yt, rcond=None)[0]
    return np.concatenate([sol, (delta - sol.sum(axis=0))[None, :]], axis=0).T


def explain(probs_coalitions: np.ndarray, prob_full: np.ndarray, prob_null: np.ndarray, Z: np.ndarray, w: np.ndarray
            ) -> np.ndarray:
    """probs (S,C) = background-averaged model outputs per coalition; logit link as in the reference call."""
    y = logit(probs_coalitions) - logit(prob_null)[None, :]
    delta = logit(prob_full) - logit(prob_null)
    return wls_solve(Z, w, y, delta)


def varying_features(x: np.ndarray, background: np.ndarray) -> np.ndarray:
    """Indices of the features of x that differ from at least one background row (shap.KernelExplainer.varying_groups)."""
    return np.nonzero((np.asarray(background) != np.asarray(x)[None, :]).any(axis=0))[0]


def explain_varying(model, x: np.ndarray, background: np.ndarray, Zm: np.ndarray, w: np.ndarray) -> np.ndarray:
    """KernelSHAP for one row the way shap.KernelExplainer.explain organises it (logit link): regression over the M
    varying features only, the rest get 0; M = 0 -> zeros, M = 1 -> that feature gets link(f(x)) - link(E f).
    model: (n, T) int64 ids -> (n, C) probabilities; Zm (S, M) coalitions over the varying features, w (S,) weights.
    Under-determined systems take the minimum-norm least-squares solution (numpy lstsq).  -> phi (C, T)"""
    x = np.asarray(x)
    background = np.asarray(background)
    K, T = background.shape
    f_null = model(background).astype(np.float64).mean(axis=0)
    f_x = model(x[None, :]).astype(np.float64)[0]
    C = f_x.shape[0]
    idx = varying_features(x, background)
    M = idx.size
    phi = np.zeros((C, T), dtype=np.float64)
    if M == 0:
        return phi
    if M == 1:
        phi[:, idx[0]] = logit(f_x) - logit(f_null)
        return phi
    S = Zm.shape[0]
    assert Zm.shape[1] == M
    Z = np.zeros((S, T), dtype=bool)
    Z[:, idx] = Zm != 0
    synth = np.where(Z[:, None, :], x[None, None, :], background[None, :, :]).reshape(S * K, T)
    probs = model(synth).astype(np.float64).reshape(S, K, C).mean(axis=1)
    y = logit(probs) - logit(f_null)[None, :]
    delta = logit(f_x) - logit(f_null)
    E = Zm[:, :-1].astype(np.float64) - Zm[:, -1:].astype(np.float64)
    if np.linalg.matrix_rank(E) < M - 1:
        phi[:, idx] = wls_solve_lstsq(Zm, w, y, delta)
    else:
        phi[:, idx] = wls_solve(Zm, w, y, delta)
    return phi
