"""CPU: self-consistency of the KernelSHAP restatement (parity UNPINNED — shap is not available; see
oracle/kernelshap.py).  These pin the estimator's defining properties rather than a third-party output."""
import itertools
from math import comb, factorial

import numpy as np

from oracle import kernelshap as oks


def _exact_shapley(v, d):
    phi = np.zeros(d)
    for j in range(d):
        others = [i for i in range(d) if i != j]
        for r in range(d):
            for S in itertools.combinations(others, r):
                wt = factorial(r) * factorial(d - r - 1) / factorial(d)
                phi[j] += wt * (v(set(S) | {j}) - v(set(S)))
    return phi


def test_full_enumeration_recovers_exact_shapley_values():
    d = 6
    rng = np.random.default_rng(0)
    table = {frozenset(s): rng.standard_normal() for r in range(d + 1) for s in itertools.combinations(range(d), r)}
    v = lambda s: table[frozenset(s)]
    Z, w, y = [], [], []
    for r in range(1, d):
        for s in itertools.combinations(range(d), r):
            z = np.zeros(d); z[list(s)] = 1
            Z.append(z); w.append(oks.shapley_kernel_weight(d, r)); y.append(v(s) - v(()))
    Z, w, y = np.array(Z), np.array(w), np.array(y)[:, None]
    delta = np.array([v(range(d)) - v(())])
    phi = oks.wls_solve(Z, w, y, delta)[0]
    np.testing.assert_allclose(phi, _exact_shapley(v, d), rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(phi.sum(), delta[0], rtol=1e-12)


def test_cholesky_and_lstsq_forms_agree_and_efficiency_holds():
    d, S, C = 33, 400, 2
    Z, w = oks.sample_coalitions(d, S, seed=1)
    assert Z.shape == (S, d) and abs(w.sum() - 1.0) < 1e-12
    rng = np.random.default_rng(2)
    y = rng.standard_normal((S, C))
    delta = rng.standard_normal(C)
    a, b = oks.wls_solve(Z, w, y, delta), oks.wls_solve_lstsq(Z, w, y, delta)
    np.testing.assert_allclose(a, b, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(a.sum(axis=1), delta, rtol=1e-10)


def test_additive_model_is_recovered_exactly():
    d, S = 20, 300
    Z, w = oks.sample_coalitions(d, S, seed=3)
    beta = np.random.default_rng(4).standard_normal(d)
    y = (Z @ beta)[:, None]
    phi = oks.wls_solve(Z, w, y, np.array([beta.sum()]))[0]
    np.testing.assert_allclose(phi, beta, rtol=1e-8, atol=1e-9)


def test_pack_features_layout():
    Z = (np.random.default_rng(5).random((9, 70)) > 0.5).astype(np.uint8)
    P = oks.pack_features(Z)
    assert P.shape == (9, 3) and P.dtype == np.uint32
    for j in range(70):
        np.testing.assert_array_equal((P[:, j // 32] >> (j % 32)) & 1, Z[:, j])


def _toy_model(beta, bias=-0.3):
    """probabilities of a 2-class model whose class-1 logit is additive in 'token present at position j == x_j' features"""
    def model(ids):
        ids = np.asarray(ids)
        logit1 = bias + (np.sin(ids * 0.37) * beta[None, :]).sum(axis=1)
        p1 = 1.0 / (1.0 + np.exp(-logit1))
        return np.stack([1.0 - p1, p1], axis=1)
    return model


def test_explain_varying_restricts_to_varying_features_and_keeps_efficiency():
    """shap.KernelExplainer only regresses over features that differ from the background; the others get 0."""
    T, K = 24, 3
    rng = np.random.default_rng(7)
    background = rng.integers(5, 50, size=(K, T))
    x = background[0].copy()
    vary = np.array([1, 4, 5, 9, 17, 23])
    x[vary] += 100
    background[:, 0] = x[0]                       # CLS-like column: identical everywhere
    model = _toy_model(rng.standard_normal(T))
    idx = oks.varying_features(x, background)
    assert set(vary) <= set(idx)
    Zm, w = oks.sample_coalitions(idx.size, 400, seed=2)
    phi = oks.explain_varying(model, x, background, Zm, w)
    assert phi.shape == (2, T)
    mask = np.zeros(T, bool); mask[idx] = True
    assert np.all(phi[:, ~mask] == 0.0)
    delta = oks.logit(model(x[None])[0]) - oks.logit(model(background).mean(axis=0))
    np.testing.assert_allclose(phi.sum(axis=1), delta, rtol=1e-9, atol=1e-10)
    # M = 1 and M = 0
    x1 = background[0].copy(); x1[3] += 7
    bg1 = np.repeat(background[:1], K, axis=0)
    phi1 = oks.explain_varying(model, x1, bg1, np.zeros((0, 1)), np.zeros(0))
    d1 = oks.logit(model(x1[None])[0]) - oks.logit(model(bg1).mean(axis=0))
    np.testing.assert_allclose(phi1[:, 3], d1, rtol=1e-12)
    assert np.count_nonzero(phi1) == 2
    assert np.count_nonzero(oks.explain_varying(model, bg1[0], bg1, np.zeros((0, 0)), np.zeros(0))) == 0


def test_underdetermined_system_takes_the_minimum_norm_solution():
    """Fewer (paired) coalitions than unknowns — the reference's own hparams (512 positions, 512 samples): the Gram matrix
    is singular; like numpy lstsq inside shap the restatement returns the minimum-norm solution, efficiency still holds."""
    T, K, S = 64, 2, 40
    rng = np.random.default_rng(11)
    background = rng.integers(5, 50, size=(K, T))
    x = background[0] + 100
    model = _toy_model(rng.standard_normal(T) * 0.2)
    Zm, w = oks.sample_coalitions(T, S, seed=5)
    E = Zm[:, :-1].astype(float) - Zm[:, -1:].astype(float)
    assert np.linalg.matrix_rank(E) < T - 1
    phi = oks.explain_varying(model, x, background, Zm, w)
    delta = oks.logit(model(x[None])[0]) - oks.logit(model(background).mean(axis=0))
    np.testing.assert_allclose(phi.sum(axis=1), delta, rtol=1e-9, atol=1e-10)
    assert np.isfinite(phi).all()
