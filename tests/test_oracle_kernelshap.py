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
