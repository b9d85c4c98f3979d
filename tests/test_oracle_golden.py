"""CPU: the numpy oracle reproduces the reference's own outputs (tests/golden/*.npz, produced by
tests/golden/make_golden.py from the unmodified reference).  This is what pins the oracle."""
import json
import os

import numpy as np
import pytest

from oracle import configs as ocfg
from oracle import shapley as osh
from oracle import synth
from oracle import transformer as otr


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("n", [196, 127, 511, 16, 2])
def test_sampler_bit_exact(golden_dir, n):
    g = _load(golden_dir, "sampler.npz")
    masks = osh.masks_from_uniforms(g[f"n{n}_u_players"], g[f"n{n}_u_size"], g[f"n{n}_prefix"], n)
    assert masks.dtype == np.int64
    np.testing.assert_array_equal(masks, g[f"n{n}_masks"].astype(np.int64))
    # pairs are complements (reference models/shapley.py:75-78)
    np.testing.assert_array_equal(masks[0::2] + masks[1::2], np.ones_like(masks[0::2]))


@pytest.mark.parametrize("n", [196, 127, 511, 16, 2])
def test_prefix_table_matches_reference(golden_dir, n):
    g = _load(golden_dir, "sampler.npz")
    ref = g[f"n{n}_prefix"]
    mine = osh.shapley_size_prefix(n)
    # float32 reduction order differs between numpy and torch: allow 2 ulp, never more
    assert np.max(np.abs(mine - ref)) <= 2 * np.spacing(np.float32(1.0))


def test_purely_uniform_bit_exact(golden_dir):
    g = _load(golden_dir, "sampler.npz")
    m = osh.mask_purely_uniform_from_uniforms(g["pu_u_players"], g["pu_u_row"])
    np.testing.assert_array_equal(m, g["pu_masks"].astype(np.int64))


def test_pack_roundtrip_and_layout():
    rng = np.random.default_rng(0)
    for n in (1, 2, 30, 31, 32, 63, 127, 196, 511):
        m = (rng.random((7, n)) > 0.5).astype(np.int64)
        packed = osh.pack_player_mask(m)
        assert packed.shape == (7, (n + 1 + 31) // 32) and packed.dtype == np.uint32
        assert np.all(packed[:, 0] & 1 == 1)  # bit 0 = CLS, always kept
        back = osh.unpack_token_mask(packed, n + 1)
        np.testing.assert_array_equal(back[:, 1:], m)
        for r in range(3):  # explicit bit arithmetic, independent of the vectorised packer
            for j in range(n):
                assert (int(packed[r, (j + 1) // 32]) >> ((j + 1) % 32)) & 1 == m[r, j]


@pytest.mark.parametrize("tag", ["vit", "bert", "tiny"])
def test_normalize_and_loss(golden_dir, tag):
    g = _load(golden_dir, "shapley_math.npz")
    pred, grand, null = g[f"{tag}_pred"], g[f"{tag}_grand"], g[f"{tag}_null"]
    norm = osh.normalize_shapley_explanation(pred, grand, null)
    np.testing.assert_allclose(norm, g[f"{tag}_norm"], rtol=1e-5, atol=1e-6)
    phi = osh.explainer_output(pred, grand, null)
    np.testing.assert_allclose(phi, g[f"{tag}_phi"], rtol=1e-5, atol=1e-6)
    loss, dphi = osh.loss_shapley_new(g[f"{tag}_mask"].astype(np.int64), null, g[f"{tag}_v_s"], g[f"{tag}_phi"])
    np.testing.assert_allclose(loss, g[f"{tag}_loss"], rtol=1e-5)
    np.testing.assert_allclose(dphi, g[f"{tag}_dphi"], rtol=1e-4, atol=1e-6)
    # the CLS row keeps its share: phi does NOT sum to grand - null
    T = pred.shape[1]
    gap = (grand - null) - phi.sum(axis=2)
    np.testing.assert_allclose(gap, norm[:, 0, :], rtol=1e-4, atol=1e-5)
    assert T == phi.shape[2] + 1


def test_explainer_output_grad_is_adjoint():
    rng = np.random.default_rng(3)
    B, T, C = 2, 9, 3
    pred = rng.standard_normal((B, T, C))
    grand, null = rng.random((B, C)), rng.random((1, C))
    dphi = rng.standard_normal((B, C, T - 1))
    d = osh.explainer_output_grad(dphi, T)
    eps = 1e-6
    num = np.zeros_like(pred)
    for idx in np.ndindex(*pred.shape):
        p2 = pred.copy(); p2[idx] += eps
        num[idx] = ((osh.explainer_output(p2, grand, null) - osh.explainer_output(pred, grand, null)) * dphi).sum() / eps
    np.testing.assert_allclose(d, num, rtol=1e-4, atol=1e-6)


def test_state_dict_keys_match_reference(golden_dir):
    with open(os.path.join(golden_dir, "state_dict_keys.json")) as f:
        ref = json.load(f)
    for name, kinds in ref.items():
        cfg = ocfg.get_config(name)
        mine_s = {k: list(s) for k, s in synth.surrogate_shapes(cfg)}
        mine_e = {k: list(s) for k, s in synth.explainer_shapes(cfg)}
        assert mine_s == kinds["surrogate"], name
        assert mine_e == kinds["explainer"], name


MODEL_CASES = ["vit_mini", "vit_mini_px64", "vit_tiny", "bert_mini", "vit_base", "bert_base_128", "bert_mini_512"]


@pytest.mark.parametrize("name", MODEL_CASES)
def test_model_forward_matches_reference(golden_dir, name):
    g = _load(golden_dir, f"model_{name}.npz")
    B, S, n = (int(v) for v in g["meta"])
    cfg = ocfg.get_config(name)
    assert ocfg.n_players(cfg) == n
    srg = synth.surrogate_state(cfg, seed=0)
    exp = synth.explainer_state(cfg, seed=1)
    xs = synth.inputs(cfg, B, seed=0)
    masks = g["masks"].astype(np.int64)
    xs_ext = np.repeat(xs, S, axis=0)
    v_s = otr.fw_surrogate(srg, cfg, xs_ext, masks)
    np.testing.assert_allclose(v_s, g["v_s"], rtol=1e-4, atol=2e-6)
    ones = np.ones((B, n), dtype=np.int64)
    grand = otr.fw_surrogate(srg, cfg, xs, ones)
    np.testing.assert_allclose(grand, g["grand"], rtol=1e-4, atol=2e-6)
    null = otr.fw_surrogate(srg, cfg, otr.null_input(cfg), np.ones((1, n), dtype=np.int64))
    np.testing.assert_allclose(null, g["null"], rtol=1e-4, atol=2e-6)
    phi, _ = otr.fw_explainer(exp, cfg, xs, ones, g["grand"], g["null"])
    scale = np.abs(g["phi"]).max()
    np.testing.assert_allclose(phi, g["phi"], rtol=1e-4, atol=1e-4 * scale)
    phi_m, _ = otr.fw_explainer(exp, cfg, xs, masks.reshape(B, S, n)[:, 0, :], g["grand"], g["null"])
    np.testing.assert_allclose(phi_m, g["phi_masked"], rtol=1e-4, atol=1e-4 * scale)
    loss, _ = osh.loss_shapley_new(masks.reshape(B, S, n), g["null"], g["v_s"], g["phi"])
    np.testing.assert_allclose(loss, g["loss"], rtol=1e-4)


def test_vit_masked_keys_still_contribute_bert_masked_keys_do_not():
    """ViT mask = logit 0 (masked patches still matter); BERT mask = -inf (dead)."""
    cfg = ocfg.get_config("vit_mini_px64")
    n = ocfg.n_players(cfg)
    sd = synth.surrogate_state(cfg)
    xs = synth.inputs(cfg, 1)
    mask = np.ones((1, n), dtype=np.int64); mask[0, : n // 2] = 0
    a = otr.fw_surrogate(sd, cfg, xs, mask)
    xs2 = xs.copy(); xs2[:, :, :16, :16] += 1.0  # perturb a masked patch (patch 0)
    b = otr.fw_surrogate(sd, cfg, xs2, mask)
    assert np.abs(a - b).max() > 1e-6
    cfgb = ocfg.get_config("bert_mini")
    nb = ocfg.n_players(cfgb)
    sdb = synth.surrogate_state(cfgb)
    ids = synth.inputs(cfgb, 1)
    maskb = np.ones((1, nb), dtype=np.int64); maskb[0, 3:9] = 0
    a = otr.fw_surrogate(sdb, cfgb, ids, maskb)
    ids2 = ids.copy(); ids2[0, 4:10] = 7  # tokens 4..9 = players 3..8, all masked
    b = otr.fw_surrogate(sdb, cfgb, ids2, maskb)
    np.testing.assert_array_equal(a, b)


# ------------------------------------------------------------------------------------------------
# 8f-1: evaluator mask generators (goldens from scripts/measure_faithfulness.py and models/shapley.py)
# ------------------------------------------------------------------------------------------------
def test_perturbed_samples_bit_exact(golden_dir):
    g = _load(golden_dir, "evaluators.npz")
    for idx, (n, steps, base) in enumerate(g["perturb_cases"]):
        stops, masks = osh.perturbed_samples(g[f"perturb_{idx}_attr"], int(n), int(steps), int(base))
        np.testing.assert_array_equal(stops, g[f"perturb_{idx}_stops"])
        np.testing.assert_array_equal(masks, g[f"perturb_{idx}_masks"].astype(np.int64))
        # stop i flips exactly stops[i] players of the base mask
        np.testing.assert_array_equal((masks != base).sum(axis=1), stops)


def test_selective_masks_have_the_reference_shape_and_counts(golden_dir):
    g = _load(golden_dir, "evaluators.npz")
    rng = np.random.RandomState(0)
    for idx, (b, n, k) in enumerate(g["selective_cases"]):
        ref = g[f"selective_{idx}"].astype(np.int64)
        assert ref.shape == (b, n) and ((ref == 0).sum(axis=1) == k).all()       # what the reference guarantees
        mine = osh.selective_masks_from_keys(rng.rand(int(b), int(n)).astype(np.float32), int(k))
        assert mine.shape == ref.shape and mine.dtype == np.int64
        assert ((mine == 0).sum(axis=1) == k).all() and set(np.unique(mine)) <= {0, 1}
