"""The oracle's restatement of the Froyo / LTT / Duo variants (oracle/variants.py) pinned against outputs of the
reference's own classes (tests/golden/make_golden.py froyo | ltt | duo).  CPU only."""
import os

import numpy as np
import pytest

from oracle import configs as ocfg
from oracle import synth
from oracle import variants as ov

TOL = dict(rtol=1e-4, atol=2e-6)


def _attr_close(got, ref):
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-4 * np.abs(ref).max())


def _shapes(golden_dir, fname, name, role):
    import json
    with open(os.path.join(golden_dir, fname)) as f:
        return {k: tuple(v) for k, v in json.load(f)[name][role].items()}


@pytest.mark.parametrize("name", ["vit_mini", "bert_mini"])
def test_froyo_bundle_matches_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"froyo_{name}.npz"))
    cfg = ocfg.get_config(name)
    n = ocfg.n_players(cfg)
    B = g["ones_cls"].shape[0]
    sd = synth.froyo_final_state(cfg, seed=1)
    xs = synth.inputs(cfg, B, seed=0)
    for tag, m in (("ones", np.ones((B, n), np.int64)), ("masked", g["masks"].astype(np.int64))):
        cls, phi = ov.froyo_final(sd, cfg, xs, m)
        np.testing.assert_allclose(cls, g[f"{tag}_cls"], **TOL)
        _attr_close(phi, g[f"{tag}_phi"])


@pytest.mark.parametrize("name", ["ltt_vit_mini", "ltt_bert_mini"])
def test_ltt_matches_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    B, S, n = (int(v) for v in g["meta"])
    cfg = ocfg.get_config(name)
    xs = synth.inputs(cfg, B, seed=0)
    masks = g["masks"].astype(np.int64)
    sds = [synth.state_like(_shapes(golden_dir, "ltt_keys.json", name, role), seed=30 + i)
           for i, role in enumerate(("surrogate", "explainer", "final"))]
    v_side, v_main = ov.ltt_surrogate(sds[0], cfg, np.repeat(xs, S, axis=0), masks)
    np.testing.assert_allclose(v_side, g["v_side"], **TOL)
    np.testing.assert_allclose(v_main, g["v_main"], **TOL)
    phi, e_main = ov.ltt_explainer(sds[1], cfg, xs, np.ones((B, n), np.int64), g["grand"], g["null"])
    _attr_close(phi, g["phi"])
    np.testing.assert_allclose(e_main, g["e_main"], **TOL)
    phi_m, _ = ov.ltt_explainer(sds[1], cfg, xs, masks.reshape(B, S, n)[:, 0, :], g["grand"], g["null"])
    _attr_close(phi_m, g["phi_masked"])
    f_cls, f_phi = ov.ltt_final(sds[2], cfg, xs, np.ones((B, n), np.int64))
    np.testing.assert_allclose(f_cls, g["f_cls"], **TOL)
    _attr_close(f_phi, g["f_phi"])


@pytest.mark.parametrize("name", ["vit_mini", "bert_mini"])
def test_duo_explainer_matches_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"duo_{name}.npz"))
    gm = np.load(os.path.join(golden_dir, f"model_{name}.npz"))
    B, S, n = (int(v) for v in g["meta"])
    cfg = ocfg.get_config(name)
    sd = synth.state_like(_shapes(golden_dir, "duo_keys.json", name, "explainer"), seed=50)
    xs = synth.inputs(cfg, B, seed=0)
    phi, cls = ov.duo_explainer(sd, cfg, xs, np.ones((B, n), np.int64), gm["grand"], gm["null"])
    _attr_close(phi, g["phi"])
    np.testing.assert_allclose(cls, g["cls"], rtol=1e-4, atol=1e-5)
    phi_m, cls_m = ov.duo_explainer(sd, cfg, xs, gm["masks"].astype(np.int64).reshape(B, S, n)[:, 0, :], gm["grand"], gm["null"])
    _attr_close(phi_m, g["phi_masked"])
    np.testing.assert_allclose(cls_m, g["cls_masked"], rtol=1e-4, atol=1e-5)
