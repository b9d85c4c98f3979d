"""Froyo (frozen backbone) variants on the GPU vs goldens from the reference's own classes (SURVEY.md 8f-4;
reference models/froyo_{vit,bert}.py, recipes/froyo_{vit,bert}.py).  Goldens: tests/golden/make_golden.py froyo."""
import os

import numpy as np
import pytest
import torch

from oracle import configs as ocfg
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _np(t):
    return t.detach().float().cpu().numpy()


def _recipe(name):
    if name.startswith("vit"):
        from autognothi_b200.recipes.froyo_vit import froyo_vit_recipe
        return froyo_vit_recipe()
    from autognothi_b200.recipes.froyo_bert import froyo_bert_recipe
    return froyo_bert_recipe()


def _state(sd):
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}


@pytest.mark.parametrize("name,precision", [("vit_mini", "fp32"), ("bert_mini", "fp32"), ("vit_tiny", "fp32"),
                                            ("vit_tiny", "bf16"), ("bert_mini", "bf16")])
def test_froyo_final_vs_reference_golden(agb, golden_dir, name, precision):
    g = np.load(os.path.join(golden_dir, f"froyo_{name}.npz"))
    rec = _recipe(name)
    cfgd = ocfg.get_config(name)
    cfg = rec.t_config(**cfgd)
    n = rec.n_players(cfg)
    final = rec.t_final(cfg)
    final.load_state_dict(_state(synth.froyo_final_state(cfgd, seed=1)), strict=True)
    final = final.to(DEV).eval()
    final.agb_precision = precision
    B = g["ones_cls"].shape[0]
    xs = torch.from_numpy(synth.inputs(cfgd, B, seed=0)).to(DEV)
    with torch.no_grad():
        cls, phi = rec.fw_final(final, xs)
        tok = torch.cat([torch.ones((B, 1), dtype=torch.int64), torch.from_numpy(g["masks"].astype(np.int64))], dim=1).to(DEV)
        cls_m, phi_m = final(xs, tok)
    for got_c, got_p, tag in ((cls, phi, "ones"), (cls_m, phi_m, "masked")):
        rc, rp = g[f"{tag}_cls"], g[f"{tag}_phi"]
        assert got_p.shape == (B, cfgd["num_labels"], n)
        if precision == "fp32":
            np.testing.assert_allclose(_np(got_c), rc, rtol=1e-4, atol=2e-6)
            np.testing.assert_allclose(_np(got_p), rp, rtol=1e-4, atol=1e-4 * np.abs(rp).max())
        else:
            np.testing.assert_allclose(_np(got_c), rc, atol=2e-2)
            a, b = _np(got_p).reshape(-1).astype(np.float64), rp.reshape(-1).astype(np.float64)
            r = float(np.corrcoef(a, b)[0, 1])
            l2 = float(np.linalg.norm(a - b) / np.linalg.norm(b))
            assert r >= 0.999 and l2 <= 1e-2, (r, l2)


@pytest.mark.parametrize("name", ["vit_mini", "bert_mini"])
def test_froyo_conversion_chain_and_final_coherency(agb, name):
    """classifier -> surrogate -> explainer -> final through the recipe's conv_* rules, then the reference's
    `_verify_final_coherency` (scripts/train_all.py:199-215, eps 1e-5): the bundle equals the separate calls."""
    rec = _recipe(name)
    cfgd = ocfg.get_config(name)
    cfg = rec.t_config(**cfgd)
    n = rec.n_players(cfg)
    from autognothi_b200.recipes.froyo_bert import FroyoBertMisc
    misc = FroyoBertMisc(tokenizer=None) if not name.startswith("vit") else None
    classifier = rec.t_classifier(cfg)
    classifier.load_state_dict(_state(synth.surrogate_state(cfgd, seed=3)), strict=True)
    classifier = classifier.to(DEV).eval()
    surrogate = rec.conv_classifier_surrogate(cfg, misc, classifier)
    with torch.no_grad():      # give the surrogate its own head so that classifier != surrogate
        surrogate.classifier.weight.mul_(0.5).add_(0.01)
    explainer = rec.conv_surrogate_explainer(cfg, misc, surrogate)
    for m in (classifier, surrogate, explainer):
        m.agb_precision = "fp32"
    final = rec.conv_explainer_final(cfg, misc, classifier, surrogate, explainer).eval()
    final.agb_precision = "fp32"
    B = 2
    xs = torch.from_numpy(synth.inputs(cfgd, B, seed=4)).to(DEV)
    ones = torch.ones((B, n), dtype=torch.int64, device=DEV)
    with torch.no_grad():
        logits, phi = rec.fw_final(final, xs)
        ys, _ = rec.fw_classifier(classifier, xs, ones)
        grand, _ = rec.fw_surrogate(surrogate, xs, ones)
        null, _ = rec.fw_surrogate(surrogate, rec.gen_null(cfg, misc, DEV), torch.ones((1, n), dtype=torch.int64, device=DEV))
        phi_sep, _ = rec.fw_explainer(explainer, xs, ones, grand, null)
    np.testing.assert_allclose(_np(final.surrogate_null), _np(null), atol=1e-6)
    np.testing.assert_allclose(_np(logits), _np(ys), atol=1e-5)
    np.testing.assert_allclose(_np(phi), _np(phi_sep), atol=1e-5 * max(1.0, float(phi_sep.abs().max())))


@pytest.mark.parametrize("name,precision", [("vit_mini", "fp32"), ("bert_mini", "fp32"), ("vit_mini", "bf16")])
def test_froyo_explainer_training_touches_only_the_heads(agb, golden_dir, name, precision):
    """Frozen backbone: the explainer_* gradients equal the reference's autograd (the vanilla goldens hold them: freezing
    the backbone does not change the gradient of the remaining parameters) and no `vit.` / `bert.` gradient exists."""
    from autognothi_b200.models import shapley as ash
    g = np.load(os.path.join(golden_dir, f"model_{name}.npz"))
    t = np.load(os.path.join(golden_dir, f"train_{name}.npz"))
    B, S, n = (int(v) for v in g["meta"])
    rec = _recipe(name)
    cfgd = ocfg.get_config(name)
    cfg = rec.t_config(**cfgd)
    exp = rec.t_explainer(cfg)
    exp.load_state_dict(_state(synth.explainer_state(cfgd, seed=1)), strict=True)
    exp = exp.to(DEV).train()
    exp.agb_precision = precision
    exp.agb_dropout = False      # goldens: reference in eval() mode
    xs = torch.from_numpy(synth.inputs(cfgd, B, seed=0)).to(DEV)
    masks = torch.from_numpy(g["masks"].astype(np.int64)).to(DEV).reshape(B, S, n)
    v_s, grand, null = (torch.from_numpy(g[k]).to(DEV) for k in ("v_s", "grand", "null"))
    ones = torch.ones((B, n), dtype=torch.int64, device=DEV)
    phi, _ = rec.fw_explainer(exp, xs, ones, grand, null)
    loss = ash.loss_shapley_new(B, S, n, masks, null, v_s, grand, phi)
    loss.backward()
    tol = 1e-4 if precision == "fp32" else 2e-2
    assert abs(float(loss.detach()) - float(t["loss"])) <= tol * abs(float(t["loss"]))
    ref_norms = dict(zip([str(s) for s in t["norm_names"]], t["norm_values"]))
    floor = 1e-5 * max(ref_norms.values())
    for k, p in exp.named_parameters():
        if k.startswith(("vit.", "bert.")):
            assert not p.requires_grad and p.grad is None, k
            continue
        assert p.grad is not None, k
        got = float(p.grad.norm())
        if precision == "fp32":
            assert abs(got - ref_norms[k]) <= 2e-3 * ref_norms[k] + floor, f"{k}: |grad| {got} vs {ref_norms[k]}"
        elif ref_norms[k] > 100 * floor:
            assert abs(got - ref_norms[k]) <= 5e-2 * ref_norms[k], f"{k}: |grad| {got} vs {ref_norms[k]}"


@pytest.mark.parametrize("name", ["vit_mini", "bert_mini"])
def test_froyo_surrogate_training_touches_only_the_head(agb, golden_dir, name):
    """Froyo surrogate (reference models/froyo_vit.py:76-85): backbone frozen, head gradients from torch autograd over the
    engine's CLS rows; they equal the vanilla surrogate-training goldens for the same parameters."""
    from autognothi_b200.models import shapley as ash
    t = np.load(os.path.join(golden_dir, f"train_surrogate_{name}.npz"))
    rec = _recipe(name)
    cfgd = ocfg.get_config(name)
    cfg = rec.t_config(**cfgd)
    n = rec.n_players(cfg)
    srg = rec.t_surrogate(cfg)
    srg.load_state_dict(_state(synth.surrogate_state(cfgd, seed=0)), strict=True)
    srg = srg.to(DEV).train()
    srg.agb_precision = "fp32"
    srg.agb_dropout = False      # goldens: reference in eval() mode
    B = t["masks"].shape[0]
    xs = torch.from_numpy(synth.inputs(cfgd, B, seed=0)).to(DEV)
    masks = torch.from_numpy(t["masks"].astype(np.int64)).to(DEV)
    adapt, _ = rec.fw_surrogate(srg, xs, masks)
    np.testing.assert_allclose(_np(adapt), t["adapt"], rtol=1e-4, atol=2e-6)
    loss = ash.loss_logits_kl_divergence(torch.from_numpy(t["orig"]).to(DEV), adapt)
    loss.backward()
    ref_norms = dict(zip([str(s) for s in t["norm_names"]], t["norm_values"]))
    floor = 1e-5 * max(ref_norms.values())
    seen = 0
    for k, p in srg.named_parameters():
        if k.startswith(("vit.", "bert.")):
            assert p.grad is None, k
            continue
        seen += 1
        got = float(p.grad.norm())
        assert abs(got - ref_norms[k]) <= 2e-3 * ref_norms[k] + floor, f"{k}: |grad| {got} vs {ref_norms[k]}"
    assert seen >= 2
