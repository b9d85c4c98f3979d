"""Duo (dual-objective) variants on the GPU vs goldens from the reference's own classes (reference
models/duo_vanilla_{vit,bert}.py, scripts/train_duo_explainer.py:180-196).  Goldens: tests/golden/make_golden.py duo."""
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


def _setup(name, precision):
    if name.startswith("vit"):
        from autognothi_b200.recipes.duo_vanilla_vit import duo_vanilla_vit_recipe as r
    else:
        from autognothi_b200.recipes.duo_vanilla_bert import duo_vanilla_bert_recipe as r
    rec = r()
    cfgd = ocfg.get_config(name)
    cfg = rec.t_config(**cfgd)
    out = []
    for i, cls in enumerate((rec.t_explainer, rec.t_final)):       # seeds 50, 51 as in make_golden.py
        m = cls(cfg)
        sd = synth.state_like({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=50 + i)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
        m = m.to(DEV).eval()
        for sub in m.modules():
            if hasattr(sub, "agb_precision"):
                sub.agb_precision = precision
        out.append(m)
    return rec, cfgd, cfg, out


@pytest.mark.parametrize("name,precision", [("vit_mini", "fp32"), ("bert_mini", "fp32"), ("vit_mini", "bf16"), ("bert_mini", "bf16")])
def test_duo_forward_vs_reference_golden(agb, golden_dir, name, precision):
    g = np.load(os.path.join(golden_dir, f"duo_{name}.npz"))
    gm = np.load(os.path.join(golden_dir, f"model_{name}.npz"))
    B, S, n = (int(v) for v in g["meta"])
    rec, cfgd, cfg, (exp, fin) = _setup(name, precision)
    xs = torch.from_numpy(synth.inputs(cfgd, B, seed=0)).to(DEV)
    masks = torch.from_numpy(gm["masks"].astype(np.int64)).to(DEV).reshape(B, S, n)
    grand, null = torch.from_numpy(gm["grand"]).to(DEV), torch.from_numpy(gm["null"]).to(DEV)
    ones = torch.ones((B, n), dtype=torch.int64, device=DEV)
    with torch.no_grad():
        phi, cls = rec.fw_explainer(exp, xs, ones, grand, null)
        phi_m, cls_m = rec.fw_explainer(exp, xs, masks[:, 0, :].contiguous(), grand, null)
        f_cls, f_phi = rec.fw_final(fin, xs)
    if precision == "fp32":
        tol = dict(rtol=1e-4, atol=2e-6)
        for got, ref in ((phi, g["phi"]), (phi_m, g["phi_masked"]), (f_phi, g["f_phi"])):
            np.testing.assert_allclose(_np(got), ref, rtol=1e-4, atol=1e-4 * np.abs(ref).max())
    else:
        tol = dict(atol=3e-2)
        for got, ref in ((phi, g["phi"]), (phi_m, g["phi_masked"]), (f_phi, g["f_phi"])):
            a, b = _np(got).reshape(-1).astype(np.float64), ref.reshape(-1).astype(np.float64)
            assert float(np.corrcoef(a, b)[0, 1]) >= 0.999 and float(np.linalg.norm(a - b) / np.linalg.norm(b)) <= 1e-2
    np.testing.assert_allclose(_np(cls), g["cls"], **tol)
    np.testing.assert_allclose(_np(cls_m), g["cls_masked"], **tol)
    np.testing.assert_allclose(_np(f_cls), g["f_cls"], **tol)


@pytest.mark.parametrize("name,precision", [("vit_mini", "fp32"), ("bert_mini", "fp32"), ("vit_mini", "bf16")])
def test_duo_dual_objective_gradients(agb, golden_dir, name, precision):
    """loss = cross_entropy(class output, labels) + Shapley loss, as scripts/train_duo_explainer.py:184-196; the class
    head's gradient re-enters the hand-written adjoint at the CLS rows of the backbone output."""
    from autognothi_b200.models import shapley as ash
    g = np.load(os.path.join(golden_dir, f"duo_{name}.npz"))
    gm = np.load(os.path.join(golden_dir, f"model_{name}.npz"))
    B, S, n = (int(v) for v in g["meta"])
    rec, cfgd, cfg, (exp, fin) = _setup(name, precision)
    exp.train()
    exp.agb_dropout = False      # goldens: reference in eval() mode
    xs = torch.from_numpy(synth.inputs(cfgd, B, seed=0)).to(DEV)
    masks = torch.from_numpy(gm["masks"].astype(np.int64)).to(DEV).reshape(B, S, n)
    v_s, grand, null = (torch.from_numpy(gm[k]).to(DEV) for k in ("v_s", "grand", "null"))
    labels = torch.from_numpy(g["labels"]).to(DEV)
    ones = torch.ones((B, n), dtype=torch.int64, device=DEV)
    phi, cls = rec.fw_explainer(exp, xs, ones, grand, null)
    assert phi.requires_grad and cls.requires_grad
    loss_cls = torch.nn.functional.cross_entropy(cls, labels)
    loss_shap = ash.loss_shapley_new(B, S, n, masks, null, v_s, grand, phi)
    (loss_cls + loss_shap).backward()
    t = 1e-4 if precision == "fp32" else 3e-2
    assert abs(float(loss_cls.detach()) - float(g["loss_cls"])) <= t * abs(float(g["loss_cls"]))
    assert abs(float(loss_shap.detach()) - float(g["loss_shap"])) <= t * abs(float(g["loss_shap"]))
    ref_norms = dict(zip([str(s) for s in g["norm_names"]], g["norm_values"]))
    floor = 1e-5 * max(ref_norms.values())
    params = dict(exp.named_parameters())
    assert set(params) == set(ref_norms)
    for k, p in params.items():
        assert p.grad is not None, k
        if precision == "fp32":
            got = float(p.grad.norm())
            assert abs(got - ref_norms[k]) <= 2e-3 * ref_norms[k] + floor, f"{k}: |grad| {got} vs {ref_norms[k]}"
    for key in g.files:
        if not key.startswith("grad::"):
            continue
        k = key[len("grad::"):]
        ref, got = g[key].reshape(-1).astype(np.float64), _np(params[k].grad).reshape(-1).astype(np.float64)
        if precision == "fp32":
            np.testing.assert_allclose(got, ref, rtol=2e-3, atol=2e-3 * np.abs(ref).max() + floor, err_msg=k)
        elif np.linalg.norm(ref) > 100 * floor:
            cos = float(ref @ got / (np.linalg.norm(ref) * np.linalg.norm(got) + 1e-30))
            assert cos > 0.99, f"{k}: cosine {cos}"
