"""Host-side checks of the Froyo mirror (reference models/froyo_{vit,bert}.py, recipes/froyo_{vit,bert}.py) that need no
GPU: state-dict ABI of every class and which parameters train() leaves trainable, against tables dumped from the
reference's own classes (tests/golden/make_golden.py froyo -> froyo_keys.json)."""
import json
import os

import pytest

from oracle import configs as ocfg


@pytest.fixture(scope="module")
def froyo_keys(golden_dir):
    with open(os.path.join(golden_dir, "froyo_keys.json")) as f:
        return json.load(f)


def _recipe(name):
    if name.startswith("vit"):
        from autognothi_b200.recipes.froyo_vit import froyo_vit_recipe
        return froyo_vit_recipe()
    from autognothi_b200.recipes.froyo_bert import froyo_bert_recipe
    return froyo_bert_recipe()


@pytest.mark.parametrize("name", ["vit_mini", "bert_mini"])
def test_froyo_state_dict_abi_and_freezes(froyo_keys, name):
    rec = _recipe(name)
    cfg = rec.t_config(**ocfg.get_config(name))
    ref = froyo_keys[name]
    for role, cls in (("classifier", rec.t_classifier), ("surrogate", rec.t_surrogate), ("explainer", rec.t_explainer)):
        m = cls(cfg).train()
        assert {k: list(v.shape) for k, v in m.state_dict().items()} == ref[role], role
        assert sorted(k for k, p in m.named_parameters() if p.requires_grad) == ref[role + "_trainable"], role
    final = rec.t_final(cfg)
    assert {k: list(v.shape) for k, v in final.state_dict().items()} == ref["final"]
    final.train()
    assert not any(p.requires_grad for k, p in final.named_parameters()
                   if k.startswith(("vit.", "bert.", "bert_pooler.", "classifier.")) or k == "surrogate_null")
    assert rec.id == ("froyo_vit" if name.startswith("vit") else "froyo_bert") and rec.version == "beta.1.01"
    assert rec.n_players(cfg) == ocfg.n_players(ocfg.get_config(name))


def test_froyo_models_refuse_to_run_on_the_cpu():
    import torch
    rec = _recipe("vit_mini")
    cfgd = ocfg.get_config("vit_mini")
    cfg = rec.t_config(**cfgd)
    final = rec.t_final(cfg).eval()
    with pytest.raises((RuntimeError, AssertionError)):
        final(torch.zeros(1, 3, cfgd["img_px_size"], cfgd["img_px_size"]), torch.ones(1, ocfg.n_players(cfgd) + 1, dtype=torch.int64))
