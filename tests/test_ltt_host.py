"""Host-side checks of the LTT mirror (reference models/ltt_{vit,bert}.py, recipes/ltt_{vit,bert}.py) that need no GPU:
state-dict ABI of every class and which parameters train() leaves trainable, against tables dumped from the reference's
own classes (tests/golden/make_golden.py ltt -> ltt_keys.json); the bundle conversion's ladder re-indexing."""
import json
import os

import pytest
import torch

from oracle import configs as ocfg


@pytest.fixture(scope="module")
def ltt_keys(golden_dir):
    with open(os.path.join(golden_dir, "ltt_keys.json")) as f:
        return json.load(f)


def _recipe(name):
    if "vit" in name:
        from autognothi_b200.recipes.ltt_vit import ltt_vit_recipe
        return ltt_vit_recipe()
    from autognothi_b200.recipes.ltt_bert import ltt_bert_recipe
    return ltt_bert_recipe()


@pytest.mark.parametrize("name", ["ltt_vit_mini", "ltt_bert_mini"])
def test_ltt_state_dict_abi_and_freezes(ltt_keys, name):
    rec = _recipe(name)
    cfg = rec.t_config(**ocfg.get_config(name))
    ref = ltt_keys[name]
    assert rec.t_classifier is rec.t_surrogate          # sic, as in the reference
    for role, cls in (("surrogate", rec.t_surrogate), ("explainer", rec.t_explainer), ("final", rec.t_final)):
        m = cls(cfg).train()
        assert {k: list(v.shape) for k, v in m.state_dict().items()} == ref[role], role
        assert sorted(k for k, p in m.named_parameters() if p.requires_grad) == ref[role + "_trainable"], role
    assert rec.id == ("ltt_vit" if "vit" in name else "ltt_bert") and rec.version == "beta.1.01"
    assert rec.n_players(cfg) == ocfg.n_players(ocfg.get_config(name))


def test_ladder_reindexing_for_the_bundle():
    from autognothi_b200.recipes.ltt_vit import _ladder_as
    sd = {"vit.encoder.s_attn_maps.0_3.weight": torch.zeros(1), "vit.encoder.s_attn_layers.0_11.output.dense.bias": torch.zeros(1),
          "vit.s_attn_layernorm.0.weight": torch.zeros(1), "vit.encoder.layers.0.output.dense.bias": torch.zeros(1)}
    out = _ladder_as(sd, 0, 1, "vit")
    assert sorted(out) == ["vit.encoder.s_attn_layers.1_11.output.dense.bias", "vit.encoder.s_attn_maps.1_3.weight",
                           "vit.s_attn_layernorm.1.weight"]
