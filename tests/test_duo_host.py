"""Host-side checks of the Duo mirror (reference models/duo_vanilla_{vit,bert}.py): state-dict ABI against tables dumped
from the reference's own classes (tests/golden/make_golden.py duo -> duo_keys.json) and the recipe flags."""
import json
import os

import pytest

from oracle import configs as ocfg


def _recipe(name):
    if name.startswith("vit"):
        from autognothi_b200.recipes.duo_vanilla_vit import duo_vanilla_vit_recipe
        return duo_vanilla_vit_recipe()
    from autognothi_b200.recipes.duo_vanilla_bert import duo_vanilla_bert_recipe
    return duo_vanilla_bert_recipe()


@pytest.mark.parametrize("name", ["vit_mini", "bert_mini"])
def test_duo_state_dict_abi(golden_dir, name):
    with open(os.path.join(golden_dir, "duo_keys.json")) as f:
        ref = json.load(f)[name]
    rec = _recipe(name)
    cfg = rec.t_config(**ocfg.get_config(name))
    assert {k: list(v.shape) for k, v in rec.t_explainer(cfg).state_dict().items()} == ref["explainer"]
    assert {k: list(v.shape) for k, v in rec.t_final(cfg).state_dict().items()} == ref["final"]
    assert rec.training.exp_variant_duo and not rec.measurements.verify_final_coherency
    assert rec.id == ("duo_vanilla_vit" if name.startswith("vit") else "duo_vanilla_bert")
