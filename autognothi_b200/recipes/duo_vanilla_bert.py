"""Drop-in for reference recipes/duo_vanilla_bert.py (dual-objective BERT pipeline); see recipes/duo_vanilla_vit.py."""
from __future__ import annotations

import dataclasses
from typing import Any, Optional, Tuple

import torch
from torch import Tensor

from ..models.duo_vanilla_bert import (DuoVanillaBertClassifier, DuoVanillaBertConfig, DuoVanillaBertExplainer,
                                       DuoVanillaBertFinal, DuoVanillaBertSurrogate)
from ..models.shapley import MaskLike, PackedMasks
from ._common import copy_matching, resolve_masks
from .types import ModelRecipe, ModelRecipe_Measurements, ModelRecipe_Training
from .vanilla_bert import _fw_classifier, _fw_surrogate, _gen_input, _gen_null, pre_conv_bert


@dataclasses.dataclass
class DuoVanillaBertMisc:
    tokenizer: Any = None


def _n_players(cfg) -> int:
    return cfg.max_position_embeddings - 1


def duo_vanilla_bert_recipe() -> ModelRecipe:
    return ModelRecipe(
        id="duo_vanilla_bert",
        version="beta.1.01",
        t_config=DuoVanillaBertConfig,
        t_classifier=DuoVanillaBertClassifier,
        t_surrogate=DuoVanillaBertSurrogate,
        t_explainer=DuoVanillaBertExplainer,
        t_final=DuoVanillaBertFinal,
        load_misc=_load_misc,
        conv_pretrained_classifier=_conv_pretrained_classifier,
        conv_classifier_surrogate=_conv_classifier_surrogate,
        conv_surrogate_explainer=_conv_surrogate_explainer,
        conv_explainer_final=_conv_explainer_final,
        n_players=_n_players,
        gen_input=lambda cfg, misc, device: _gen_input(cfg.max_position_embeddings, misc.tokenizer, device),
        gen_null=lambda cfg, misc, device: _gen_null(cfg.max_position_embeddings, misc.tokenizer, device),
        training=ModelRecipe_Training(True, True, True, True, False),           # exp_variant_duo
        fw_classifier=_fw_classifier,
        fw_surrogate=_fw_surrogate,
        fw_explainer=_fw_explainer,
        fw_final=_fw_final,
        measurements=ModelRecipe_Measurements(False, True, True, True, True, True, True, True, False, True),
    )


def _load_misc(m_path, cfg) -> DuoVanillaBertMisc:
    from transformers import AutoTokenizer  # host-side text preprocessing only
    return DuoVanillaBertMisc(tokenizer=AutoTokenizer.from_pretrained(m_path / "tokenizer"))


_KEEP = ("bert.", "bert_pooler.", "classifier.")


def _conv_pretrained_classifier(cfg: DuoVanillaBertConfig, model) -> DuoVanillaBertClassifier:
    v_classifier = pre_conv_bert(cfg.into(), model)
    classifier = DuoVanillaBertClassifier(cfg)
    copy_matching(v_classifier.state_dict(), classifier, _KEEP)
    return classifier


def _conv_classifier_surrogate(cfg, _misc, classifier) -> DuoVanillaBertSurrogate:
    surrogate = DuoVanillaBertSurrogate(cfg).to(next(classifier.parameters()).device)
    copy_matching(classifier.state_dict(), surrogate, _KEEP)
    return surrogate


def _conv_surrogate_explainer(cfg, _misc, surrogate) -> DuoVanillaBertExplainer:
    """the explainer keeps the pooler and the classification head (reference recipes/duo_vanilla_bert.py:122-146)"""
    explainer = DuoVanillaBertExplainer(cfg).to(next(surrogate.parameters()).device)
    copy_matching(surrogate.state_dict(), explainer, _KEEP)
    return explainer


def _conv_explainer_final(cfg, misc, classifier, surrogate, explainer) -> DuoVanillaBertFinal:
    device = next(classifier.parameters()).device
    n_players = _n_players(cfg)
    surrogate.eval()
    with torch.no_grad():
        surrogate_null, _ = _fw_surrogate(surrogate, _gen_null(cfg.max_position_embeddings, misc.tokenizer, device),
                                          PackedMasks.ones(1, n_players, device))
    final = DuoVanillaBertFinal(cfg).to(device)
    copy_matching(surrogate.state_dict(), final, ("",), "surrogate.")
    copy_matching(explainer.state_dict(), final, ("",), "explainer.")
    with torch.no_grad():
        final.surrogate_null.copy_(surrogate_null)
    return final


def _fw_explainer(model, xs: Tensor, mask: MaskLike, surrogate_grand: Tensor, surrogate_null: Tensor
                  ) -> Tuple[Tensor, Optional[Tensor]]:
    pm, S = resolve_masks(xs, mask, _n_players(model.config))
    assert S == 1, "the explainer takes one mask row per input"
    logits, attr = model(xs, pm, None, surrogate_grand, surrogate_null)
    return attr, logits


def _fw_final(model, xs: Tensor) -> Tuple[Tensor, Tensor]:
    pm = PackedMasks.ones(xs.shape[0], _n_players(model.config), xs.device)
    return model(xs, pm, None)
