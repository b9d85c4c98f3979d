"""Drop-in for reference recipes/froyo_bert.py (ModelRecipe of the frozen-backbone BERT pipeline)."""
from __future__ import annotations

import dataclasses
from typing import Any, Optional, Tuple

import torch
from torch import Tensor

from ..models.froyo_bert import (FroyoBertClassifier, FroyoBertConfig, FroyoBertExplainer, FroyoBertFinal,
                                 FroyoBertSurrogate)
from ..models.shapley import MaskLike, PackedMasks
from ._common import copy_matching, resolve_masks
from .types import ModelRecipe, ModelRecipe_Measurements, ModelRecipe_Training
from .vanilla_bert import _gen_input, _gen_null, pre_conv_bert


@dataclasses.dataclass
class FroyoBertMisc:
    tokenizer: Any = None


def _n_players(cfg) -> int:
    return cfg.max_position_embeddings - 1  # reference recipes/froyo_bert.py:52


def froyo_bert_recipe() -> ModelRecipe:
    return ModelRecipe(
        id="froyo_bert",
        version="beta.1.01",
        t_config=FroyoBertConfig,
        t_classifier=FroyoBertClassifier,
        t_surrogate=FroyoBertSurrogate,
        t_explainer=FroyoBertExplainer,
        t_final=FroyoBertFinal,
        load_misc=_load_misc,
        conv_pretrained_classifier=_conv_pretrained_classifier,
        conv_classifier_surrogate=_conv_classifier_surrogate,
        conv_surrogate_explainer=_conv_surrogate_explainer,
        conv_explainer_final=_conv_explainer_final,
        n_players=_n_players,
        gen_input=lambda cfg, misc, device: _gen_input(cfg.max_position_embeddings, misc.tokenizer, device),
        gen_null=lambda cfg, misc, device: _gen_null(cfg.max_position_embeddings, misc.tokenizer, device),
        training=ModelRecipe_Training(True, True, True, False, False),
        fw_classifier=_fw_classifier,
        fw_surrogate=_fw_surrogate,
        fw_explainer=_fw_explainer,
        fw_final=_fw_final,
        measurements=ModelRecipe_Measurements(True, True, True, True, True, True, True, True, False, True),
    )


def _load_misc(m_path, cfg) -> FroyoBertMisc:
    """reference recipes/froyo_bert.py:88-90: the tokenizer stored under <base model>/tokenizer"""
    from transformers import AutoTokenizer  # host-side text preprocessing only
    return FroyoBertMisc(tokenizer=AutoTokenizer.from_pretrained(m_path / "tokenizer"))


def _conv_pretrained_classifier(cfg: FroyoBertConfig, model) -> FroyoBertClassifier:
    v_classifier = pre_conv_bert(cfg.into(), model)
    classifier = FroyoBertClassifier(cfg)
    copy_matching(v_classifier.state_dict(), classifier, ("bert.", "bert_pooler.", "classifier."))
    return classifier


def _conv_classifier_surrogate(cfg, _misc, classifier) -> FroyoBertSurrogate:
    surrogate = FroyoBertSurrogate(cfg).to(next(classifier.parameters()).device)
    copy_matching(classifier.state_dict(), surrogate, ("bert.", "bert_pooler.", "classifier."))
    return surrogate


def _conv_surrogate_explainer(cfg, _misc, surrogate) -> FroyoBertExplainer:
    explainer = FroyoBertExplainer(cfg).to(next(surrogate.parameters()).device)
    copy_matching(surrogate.state_dict(), explainer, ("bert.",))
    return explainer


def _conv_explainer_final(cfg, misc, classifier, surrogate, explainer) -> FroyoBertFinal:
    """reference recipes/froyo_bert.py:148-193: the surrogate's pooler / head become `srg_bert_pooler.*` / `srg_classifier.*`"""
    device = next(classifier.parameters()).device
    n_players = _n_players(cfg)
    nil_xs = _gen_null(cfg.max_position_embeddings, misc.tokenizer, device)
    surrogate.eval()
    with torch.no_grad():
        surrogate_null, _ = _fw_surrogate(surrogate, nil_xs, PackedMasks.ones(1, n_players, device))
    final = FroyoBertFinal(cfg).to(device)
    copy_matching(classifier.state_dict(), final, ("bert.", "bert_pooler.", "classifier."))
    copy_matching({k: v for k, v in surrogate.state_dict().items() if k.startswith(("bert_pooler.", "classifier."))},
                  final, ("",), "srg_")
    copy_matching(explainer.state_dict(), final, ("explainer_attn.", "explainer_mlp."))
    with torch.no_grad():
        final.surrogate_null.copy_(surrogate_null)
    return final


def _fw_classifier(model, xs: Tensor, mask: MaskLike) -> Tuple[Tensor, Tensor]:
    pm, S = resolve_masks(xs, mask, _n_players(model.config))
    probs = model(xs, pm, None, n_mask_samples=S)
    return probs, probs


def _fw_surrogate(model, xs: Tensor, mask: MaskLike) -> Tuple[Tensor, Optional[Tensor]]:
    pm, S = resolve_masks(xs, mask, _n_players(model.config))
    return model(xs, pm, None, n_mask_samples=S), None


def _fw_explainer(model, xs: Tensor, mask: MaskLike, surrogate_grand: Tensor, surrogate_null: Tensor
                  ) -> Tuple[Tensor, Optional[Tensor]]:
    pm, S = resolve_masks(xs, mask, _n_players(model.config))
    assert S == 1, "the explainer takes one mask row per input"
    return model(xs, pm, None, surrogate_grand, surrogate_null), None


def _fw_final(model, xs: Tensor) -> Tuple[Tensor, Tensor]:
    pm = PackedMasks.ones(xs.shape[0], _n_players(model.config), xs.device)
    return model(xs, pm, None)
