"""Drop-in for reference recipes/ltt_bert.py (ModelRecipe of the ladder-side-tuning BERT pipeline)."""
from __future__ import annotations

import dataclasses
from typing import Any, Optional, Tuple

import torch
from torch import Tensor

from ..models.ltt_bert import LttBertConfig, LttBertExplainer, LttBertFinal, LttBertSurrogate
from ..models.shapley import MaskLike, PackedMasks
from ._common import copy_matching, resolve_masks
from .ltt_vit import _ladder_as
from .types import ModelRecipe, ModelRecipe_Measurements, ModelRecipe_Training
from .vanilla_bert import _gen_input, _gen_null, pre_conv_bert


@dataclasses.dataclass
class LttBertMisc:
    tokenizer: Any = None


def _n_players(cfg) -> int:
    return cfg.max_position_embeddings - 1  # reference recipes/ltt_bert.py:49


def ltt_bert_recipe() -> ModelRecipe:
    return ModelRecipe(
        id="ltt_bert",
        version="beta.1.01",
        t_config=LttBertConfig,
        t_classifier=LttBertSurrogate,     # sic (reference recipes/ltt_bert.py:41)
        t_surrogate=LttBertSurrogate,
        t_explainer=LttBertExplainer,
        t_final=LttBertFinal,
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


def _load_misc(m_path, cfg) -> LttBertMisc:
    from transformers import AutoTokenizer  # host-side text preprocessing only
    return LttBertMisc(tokenizer=AutoTokenizer.from_pretrained(m_path / "tokenizer"))


_BACKBONE = ("bert.embeddings.", "bert.encoder.layers.", "bert_pooler.", "classifier.")


def _conv_pretrained_classifier(cfg: LttBertConfig, model) -> LttBertSurrogate:
    """reference recipes/ltt_bert.py:92-119"""
    v_classifier = pre_conv_bert(cfg.into(), model)
    classifier = LttBertSurrogate(cfg)
    copy_matching(v_classifier.state_dict(), classifier, _BACKBONE)
    return classifier


def _conv_classifier_surrogate(cfg, _misc, classifier) -> LttBertSurrogate:
    """reference recipes/ltt_bert.py:122-134"""
    surrogate = LttBertSurrogate(cfg).to(next(classifier.parameters()).device)
    copy_matching(classifier.state_dict(), surrogate, ("bert.", "bert_pooler.", "classifier.", "bert_s_attn_pooler.", "s_attn_classifier."))
    return surrogate


def _conv_surrogate_explainer(cfg, _misc, surrogate) -> LttBertExplainer:
    """reference recipes/ltt_bert.py:137-162: keeps the surrogate's ladder, drops its pooler/head"""
    explainer = LttBertExplainer(cfg).to(next(surrogate.parameters()).device)
    copy_matching(surrogate.state_dict(), explainer, ("bert.", "bert_pooler.", "classifier."))
    return explainer


def _conv_explainer_final(cfg, misc, classifier, surrogate, explainer) -> LttBertFinal:
    """reference recipes/ltt_bert.py:165-262"""
    device = next(classifier.parameters()).device
    n_players = _n_players(cfg)
    nil_xs = _gen_null(cfg.max_position_embeddings, misc.tokenizer, device)
    surrogate.eval()
    with torch.no_grad():
        surrogate_null, _ = _fw_surrogate(surrogate, nil_xs, PackedMasks.ones(1, n_players, device))
    final = LttBertFinal(cfg).to(device)
    copy_matching(classifier.state_dict(), final, _BACKBONE)
    copy_matching(_ladder_as(surrogate.state_dict(), 0, 0, "bert"), final, ("",))
    copy_matching(surrogate.state_dict(), final, ("bert_s_attn_pooler.", "s_attn_classifier."))
    copy_matching(_ladder_as(explainer.state_dict(), 0, 1, "bert"), final, ("",))
    copy_matching(explainer.state_dict(), final, ("s_attn_attention_layers.", "s_attn_explainer."))
    with torch.no_grad():
        final.surrogate_null.copy_(surrogate_null)
    return final


def _fw_classifier(model, xs: Tensor, mask: MaskLike) -> Tuple[Tensor, Tensor]:
    pm, S = resolve_masks(xs, mask, _n_players(model.config))
    side, main = model(xs, pm, None, n_mask_samples=S)
    return side, main


def _fw_surrogate(model, xs: Tensor, mask: MaskLike) -> Tuple[Tensor, Optional[Tensor]]:
    pm, S = resolve_masks(xs, mask, _n_players(model.config))
    side, main = model(xs, pm, None, n_mask_samples=S)
    return side, main


def _fw_explainer(model, xs: Tensor, mask: MaskLike, surrogate_grand: Tensor, surrogate_null: Tensor
                  ) -> Tuple[Tensor, Optional[Tensor]]:
    pm, S = resolve_masks(xs, mask, _n_players(model.config))
    assert S == 1, "the explainer takes one mask row per input"
    attr, main = model(xs, pm, None, surrogate_grand, surrogate_null)
    return attr, main


def _fw_final(model, xs: Tensor) -> Tuple[Tensor, Tensor]:
    pm = PackedMasks.ones(xs.shape[0], _n_players(model.config), xs.device)
    return model(xs, pm, None)
