"""Drop-in for reference recipes/kernel_shap_bert.py: the KernelSHAP baseline recipe."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from ..models.kernel_shap_bert import (KernelShapBertClassifier, KernelShapBertConfig, KernelShapBertExplainer,
                                       KernelShapBertFinal, KernelShapBertSurrogate, kernel_shap_torch)
from ..models.shapley import MaskLike, PackedMasks
from . import vanilla_bert as vb
from ._common import copy_matching, resolve_masks
from .types import ModelRecipe, ModelRecipe_Measurements, ModelRecipe_Training


def _n_players(cfg) -> int:
    return cfg.max_position_embeddings - 1


def kernel_shap_bert_recipe() -> ModelRecipe:
    """reference recipes/kernel_shap_bert.py:38-86"""
    return ModelRecipe(
        id="kernel_shap_bert",
        version="beta.1.01",
        t_config=KernelShapBertConfig,
        t_classifier=KernelShapBertClassifier,
        t_surrogate=KernelShapBertSurrogate,
        t_explainer=KernelShapBertExplainer,
        t_final=KernelShapBertFinal,
        load_misc=vb._load_misc,
        conv_pretrained_classifier=_conv_pretrained_classifier,
        conv_classifier_surrogate=_conv_classifier_surrogate,
        conv_surrogate_explainer=lambda cfg, misc, srg: KernelShapBertExplainer(cfg).to(next(srg.parameters()).device),
        conv_explainer_final=_conv_explainer_final,
        n_players=_n_players,
        gen_input=lambda cfg, misc, device: vb._gen_input(cfg.max_position_embeddings, misc.tokenizer, device),
        gen_null=lambda cfg, misc, device: vb._gen_null(cfg.max_position_embeddings, misc.tokenizer, device),
        training=ModelRecipe_Training(True, False, False, False, True),
        fw_classifier=_fw_classifier,
        fw_surrogate=_fw_classifier_opt,
        fw_explainer=_fw_explainer,
        fw_final=_fw_final,
        measurements=ModelRecipe_Measurements(False, True, True, True, True, False, True, False, False, False),
    )


def _conv_pretrained_classifier(cfg, model) -> KernelShapBertClassifier:
    classifier = KernelShapBertClassifier(cfg)
    src = vb.pre_conv_bert(cfg.into(), model)
    copy_matching(src.state_dict(), classifier, ("bert.", "bert_pooler.", "classifier."))
    return classifier


def _conv_classifier_surrogate(cfg, _misc, classifier) -> KernelShapBertSurrogate:
    surrogate = KernelShapBertSurrogate(cfg).to(next(classifier.parameters()).device)
    copy_matching(classifier.state_dict(), surrogate, ("bert.", "bert_pooler.", "classifier."))
    return surrogate


def _conv_explainer_final(cfg, misc, classifier, surrogate, explainer) -> KernelShapBertFinal:
    final = KernelShapBertFinal(cfg).to(next(classifier.parameters()).device)
    copy_matching(classifier.state_dict(), final, ("",), "classifier.")
    copy_matching(explainer.state_dict(), final, ("",), "explainer.")
    return final


def _fw_classifier(model, xs: Tensor, mask: MaskLike) -> Tuple[Tensor, Tensor]:
    pm, S = resolve_masks(xs, mask, _n_players(model.config))
    probs = model(xs, pm, None, n_mask_samples=S)
    return probs, probs


def _fw_classifier_opt(model, xs: Tensor, mask: MaskLike) -> Tuple[Tensor, Optional[Tensor]]:
    return _fw_classifier(model, xs, mask)[0], None


def _fw_explainer(model, xs, mask, surrogate_grand, surrogate_null):
    """reference recipes/kernel_shap_bert.py:165-172: the KernelSHAP 'explainer' has no forward of its own."""
    raise NotImplementedError("KernelSHAP explanations are produced by fw_final (kernel_shap_torch)")


def _fw_final(model: KernelShapBertFinal, xs: Tensor) -> Tuple[Tensor, Tensor]:
    """reference recipes/kernel_shap_bert.py:175-188: plain classifier logits (no attention masking) + KernelSHAP."""
    cfg = model.config
    n = _n_players(cfg)

    def fw(ids: Tensor) -> Tensor:
        return model.classifier(ids, PackedMasks.ones(ids.shape[0], n, ids.device), None)

    logits = fw(xs)
    attr = kernel_shap_torch(fw, model.explainer.Xs_train, xs, n_samples=cfg.kernel_shap_n_samples,
                             batch_size=64, silent=True)
    return logits, attr
