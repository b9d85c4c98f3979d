"""Drop-in for reference recipes/duo_vanilla_vit.py (dual-objective ViT pipeline; training.exp_variant_duo = True routes
scripts/train_explainer.py:26-27 to train_duo_explainer).  The gradient-similarity analysis of the reference's
`duo_vanilla_vit_inspect` recipe is not part of this path (allow_dual_task_similarity = False)."""
from __future__ import annotations

import dataclasses
from typing import Optional, Tuple

import torch
from torch import Tensor

from ..models.duo_vanilla_vit import (DuoVanillaViTClassifier, DuoVanillaViTConfig, DuoVanillaViTExplainer, DuoVanillaViTFinal,
                                      DuoVanillaViTSurrogate)
from ..models.shapley import MaskLike, PackedMasks
from ._common import copy_matching, resolve_masks
from .types import ModelRecipe, ModelRecipe_Measurements, ModelRecipe_Training
from .vanilla_vit import _fw_classifier, _fw_surrogate, _gen_input, _gen_null, pre_conv_vit


@dataclasses.dataclass
class DuoVanillaViTMisc:
    pass


def _n_players(cfg) -> int:
    return (cfg.img_px_size // cfg.img_patch_size) ** 2


def duo_vanilla_vit_recipe() -> ModelRecipe:
    return ModelRecipe(
        id="duo_vanilla_vit",
        version="beta.1.01",
        t_config=DuoVanillaViTConfig,
        t_classifier=DuoVanillaViTClassifier,
        t_surrogate=DuoVanillaViTSurrogate,
        t_explainer=DuoVanillaViTExplainer,
        t_final=DuoVanillaViTFinal,
        load_misc=lambda m_path, cfg: DuoVanillaViTMisc(),
        conv_pretrained_classifier=_conv_pretrained_classifier,
        conv_classifier_surrogate=_conv_classifier_surrogate,
        conv_surrogate_explainer=_conv_surrogate_explainer,
        conv_explainer_final=_conv_explainer_final,
        n_players=_n_players,
        gen_input=lambda cfg, misc, device: _gen_input(cfg.img_px_size, cfg.img_patch_size, device),
        gen_null=lambda cfg, misc, device: _gen_null(cfg.img_px_size, cfg.img_patch_size, device),
        training=ModelRecipe_Training(True, True, True, True, False),           # exp_variant_duo
        fw_classifier=_fw_classifier,
        fw_surrogate=_fw_surrogate,
        fw_explainer=_fw_explainer,
        fw_final=_fw_final,
        measurements=ModelRecipe_Measurements(False, True, True, True, True, True, True, True, False, True),
    )


def _conv_pretrained_classifier(cfg: DuoVanillaViTConfig, model) -> DuoVanillaViTClassifier:
    v_classifier = pre_conv_vit(cfg.into(), model)
    classifier = DuoVanillaViTClassifier(cfg)
    copy_matching(v_classifier.state_dict(), classifier, ("vit.", "classifier."))
    return classifier


def _conv_classifier_surrogate(cfg, _misc, classifier) -> DuoVanillaViTSurrogate:
    surrogate = DuoVanillaViTSurrogate(cfg).to(next(classifier.parameters()).device)
    copy_matching(classifier.state_dict(), surrogate, ("vit.", "classifier."))
    return surrogate


def _conv_surrogate_explainer(cfg, _misc, surrogate) -> DuoVanillaViTExplainer:
    """reference recipes/duo_vanilla_vit.py:118-141: the explainer KEEPS the classification head"""
    explainer = DuoVanillaViTExplainer(cfg).to(next(surrogate.parameters()).device)
    copy_matching(surrogate.state_dict(), explainer, ("vit.", "classifier."))
    return explainer


def _conv_explainer_final(cfg, misc, classifier, surrogate, explainer) -> DuoVanillaViTFinal:
    """reference recipes/duo_vanilla_vit.py:144-176: surrogate + explainer only (the classifier is not bundled)"""
    device = classifier.vit.embeddings.cls_token.device
    n_players = _n_players(cfg)
    surrogate.eval()
    with torch.no_grad():
        surrogate_null, _ = _fw_surrogate(surrogate, _gen_null(cfg.img_px_size, cfg.img_patch_size, device),
                                          PackedMasks.ones(1, n_players, device))
    final = DuoVanillaViTFinal(cfg).to(device)
    copy_matching(surrogate.state_dict(), final, ("",), "surrogate.")
    copy_matching(explainer.state_dict(), final, ("",), "explainer.")
    with torch.no_grad():
        final.surrogate_null.copy_(surrogate_null)
    return final


def _fw_explainer(model: DuoVanillaViTExplainer, xs: Tensor, mask: MaskLike, surrogate_grand: Tensor, surrogate_null: Tensor
                  ) -> Tuple[Tensor, Optional[Tensor]]:
    pm, S = resolve_masks(xs, mask, _n_players(model.config))
    assert S == 1, "the explainer takes one mask row per input"
    attr, logits = model(xs, pm, surrogate_grand, surrogate_null)
    return attr, logits


def _fw_final(model: DuoVanillaViTFinal, xs: Tensor) -> Tuple[Tensor, Tensor]:
    pm = PackedMasks.ones(xs.shape[0], _n_players(model.config), xs.device)
    return model(xs, pm)
