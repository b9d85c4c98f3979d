"""Drop-in for reference recipes/froyo_vit.py (ModelRecipe of the frozen-backbone ViT pipeline)."""
from __future__ import annotations

import dataclasses
from typing import Optional, Tuple

import torch
from torch import Tensor

from ..models.froyo_vit import FroyoViTClassifier, FroyoViTConfig, FroyoViTExplainer, FroyoViTFinal, FroyoViTSurrogate
from ..models.shapley import MaskLike, PackedMasks
from ._common import copy_matching, resolve_masks
from .types import ModelRecipe, ModelRecipe_Measurements, ModelRecipe_Training
from .vanilla_vit import _gen_input, _gen_null, pre_conv_vit


@dataclasses.dataclass
class FroyoViTMisc:
    pass


def _n_players(cfg) -> int:
    return (cfg.img_px_size // cfg.img_patch_size) ** 2  # reference recipes/froyo_vit.py:49


def froyo_vit_recipe() -> ModelRecipe:
    return ModelRecipe(
        id="froyo_vit",
        version="beta.1.01",
        t_config=FroyoViTConfig,
        t_classifier=FroyoViTClassifier,
        t_surrogate=FroyoViTSurrogate,
        t_explainer=FroyoViTExplainer,
        t_final=FroyoViTFinal,
        load_misc=lambda m_path, cfg: FroyoViTMisc(),
        conv_pretrained_classifier=_conv_pretrained_classifier,
        conv_classifier_surrogate=_conv_classifier_surrogate,
        conv_surrogate_explainer=_conv_surrogate_explainer,
        conv_explainer_final=_conv_explainer_final,
        n_players=_n_players,
        gen_input=lambda cfg, misc, device: _gen_input(cfg.img_px_size, cfg.img_patch_size, device),
        gen_null=lambda cfg, misc, device: _gen_null(cfg.img_px_size, cfg.img_patch_size, device),
        training=ModelRecipe_Training(True, True, True, False, False),
        fw_classifier=_fw_classifier,
        fw_surrogate=_fw_surrogate,
        fw_explainer=_fw_explainer,
        fw_final=_fw_final,
        measurements=ModelRecipe_Measurements(True, True, True, True, True, True, True, True, False, True),
    )


def _conv_pretrained_classifier(cfg: FroyoViTConfig, model) -> FroyoViTClassifier:
    """reference recipes/froyo_vit.py:92-103"""
    v_classifier = pre_conv_vit(cfg.into(), model)
    classifier = FroyoViTClassifier(cfg)
    copy_matching(v_classifier.state_dict(), classifier, ("vit.", "classifier."))
    return classifier


def _conv_classifier_surrogate(cfg, _misc, classifier) -> FroyoViTSurrogate:
    """reference recipes/froyo_vit.py:106-116: keep vit.* and classifier.*"""
    surrogate = FroyoViTSurrogate(cfg).to(next(classifier.parameters()).device)
    copy_matching(classifier.state_dict(), surrogate, ("vit.", "classifier."))
    return surrogate


def _conv_surrogate_explainer(cfg, _misc, surrogate) -> FroyoViTExplainer:
    """reference recipes/froyo_vit.py:119-143: keep vit.*, drop classifier.*, fresh explainer_* parameters"""
    explainer = FroyoViTExplainer(cfg).to(next(surrogate.parameters()).device)
    copy_matching(surrogate.state_dict(), explainer, ("vit.",))
    return explainer


def _conv_explainer_final(cfg, misc, classifier, surrogate, explainer) -> FroyoViTFinal:
    """reference recipes/froyo_vit.py:146-190: replay the surrogate on the null input; the backbone and `classifier.*`
    come from the classifier, the surrogate's head becomes `srg_classifier.*`, the explainer contributes its tail."""
    device = classifier.vit.embeddings.cls_token.device
    n_players = _n_players(cfg)
    nil_xs = _gen_null(cfg.img_px_size, cfg.img_patch_size, device)
    surrogate.eval()
    with torch.no_grad():
        surrogate_null, _ = _fw_surrogate(surrogate, nil_xs, PackedMasks.ones(1, n_players, device))
    final = FroyoViTFinal(cfg).to(device)
    copy_matching(classifier.state_dict(), final, ("vit.", "classifier."))
    copy_matching({k: v for k, v in surrogate.state_dict().items() if k.startswith("classifier.")}, final, ("",), "srg_")
    copy_matching(explainer.state_dict(), final, ("explainer_attn.", "explainer_mlp."))
    with torch.no_grad():
        final.surrogate_null.copy_(surrogate_null)
    return final


def _fw_classifier(model: FroyoViTClassifier, xs: Tensor, mask: MaskLike) -> Tuple[Tensor, Tensor]:
    pm, S = resolve_masks(xs, mask, _n_players(model.config))
    probs = model(xs, pm, n_mask_samples=S)
    return probs, probs


def _fw_surrogate(model: FroyoViTSurrogate, xs: Tensor, mask: MaskLike) -> Tuple[Tensor, Optional[Tensor]]:
    pm, S = resolve_masks(xs, mask, _n_players(model.config))
    return model(xs, pm, n_mask_samples=S), None


def _fw_explainer(model: FroyoViTExplainer, xs: Tensor, mask: MaskLike, surrogate_grand: Tensor, surrogate_null: Tensor
                  ) -> Tuple[Tensor, Optional[Tensor]]:
    pm, S = resolve_masks(xs, mask, _n_players(model.config))
    assert S == 1, "the explainer takes one mask row per input"
    return model(xs, pm, surrogate_grand, surrogate_null), None


def _fw_final(model: FroyoViTFinal, xs: Tensor) -> Tuple[Tensor, Tensor]:
    pm = PackedMasks.ones(xs.shape[0], _n_players(model.config), xs.device)
    return model(xs, pm)
