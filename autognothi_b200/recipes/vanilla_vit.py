"""Drop-in for reference recipes/vanilla_vit.py (the ModelRecipe of the vanilla ViT pipeline)."""
from __future__ import annotations

import dataclasses
import re
from typing import Any, Callable, List, Optional, Tuple

import torch
from torch import Tensor, nn

from ..models.shapley import MaskLike, PackedMasks
from ..models.vanilla_vit import (VanillaViTClassifier, VanillaViTConfig, VanillaViTExplainer, VanillaViTFinal,
                                  VanillaViTSurrogate)
from ._common import copy_matching, resolve_masks
from .types import ModelRecipe, ModelRecipe_Measurements, ModelRecipe_Training


@dataclasses.dataclass
class VanillaViTMisc:
    pass


def _n_players(cfg: VanillaViTConfig) -> int:
    return (cfg.img_px_size // cfg.img_patch_size) ** 2  # reference recipes/vanilla_vit.py:49


def vanilla_vit_recipe() -> ModelRecipe:
    return ModelRecipe(
        id="vanilla_bert",  # sic — the reference's id string (recipes/vanilla_vit.py:37)
        version="beta.1.01",
        t_config=VanillaViTConfig,
        t_classifier=VanillaViTClassifier,
        t_surrogate=VanillaViTSurrogate,
        t_explainer=VanillaViTExplainer,
        t_final=VanillaViTFinal,
        load_misc=lambda m_path, cfg: VanillaViTMisc(),
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


_HF_RULES = [  # HF ViTForImageClassification -> ours (reference recipes/vanilla_vit.py:94-109)
    (r"^vit\.encoder\.layer\.(\d+)\.attention\.attention\.(query|key|value)\.(weight|bias)$",
     r"vit.encoder.layers.\1.attention.self.\2.\3"),
    (r"^vit\.encoder\.layer\.(\d+)\.(attention\.output\.dense|intermediate\.dense|output\.dense|layernorm_before|layernorm_after)\.(weight|bias)$",
     r"vit.encoder.layers.\1.\2.\3"),
]


def pre_conv_vit(cfg: VanillaViTConfig, model: Any) -> VanillaViTClassifier:
    """reference recipes/vanilla_vit.py:90-113: accepts one of our classifiers or an HF ViT state dict /
    module; the classification head of a foreign checkpoint is dropped and re-initialised."""
    classifier = VanillaViTClassifier(cfg)
    sd = model.state_dict() if isinstance(model, nn.Module) else dict(model)
    if any(k.startswith("vit.encoder.layers.") for k in sd):
        copy_matching(sd, classifier, ("vit.", "classifier."))
        return classifier
    renamed = {}
    for k, v in sd.items():
        if k.startswith("classifier."):
            continue
        for pat, rep in _HF_RULES:
            if re.match(pat, k):
                k = re.sub(pat, rep, k)
                break
        renamed[k] = v
    copy_matching(renamed, classifier, ("vit.",))
    return classifier


def _conv_pretrained_classifier(cfg, model) -> VanillaViTClassifier:
    return pre_conv_vit(cfg, model)


def _conv_classifier_surrogate(cfg, _misc, classifier) -> VanillaViTSurrogate:
    """reference recipes/vanilla_vit.py:123-134: keep vit.* and classifier.*"""
    surrogate = VanillaViTSurrogate(cfg).to(next(classifier.parameters()).device)
    copy_matching(classifier.state_dict(), surrogate, ("vit.", "classifier."))
    return surrogate


def _conv_surrogate_explainer(cfg, _misc, surrogate) -> VanillaViTExplainer:
    """reference recipes/vanilla_vit.py:137-158: keep vit.*, drop classifier.*, new explainer_* params"""
    explainer = VanillaViTExplainer(cfg).to(next(surrogate.parameters()).device)
    copy_matching(surrogate.state_dict(), explainer, ("vit.",))
    return explainer


def _conv_explainer_final(cfg, misc, classifier, surrogate, explainer) -> VanillaViTFinal:
    """reference recipes/vanilla_vit.py:161-194: replay the surrogate on the null input, bundle all three."""
    device = classifier.vit.embeddings.cls_token.device
    n_players = _n_players(cfg)
    nil_xs = _gen_null(cfg.img_px_size, cfg.img_patch_size, device)
    surrogate.eval()
    with torch.no_grad():
        surrogate_null, _ = _fw_surrogate(surrogate, nil_xs, PackedMasks.ones(1, n_players, device))
    final = VanillaViTFinal(cfg).to(device)
    copy_matching(classifier.state_dict(), final, ("",), "classifier.")
    copy_matching(surrogate.state_dict(), final, ("",), "surrogate.")
    copy_matching(explainer.state_dict(), final, ("",), "explainer.")
    with torch.no_grad():
        final.surrogate_null.copy_(surrogate_null)
    return final


def _gen_input(img_px_size: int, img_patch_size: int, device) -> Callable[[Any, Any], Tuple[Tensor, Tensor]]:
    """collate: list of (C,px,px) tensors + labels -> (B,C,px,px) on device (reference l.197-210)"""

    def mask_input(raw_xs: List[Tensor], raw_ys: List[int]):
        xs = torch.stack(raw_xs, dim=0).to(device, non_blocking=True)
        ys = torch.tensor(raw_ys).to(device, non_blocking=True)
        return xs, ys

    return mask_input


def _gen_null(img_px_size: int, img_patch_size: int, device) -> Tensor:
    """zero image (reference l.213-216)"""
    return torch.zeros((1, 3, img_px_size, img_px_size), device=device)


def _fw_classifier(model: VanillaViTClassifier, xs: Tensor, mask: MaskLike) -> Tuple[Tensor, Tensor]:
    pm, S = resolve_masks(xs, mask, _n_players(model.config))
    probs = model(xs, pm, n_mask_samples=S)
    return probs, probs


def _fw_surrogate(model: VanillaViTSurrogate, xs: Tensor, mask: MaskLike) -> Tuple[Tensor, Optional[Tensor]]:
    pm, S = resolve_masks(xs, mask, _n_players(model.config))
    return model(xs, pm, n_mask_samples=S), None


def _fw_explainer(model: VanillaViTExplainer, xs: Tensor, mask: MaskLike, surrogate_grand: Tensor,
                  surrogate_null: Tensor) -> Tuple[Tensor, Optional[Tensor]]:
    pm, S = resolve_masks(xs, mask, _n_players(model.config))
    assert S == 1, "the explainer takes one mask row per input"
    return model(xs, pm, surrogate_grand, surrogate_null), None


def _fw_final(model: VanillaViTFinal, xs: Tensor) -> Tuple[Tensor, Tensor]:
    pm = PackedMasks.ones(xs.shape[0], _n_players(model.config), xs.device)
    return model(xs, pm)
