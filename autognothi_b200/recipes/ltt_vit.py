"""Drop-in for reference recipes/ltt_vit.py (ModelRecipe of the ladder-side-tuning ViT pipeline)."""
from __future__ import annotations

import dataclasses
from typing import Optional, Tuple

import torch
from torch import Tensor

from ..models.ltt_vit import LttViTConfig, LttViTExplainer, LttViTFinal, LttViTSurrogate
from ..models.shapley import MaskLike, PackedMasks
from ._common import copy_matching, resolve_masks
from .types import ModelRecipe, ModelRecipe_Measurements, ModelRecipe_Training
from .vanilla_vit import _gen_input, _gen_null, pre_conv_vit


@dataclasses.dataclass
class LttViTMisc:
    pass


def _n_players(cfg) -> int:
    return (cfg.img_px_size // cfg.img_patch_size) ** 2  # reference recipes/ltt_vit.py:41


def ltt_vit_recipe() -> ModelRecipe:
    return ModelRecipe(
        id="ltt_vit",
        version="beta.1.01",
        t_config=LttViTConfig,
        t_classifier=LttViTSurrogate,      # sic: the classifier of this pipeline is the surrogate class (reference l.33)
        t_surrogate=LttViTSurrogate,
        t_explainer=LttViTExplainer,
        t_final=LttViTFinal,
        load_misc=lambda m_path, cfg: LttViTMisc(),
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


def _conv_pretrained_classifier(cfg: LttViTConfig, model) -> LttViTSurrogate:
    """reference recipes/ltt_vit.py:82-106: backbone + classifier from the pretrained model, fresh side ladder"""
    v_classifier = pre_conv_vit(cfg.into(), model)
    classifier = LttViTSurrogate(cfg)
    copy_matching(v_classifier.state_dict(), classifier, ("vit.embeddings.", "vit.encoder.layers.", "vit.layernorm.", "classifier."))
    return classifier


def _conv_classifier_surrogate(cfg, _misc, classifier) -> LttViTSurrogate:
    """reference recipes/ltt_vit.py:109-119: everything carries over"""
    surrogate = LttViTSurrogate(cfg).to(next(classifier.parameters()).device)
    copy_matching(classifier.state_dict(), surrogate, ("vit.", "classifier.", "s_attn_classifier."))
    return surrogate


def _conv_surrogate_explainer(cfg, _misc, surrogate) -> LttViTExplainer:
    """reference recipes/ltt_vit.py:122-139: the explainer STARTS from the surrogate's side ladder (vit.* incl. the
    ladder), drops the side classifier, fresh s_explainer_* parameters"""
    explainer = LttViTExplainer(cfg).to(next(surrogate.parameters()).device)
    copy_matching(surrogate.state_dict(), explainer, ("vit.", "classifier."))
    return explainer


def _ladder_as(sd, src: int, dst: int, root: str):
    """re-index side ladder `src` of a state dict as ladder `dst` (reference recipes/ltt_vit.py:207-219)"""
    out = {}
    for k, v in sd.items():
        for stem in (f"{root}.encoder.s_attn_maps.", f"{root}.encoder.s_attn_layers."):
            if k.startswith(stem + f"{src}_"):
                out[stem + f"{dst}_" + k[len(stem) + len(f"{src}_"):]] = v
        stem = f"{root}.s_attn_layernorm."
        if k.startswith(stem + f"{src}."):
            out[stem + f"{dst}." + k[len(stem) + len(f"{src}."):]] = v
    return out


def _conv_explainer_final(cfg, misc, classifier, surrogate, explainer) -> LttViTFinal:
    """reference recipes/ltt_vit.py:142-232: backbone + classifier head from the classifier, the surrogate's ladder as
    ladder 0 (+ its side head), the explainer's ladder as ladder 1 (+ its side explainer), the replayed null value."""
    device = classifier.vit.embeddings.cls_token.device
    n_players = _n_players(cfg)
    nil_xs = _gen_null(cfg.img_px_size, cfg.img_patch_size, device)
    surrogate.eval()
    with torch.no_grad():
        surrogate_null, _ = _fw_surrogate(surrogate, nil_xs, PackedMasks.ones(1, n_players, device))
    final = LttViTFinal(cfg).to(device)
    copy_matching(classifier.state_dict(), final, ("vit.embeddings.", "vit.encoder.layers.", "vit.layernorm.", "classifier."))
    copy_matching(_ladder_as(surrogate.state_dict(), 0, 0, "vit"), final, ("",))
    copy_matching(surrogate.state_dict(), final, ("s_attn_classifier.",))
    copy_matching(_ladder_as(explainer.state_dict(), 0, 1, "vit"), final, ("",))
    copy_matching(explainer.state_dict(), final, ("s_explainer_attn.", "s_explainer_mlp."))
    with torch.no_grad():
        final.surrogate_null.copy_(surrogate_null)
    return final


def _fw_classifier(model: LttViTSurrogate, xs: Tensor, mask: MaskLike) -> Tuple[Tensor, Tensor]:
    pm, S = resolve_masks(xs, mask, _n_players(model.config))
    side, main = model(xs, pm, n_mask_samples=S)
    return side, main


def _fw_surrogate(model: LttViTSurrogate, xs: Tensor, mask: MaskLike) -> Tuple[Tensor, Optional[Tensor]]:
    pm, S = resolve_masks(xs, mask, _n_players(model.config))
    side, main = model(xs, pm, n_mask_samples=S)
    return side, main


def _fw_explainer(model: LttViTExplainer, xs: Tensor, mask: MaskLike, surrogate_grand: Tensor, surrogate_null: Tensor
                  ) -> Tuple[Tensor, Optional[Tensor]]:
    pm, S = resolve_masks(xs, mask, _n_players(model.config))
    assert S == 1, "the explainer takes one mask row per input"
    attr, main = model(xs, pm, surrogate_grand, surrogate_null)
    return attr, main


def _fw_final(model: LttViTFinal, xs: Tensor) -> Tuple[Tensor, Tensor]:
    pm = PackedMasks.ones(xs.shape[0], _n_players(model.config), xs.device)
    return model(xs, pm)
