"""Drop-in for reference models/duo_vanilla_vit.py: the vanilla ViT pipeline whose explainer ALSO carries a classification
head and is trained with both objectives (cross-entropy + Shapley loss, reference scripts/train_duo_explainer.py:180-196).
Same class names, forward signatures, return order and state-dict keys."""
from __future__ import annotations

from typing import Optional, Tuple

import pydantic
import torch
from torch import Tensor, nn

from .. import engine
from . import _tree
from .shapley import MaskLike
from .vanilla_vit import (VanillaViTClassifier, VanillaViTConfig, VanillaViTSurrogate, _EngineModule, pack_token_mask)


class DuoVanillaViTConfig(pydantic.BaseModel):
    """reference models/duo_vanilla_vit.py:17-58 (identical fields)"""

    attention_probs_dropout_prob: float
    explainer_attn_num_layers: int
    explainer_head_hidden_size: int
    explainer_normalize: bool
    hidden_dropout_prob: float
    hidden_size: int
    intermediate_size: int
    layer_norm_eps: float
    num_attention_heads: int
    num_hidden_layers: int
    num_labels: int
    img_channels: int
    img_px_size: int
    img_patch_size: int

    @property
    def is_decoder(self) -> bool:
        return False

    def into(self) -> VanillaViTConfig:
        return VanillaViTConfig(**self.model_dump())


class DuoVanillaViTClassifier(VanillaViTClassifier):
    """reference models/duo_vanilla_vit.py:61-65"""

    def __init__(self, config: DuoVanillaViTConfig):
        super().__init__(config.into())


class DuoVanillaViTSurrogate(VanillaViTSurrogate):
    """reference models/duo_vanilla_vit.py:68-72"""

    def __init__(self, config: DuoVanillaViTConfig):
        super().__init__(config.into())


class DuoVanillaViTExplainer(_EngineModule):
    """reference models/duo_vanilla_vit.py:75-137 — returns (phi (B, C, n), class probabilities (B, C))"""

    def __init__(self, config: DuoVanillaViTConfig):
        super().__init__()
        self.config = config
        H, C = config.hidden_size, config.num_labels
        _tree.build_tree(self, _tree.vit_backbone_shapes(config) + [("classifier.weight", (C, H)), ("classifier.bias", (C,))]
                         + _tree.explainer_extra_shapes(config, True))

    def forward(self, pixel_values: Tensor, attention_mask: MaskLike, surrogate_grand: Optional[Tensor],
                surrogate_null: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
        words = pack_token_mask(attention_mask, pixel_values.shape[0], engine.n_players_of(self.config))
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .. import training
            return training.duo_explainer_forward_train(self, pixel_values, words, surrogate_grand, surrogate_null)
        return self._engine(engine.DuoExplainerEngine).duo(pixel_values, words, surrogate_grand, surrogate_null)


class DuoVanillaViTFinal(nn.Module):
    """reference models/duo_vanilla_vit.py:140-175 — the class output comes from the explainer's own head"""

    def __init__(self, config: DuoVanillaViTConfig):
        super().__init__()
        self.config = config
        self.surrogate = VanillaViTSurrogate(config.into())
        self.surrogate_null = nn.Parameter(torch.zeros((1, config.num_labels)), requires_grad=False)
        self.explainer = DuoVanillaViTExplainer(config)

    def forward(self, pixel_values: Tensor, attention_mask: MaskLike) -> Tuple[Tensor, Tensor]:
        grand = self.surrogate(pixel_values, attention_mask) if self.config.explainer_normalize else None
        phi, logits = self.explainer(pixel_values, attention_mask, grand, self.surrogate_null)
        return logits, phi
