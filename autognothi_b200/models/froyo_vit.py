"""Drop-in for reference models/froyo_vit.py ("FroYo": the backbone is frozen, only the heads train; otherwise the
vanilla ViT pipeline).  Same class names, constructor/forward signatures and state-dict keys.

What the freeze buys on this implementation: the frozen encoder stack runs on the inference engine (LayerNorm-folded
tcgen05 GEMMs, first-block sharing) without keeping a tape, and the training adjoint stops at the first
`explainer_attn` block (autognothi_b200/training.py `train_backbone=False`); `FroyoViTFinal` feeds the classifier
head, the surrogate head and the explainer tail from ONE backbone pass.
"""
from __future__ import annotations

from typing import Optional, Tuple

import pydantic
import torch
from torch import Tensor, nn

from .. import engine
from . import _tree
from .shapley import MaskLike
from .vanilla_vit import (VanillaViTClassifier, VanillaViTConfig, VanillaViTExplainer, VanillaViTSurrogate,
                          _EngineModule, pack_token_mask)


class FroyoViTConfig(pydantic.BaseModel):
    """reference models/froyo_vit.py:19-58 (identical fields)"""

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


class FroyoViTClassifier(VanillaViTClassifier):
    """reference models/froyo_vit.py:63-73"""

    def __init__(self, config: FroyoViTConfig):
        super().__init__(config.into())

    def train(self, mode: bool = True):
        nn.Module.train(self, mode)
        _tree.freeze_model_parameters(self, "vit")
        _tree.freeze_model_parameters(self, "classifier")
        return self


class FroyoViTSurrogate(VanillaViTSurrogate):
    """reference models/froyo_vit.py:76-85 — only `classifier.*` trains"""

    def __init__(self, config: FroyoViTConfig):
        super().__init__(config.into())

    def train(self, mode: bool = True):
        nn.Module.train(self, mode)
        _tree.freeze_model_parameters(self, "vit")
        return self


class FroyoViTExplainer(VanillaViTExplainer):
    """reference models/froyo_vit.py:88-97 — only `explainer_attn.*` / `explainer_mlp.*` train"""

    def __init__(self, config: FroyoViTConfig):
        super().__init__(config.into())

    def train(self, mode: bool = True):
        nn.Module.train(self, mode)
        _tree.freeze_model_parameters(self, "vit")
        return self


class FroyoViTFinal(_EngineModule):
    """reference models/froyo_vit.py:100-177.  The reference's `forward` lists `surrogate_grand` / `surrogate_null`
    as required arguments although it ignores them (l.165-166 overwrite both) and its own recipe calls
    `model(xs, mask)` (recipes/froyo_vit.py:222); they are optional here so that both spellings work."""

    def __init__(self, config: FroyoViTConfig):
        super().__init__()
        self.config = config
        H, C = config.hidden_size, config.num_labels
        _tree.build_tree(self, _tree.vit_backbone_shapes(config)
                         + [("classifier.weight", (C, H)), ("classifier.bias", (C,)),
                            ("srg_classifier.weight", (C, H)), ("srg_classifier.bias", (C,))]
                         + _tree.explainer_extra_shapes(config, True))
        self.surrogate_null = nn.Parameter(torch.zeros((1, config.num_labels)), requires_grad=False)

    def forward(self, x: Tensor, attention_mask: MaskLike, surrogate_grand: Optional[Tensor] = None,
                surrogate_null: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        words = pack_token_mask(attention_mask, x.shape[0], engine.n_players_of(self.config))
        return self._engine(engine.FroyoFinalEngine).final(x, words)

    def train(self, mode: bool = True):
        super().train(mode)
        _tree.freeze_model_parameters(self, "vit")
        _tree.freeze_model_parameters(self, "classifier")
        return self
