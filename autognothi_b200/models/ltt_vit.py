"""Drop-in for reference models/ltt_vit.py — "ladder side tuning", the method of the paper: the classifier backbone is
frozen and the surrogate / explainer are narrow side ladders fed by every backbone block
(s <- block_i^side(s + GELU(W_i h_i)), reference l.420-436).  Same class names, forward signatures, return order and
state-dict keys.  Inference runs on autognothi_b200.engine.LttEngine: the backbone on the LayerNorm-folded tcgen05
path, the ladders (hidden size s_attn_hidden_size, head dim s_attn_hidden_size / num_attention_heads) on the generic
kernels; the bundle evaluates the backbone once for both ladders.
"""
from __future__ import annotations

from typing import Optional, Tuple

import pydantic
import torch
from torch import Tensor, nn

from .. import engine
from . import _tree
from .shapley import MaskLike
from .vanilla_vit import VanillaViTConfig, _EngineModule, pack_token_mask


class LttViTConfig(pydantic.BaseModel):
    """reference models/ltt_vit.py:14-52 (identical fields)"""

    attention_probs_dropout_prob: float
    explainer_s_attn_num_layers: int  # side head
    explainer_s_head_hidden_size: int  # side head
    explainer_normalize: bool  # side head
    hidden_dropout_prob: float
    hidden_size: int
    intermediate_size: int
    layer_norm_eps: float
    num_attention_heads: int
    num_hidden_layers: int
    num_labels: int
    s_attn_hidden_size: int  # side attention
    s_attn_intermediate_size: int  # side attention
    img_channels: int
    img_px_size: int
    img_patch_size: int

    def into(self) -> VanillaViTConfig:
        return VanillaViTConfig(
            attention_probs_dropout_prob=self.attention_probs_dropout_prob,
            explainer_attn_num_layers=self.explainer_s_attn_num_layers,
            explainer_head_hidden_size=self.explainer_s_head_hidden_size,
            explainer_normalize=self.explainer_normalize,
            hidden_dropout_prob=self.hidden_dropout_prob, hidden_size=self.hidden_size,
            intermediate_size=self.intermediate_size, layer_norm_eps=self.layer_norm_eps,
            num_attention_heads=self.num_attention_heads, num_hidden_layers=self.num_hidden_layers,
            num_labels=self.num_labels, img_channels=self.img_channels, img_px_size=self.img_px_size,
            img_patch_size=self.img_patch_size)


_FROZEN = ("vit.embeddings", "vit.encoder.layers", "vit.layernorm", "classifier")


class _LttModule(_EngineModule):
    _kind = "surrogate"
    _frozen = _FROZEN

    def train(self, mode: bool = True):
        super().train(mode)
        for prefix in self._frozen:
            _tree.freeze_model_parameters(self, prefix)
        return self

    def ltt_freeze_layers_until(self, layer_id: int) -> None:
        """reference models/ltt_vit.py:402-405: blocks >= layer_id no longer feed the ladders"""
        self._ltt_freeze_layer = max(1, min(self.config.num_hidden_layers, int(layer_id)))

    def _ltt(self) -> engine.LttEngine:
        kind = self._kind
        eng = self._engine(lambda sd, cfg, prec: engine.LttEngine(sd, cfg, prec, kind), kind_key=f"ltt:{kind}")
        eng.freeze_layer = getattr(self, "_ltt_freeze_layer", None)
        return eng


class LttViTSurrogate(_LttModule):
    """reference models/ltt_vit.py:55-94 — returns (side-ladder probabilities, backbone probabilities)"""

    def __init__(self, config: LttViTConfig):
        super().__init__()
        self.config = config
        _tree.build_tree(self, _tree.ltt_shapes(config, True, "surrogate"))

    def forward(self, pixel_values: Tensor, attention_mask: MaskLike, n_mask_samples: int = 1) -> Tuple[Tensor, Tensor]:
        rows = pixel_values.shape[0] * n_mask_samples
        words = pack_token_mask(attention_mask, rows, engine.n_players_of(self.config))
        if n_mask_samples == 1 and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .. import training
            return training.ltt_surrogate_forward_train(self, pixel_values, words)
        return self._ltt().surrogate(pixel_values, words, n_mask_samples)


class LttViTExplainer(_LttModule):
    """reference models/ltt_vit.py:97-183 — returns (phi (B, C, n), backbone probabilities)"""

    _kind = "explainer"

    def __init__(self, config: LttViTConfig):
        super().__init__()
        self.config = config
        _tree.build_tree(self, _tree.ltt_shapes(config, True, "explainer"))

    def forward(self, pixel_values: Tensor, attention_mask: MaskLike, surrogate_grand: Optional[Tensor],
                surrogate_null: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
        words = pack_token_mask(attention_mask, pixel_values.shape[0], engine.n_players_of(self.config))
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .. import training
            return training.ltt_explainer_forward_train(self, pixel_values, words, surrogate_grand, surrogate_null)
        return self._ltt().explainer(pixel_values, words, surrogate_grand, surrogate_null)


class LttViTFinal(_LttModule):
    """reference models/ltt_vit.py:186-287 — returns (backbone probabilities, phi); everything frozen in train()"""

    _kind = "final"
    _frozen = (...,)

    def __init__(self, config: LttViTConfig):
        super().__init__()
        self.config = config
        _tree.build_tree(self, _tree.ltt_shapes(config, True, "final"))
        self.surrogate_null = nn.Parameter(torch.zeros((1, config.num_labels)), requires_grad=False)

    def forward(self, pixel_values: Tensor, attention_mask: MaskLike) -> Tuple[Tensor, Tensor]:
        words = pack_token_mask(attention_mask, pixel_values.shape[0], engine.n_players_of(self.config))
        return self._ltt().final(pixel_values, words)
