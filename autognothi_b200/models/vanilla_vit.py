"""Drop-in for reference models/vanilla_vit.py: same class names, constructor/forward signatures and
state-dict keys; the forward passes run on the sm_100a kernels (autognothi_b200/engine.py).

`model.agb_precision` selects the arithmetic: "bf16" (default; tcgen05 tensor cores) or "fp32" (exact
CUDA-core mode used for the rtol-1e-4 parity gate).  `attention_mask` is the reference's (N, T) {0,1}
token mask (CLS column included) or a `PackedMasks`.
"""
from __future__ import annotations

from typing import Optional, Tuple

import pydantic
import torch
from torch import Tensor, nn

from .. import engine, ops
from . import _tree
from .shapley import MaskLike, PackedMasks


class VanillaViTConfig(pydantic.BaseModel):
    """reference models/vanilla_vit.py:14-32 (identical fields)"""

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


def pack_token_mask(attention_mask: MaskLike, rows: int, n_players: int) -> Tensor:
    """(rows, T) int {0,1} token mask (column 0 = CLS) or PackedMasks -> packed words (rows, W)."""
    if isinstance(attention_mask, PackedMasks):
        assert attention_mask.rows == rows and attention_mask.n_players == n_players
        return attention_mask.words
    m = attention_mask
    assert m.is_cuda, "attention_mask must be a CUDA tensor (no CPU path)"
    assert m.shape == (rows, n_players + 1), f"attention_mask must be ({rows}, {n_players + 1}), got {tuple(m.shape)}"
    return ops.pack_masks(m.to(torch.int64), prepend_cls=False)


class _EngineModule(nn.Module):
    """Shared plumbing: packed-weight cache keyed on parameter versions + precision switch."""

    agb_precision: str = "bf16"

    def _engine(self, kind, kind_key=None):
        """kind: engine class or factory (sd, config, precision) -> engine; kind_key names it in the cache when `kind` is a
        fresh closure (the LTT modules), so that a module asking for two kinds never gets the wrong one back."""
        sig = (self.agb_precision, _tree.state_signature(self))
        cache = self.__dict__.setdefault("_agb_cache", {})
        if cache.get("sig") != sig:
            cache.clear()
            cache["sig"] = sig
        key = ("eng", kind_key if kind_key is not None else getattr(kind, "__qualname__", repr(kind)))
        if key not in cache:
            sd = {k: v for k, v in self.state_dict().items()}
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("autognothi_b200 models run on CUDA only: call .to('cuda') first (no CPU fallback)")
            cache[key] = kind(sd, self.config, self.agb_precision)
        return cache[key]


class VanillaViTClassifier(_EngineModule):
    """reference models/vanilla_vit.py:35-58 — ViT + Linear + Softmax (outputs are probabilities)."""

    def __init__(self, config: VanillaViTConfig):
        super().__init__()
        self.config = config
        H, C = config.hidden_size, config.num_labels
        _tree.build_tree(self, _tree.vit_backbone_shapes(config) + [("classifier.weight", (C, H)), ("classifier.bias", (C,))])

    def train(self, mode: bool = True):
        super().train(mode)
        _tree.freeze_model_parameters(self, "vit")
        _tree.freeze_model_parameters(self, "classifier")
        return self

    def forward(self, x: Tensor, attention_mask: MaskLike, n_mask_samples: int = 1) -> Tensor:
        """x (B,C,px,px); attention_mask (B*S, T) token mask or PackedMasks with B*S rows -> (B*S, num_labels)
        probabilities in row order b*S+s.  S = n_mask_samples = 1 is the reference-shaped call."""
        n = engine.n_players_of(self.config)
        rows = x.shape[0] * n_mask_samples
        words = pack_token_mask(attention_mask, rows, n)
        return self._engine(engine.SurrogateEngine).probs(x, words, n_mask_samples)


class VanillaViTSurrogate(VanillaViTClassifier):
    """reference models/vanilla_vit.py:61-66"""

    def train(self, mode: bool = True):
        nn.Module.train(self, mode)
        return self

    def forward(self, x: Tensor, attention_mask: MaskLike, n_mask_samples: int = 1) -> Tensor:
        # surrogate training (reference scripts/train_surrogate.py:131-150): differentiable w.r.t. the parameters
        if n_mask_samples == 1 and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .. import training
            words = pack_token_mask(attention_mask, x.shape[0], engine.n_players_of(self.config))
            return training.surrogate_forward_train(self, x, words)
        return super().forward(x, attention_mask, n_mask_samples)


class VanillaViTExplainer(_EngineModule):
    """reference models/vanilla_vit.py:69-132"""

    def __init__(self, config: VanillaViTConfig):
        super().__init__()
        self.config = config
        _tree.build_tree(self, _tree.vit_backbone_shapes(config) + _tree.explainer_extra_shapes(config, True))

    def forward(self, pixel_values: Tensor, attention_mask: MaskLike, surrogate_grand: Optional[Tensor],
                surrogate_null: Optional[Tensor]) -> Tensor:
        """-> (B, num_labels, n_players) attributions (normalised over all T tokens, CLS dropped)."""
        n = engine.n_players_of(self.config)
        words = pack_token_mask(attention_mask, pixel_values.shape[0], n)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .. import training
            return training.explainer_forward_train(self, pixel_values, words, surrogate_grand, surrogate_null)
        return self._engine(engine.ExplainerEngine).phi(pixel_values, words, surrogate_grand, surrogate_null)


class VanillaViTFinal(nn.Module):
    """reference models/vanilla_vit.py:135-182"""

    def __init__(self, config: VanillaViTConfig):
        super().__init__()
        self.config = config
        self.classifier = VanillaViTClassifier(config)
        self.surrogate = VanillaViTSurrogate(config)
        self.surrogate_null = nn.Parameter(torch.zeros((1, config.num_labels)), requires_grad=False)
        self.explainer = VanillaViTExplainer(config)

    def forward(self, pixel_values: Tensor, attention_mask: MaskLike) -> Tuple[Tensor, Tensor]:
        logits = self.classifier(pixel_values, attention_mask)
        grand = self.surrogate(pixel_values, attention_mask) if self.config.explainer_normalize else None
        phi = self.explainer(pixel_values, attention_mask, grand, self.surrogate_null)
        return logits, phi

    def train(self, mode: bool = True):
        super().train(mode)
        _tree.freeze_model_parameters(self, "classifier")
        return self
