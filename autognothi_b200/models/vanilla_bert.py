"""Drop-in for reference models/vanilla_bert.py: same class names, signatures and state-dict keys;
forward passes run on the sm_100a kernels.  Token masks are additive finfo.min in the reference
(HF get_extended_attention_mask, models/vanilla_bert.py:264-266,521-523) = probability exactly 0,
which the attention kernel implements by zeroing the masked K/V rows (AGB_MASK_NEGINF).
`token_type_ids` must be all zeros on this path (reference recipes/vanilla_bert.py:289)."""
from __future__ import annotations

import os

from typing import Optional, Tuple

import pydantic
import torch
from torch import Tensor, nn

from .. import engine
from . import _tree
from .shapley import MaskLike
from .vanilla_vit import _EngineModule, pack_token_mask


class VanillaBertConfig(pydantic.BaseModel):
    """reference models/vanilla_bert.py:16-39 (identical fields)"""

    attention_probs_dropout_prob: float
    explainer_attn_num_layers: int
    explainer_head_hidden_size: int
    explainer_normalize: bool
    hidden_dropout_prob: float
    hidden_size: int
    intermediate_size: int
    layer_norm_eps: float
    max_position_embeddings: int
    num_attention_heads: int
    num_hidden_layers: int
    num_labels: int
    pad_token_id: int
    type_vocab_size: int
    vocab_size: int

    @property
    def is_decoder(self) -> bool:
        return False


def _check_token_types(token_type_ids: Optional[Tensor]) -> None:
    # a device-side all-zero check would force a sync on the hot path: shape/dtype only, unless AGB_DEBUG_CHECKS=1
    if token_type_ids is not None:
        assert token_type_ids.dtype in (torch.int64, torch.int32), "token_type_ids must be an integer tensor"
        if os.environ.get("AGB_DEBUG_CHECKS") == "1" and bool((token_type_ids != 0).any()):
            raise ValueError("token_type_ids must be all zero on this path (reference recipes/vanilla_bert.py:289 always "
                             "passes zeros; only token type 0 is embedded)")


class VanillaBertClassifier(_EngineModule):
    """reference models/vanilla_bert.py:42-79 — BERT + pooler(tanh) + Linear + Softmax."""

    def __init__(self, config: VanillaBertConfig):
        super().__init__()
        self.config = config
        H, C = config.hidden_size, config.num_labels
        _tree.build_tree(self, _tree.bert_backbone_shapes(config) + [
            ("bert_pooler.dense.weight", (H, H)), ("bert_pooler.dense.bias", (H,)),
            ("classifier.weight", (C, H)), ("classifier.bias", (C,))])
        # non-persistent buffer of the reference (models/vanilla_bert.py:297-301) — not in the state dict
        self.bert.embeddings.register_buffer(
            "position_ids", torch.arange(config.max_position_embeddings).expand((1, -1)), persistent=False)

    def train(self, mode: bool = True):
        super().train(mode)
        _tree.freeze_model_parameters(self, "bert")
        _tree.freeze_model_parameters(self, "bert_pooler")
        _tree.freeze_model_parameters(self, "classifier")
        return self

    def forward(self, input_ids: Tensor, attention_mask: MaskLike, token_type_ids: Optional[Tensor] = None,
                n_mask_samples: int = 1) -> Tensor:
        _check_token_types(token_type_ids)
        n = engine.n_players_of(self.config)
        rows = input_ids.shape[0] * n_mask_samples
        words = pack_token_mask(attention_mask, rows, n)
        return self._engine(engine.SurrogateEngine).probs(input_ids, words, n_mask_samples)


class VanillaBertSurrogate(VanillaBertClassifier):
    """reference models/vanilla_bert.py:82-87"""

    def train(self, mode: bool = True):
        nn.Module.train(self, mode)
        return self

    def forward(self, input_ids: Tensor, attention_mask: MaskLike, token_type_ids: Optional[Tensor] = None,
                n_mask_samples: int = 1) -> Tensor:
        # surrogate training (reference scripts/train_surrogate.py:131-150): differentiable w.r.t. the parameters
        if n_mask_samples == 1 and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .. import training
            _check_token_types(token_type_ids)
            words = pack_token_mask(attention_mask, input_ids.shape[0], engine.n_players_of(self.config))
            return training.surrogate_forward_train(self, input_ids, words)
        return super().forward(input_ids, attention_mask, token_type_ids, n_mask_samples)


class VanillaBertExplainer(_EngineModule):
    """reference models/vanilla_bert.py:90-164"""

    def __init__(self, config: VanillaBertConfig):
        super().__init__()
        self.config = config
        _tree.build_tree(self, _tree.bert_backbone_shapes(config) + _tree.explainer_extra_shapes(config, False))
        self.bert.embeddings.register_buffer(
            "position_ids", torch.arange(config.max_position_embeddings).expand((1, -1)), persistent=False)

    def forward(self, input_ids: Tensor, attention_mask: MaskLike, token_type_ids: Optional[Tensor],
                surrogate_grand: Optional[Tensor], surrogate_null: Optional[Tensor]) -> Tensor:
        _check_token_types(token_type_ids)
        n = engine.n_players_of(self.config)
        words = pack_token_mask(attention_mask, input_ids.shape[0], n)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .. import training
            return training.explainer_forward_train(self, input_ids, words, surrogate_grand, surrogate_null)
        return self._engine(engine.ExplainerEngine).phi(input_ids, words, surrogate_grand, surrogate_null)


class VanillaBertFinal(nn.Module):
    """reference models/vanilla_bert.py:167-227"""

    def __init__(self, config: VanillaBertConfig):
        super().__init__()
        self.config = config
        self.classifier = VanillaBertClassifier(config)
        self.surrogate = VanillaBertSurrogate(config)
        self.surrogate_null = nn.Parameter(torch.zeros((1, config.num_labels)), requires_grad=False)
        self.explainer = VanillaBertExplainer(config)

    def forward(self, input_ids: Tensor, attention_mask: MaskLike, token_type_ids: Optional[Tensor] = None
                ) -> Tuple[Tensor, Tensor]:
        logits = self.classifier(input_ids, attention_mask, token_type_ids)
        grand = self.surrogate(input_ids, attention_mask, token_type_ids) if self.config.explainer_normalize else None
        phi = self.explainer(input_ids, attention_mask, token_type_ids, grand, self.surrogate_null)
        return logits, phi

    def train(self, mode: bool = True):
        super().train(mode)
        _tree.freeze_model_parameters(self, "classifier")
        return self
