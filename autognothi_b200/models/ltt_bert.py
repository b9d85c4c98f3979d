"""Drop-in for reference models/ltt_bert.py (ladder side tuning over a frozen BERT); see models/ltt_vit.py."""
from __future__ import annotations

from typing import Optional, Tuple

import pydantic
import torch
from torch import Tensor, nn

from .. import engine
from . import _tree
from .ltt_vit import _LttModule
from .shapley import MaskLike
from .vanilla_bert import VanillaBertConfig, _check_token_types
from .vanilla_vit import pack_token_mask


class LttBertConfig(pydantic.BaseModel):
    """reference models/ltt_bert.py:20-64 (identical fields)"""

    attention_probs_dropout_prob: float
    explainer_s_attn_num_layers: int  # side head
    explainer_s_head_hidden_size: int  # side head
    explainer_normalize: bool  # side head
    hidden_dropout_prob: float
    hidden_size: int
    intermediate_size: int
    layer_norm_eps: float
    max_position_embeddings: int
    num_attention_heads: int
    num_hidden_layers: int
    num_labels: int
    pad_token_id: int
    s_attn_hidden_size: int  # side attention
    s_attn_intermediate_size: int  # side attention
    type_vocab_size: int
    vocab_size: int

    @property
    def is_decoder(self) -> bool:
        return False

    def into(self) -> VanillaBertConfig:
        return VanillaBertConfig(
            attention_probs_dropout_prob=self.attention_probs_dropout_prob,
            explainer_attn_num_layers=self.explainer_s_attn_num_layers,
            explainer_head_hidden_size=self.explainer_s_head_hidden_size,
            explainer_normalize=self.explainer_normalize,
            hidden_dropout_prob=self.hidden_dropout_prob, hidden_size=self.hidden_size,
            intermediate_size=self.intermediate_size, layer_norm_eps=self.layer_norm_eps,
            max_position_embeddings=self.max_position_embeddings, num_attention_heads=self.num_attention_heads,
            num_hidden_layers=self.num_hidden_layers, num_labels=self.num_labels, pad_token_id=self.pad_token_id,
            type_vocab_size=self.type_vocab_size, vocab_size=self.vocab_size)


_FROZEN = ("bert.embeddings", "bert.encoder.layers", "bert_pooler", "classifier")


def _position_ids(module: nn.Module, config) -> None:
    # non-persistent buffer of the reference's embeddings (models/vanilla_bert.py:297-301) — not in the state dict
    module.bert.embeddings.register_buffer(
        "position_ids", torch.arange(config.max_position_embeddings).expand((1, -1)), persistent=False)


class LttBertSurrogate(_LttModule):
    """reference models/ltt_bert.py:66-122 — returns (side-ladder probabilities, backbone probabilities)"""

    _frozen = _FROZEN

    def __init__(self, config: LttBertConfig):
        super().__init__()
        self.config = config
        _tree.build_tree(self, _tree.ltt_shapes(config, False, "surrogate"))
        _position_ids(self, config)

    def forward(self, input_ids: Tensor, attention_mask: MaskLike, token_type_ids: Optional[Tensor] = None,
                n_mask_samples: int = 1) -> Tuple[Tensor, Tensor]:
        _check_token_types(token_type_ids)
        rows = input_ids.shape[0] * n_mask_samples
        words = pack_token_mask(attention_mask, rows, engine.n_players_of(self.config))
        if n_mask_samples == 1 and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .. import training
            return training.ltt_surrogate_forward_train(self, input_ids, words)
        return self._ltt().surrogate(input_ids, words, n_mask_samples)


class LttBertExplainer(_LttModule):
    """reference models/ltt_bert.py:125-222 — returns (phi (B, C, n), backbone probabilities)"""

    _kind = "explainer"
    _frozen = _FROZEN

    def __init__(self, config: LttBertConfig):
        super().__init__()
        self.config = config
        _tree.build_tree(self, _tree.ltt_shapes(config, False, "explainer"))
        _position_ids(self, config)

    def forward(self, input_ids: Tensor, attention_mask: MaskLike, token_type_ids: Optional[Tensor],
                surrogate_grand: Optional[Tensor], surrogate_null: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
        _check_token_types(token_type_ids)
        words = pack_token_mask(attention_mask, input_ids.shape[0], engine.n_players_of(self.config))
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .. import training
            return training.ltt_explainer_forward_train(self, input_ids, words, surrogate_grand, surrogate_null)
        return self._ltt().explainer(input_ids, words, surrogate_grand, surrogate_null)


class LttBertFinal(_LttModule):
    """reference models/ltt_bert.py:225-344 — returns (backbone probabilities, phi)"""

    _kind = "final"
    _frozen = _FROZEN

    def __init__(self, config: LttBertConfig):
        super().__init__()
        self.config = config
        _tree.build_tree(self, _tree.ltt_shapes(config, False, "final"))
        self.surrogate_null = nn.Parameter(torch.zeros((1, config.num_labels)), requires_grad=False)
        _position_ids(self, config)

    def forward(self, input_ids: Tensor, attention_mask: MaskLike, token_type_ids: Optional[Tensor] = None
                ) -> Tuple[Tensor, Tensor]:
        _check_token_types(token_type_ids)
        words = pack_token_mask(attention_mask, input_ids.shape[0], engine.n_players_of(self.config))
        return self._ltt().final(input_ids, words)
