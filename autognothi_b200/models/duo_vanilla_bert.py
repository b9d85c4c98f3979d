"""Drop-in for reference models/duo_vanilla_bert.py (dual-objective explainer over BERT); see models/duo_vanilla_vit.py.
Quirk kept: this explainer returns (RAW class logits, phi) — logits first and without the softmax the ViT one applies
(reference models/duo_vanilla_bert.py:120-125,148)."""
from __future__ import annotations

from typing import Optional, Tuple

import pydantic
import torch
from torch import Tensor, nn

from .. import engine
from . import _tree
from .shapley import MaskLike
from .vanilla_bert import (VanillaBertClassifier, VanillaBertConfig, VanillaBertSurrogate, _check_token_types)
from .vanilla_vit import _EngineModule, pack_token_mask


class DuoVanillaBertConfig(pydantic.BaseModel):
    """reference models/duo_vanilla_bert.py:20-62 (identical fields)"""

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

    def into(self) -> VanillaBertConfig:
        return VanillaBertConfig(**self.model_dump())


class DuoVanillaBertClassifier(VanillaBertClassifier):
    """reference models/duo_vanilla_bert.py:65-69"""

    def __init__(self, config: DuoVanillaBertConfig):
        super().__init__(config.into())


class DuoVanillaBertSurrogate(VanillaBertSurrogate):
    """reference models/duo_vanilla_bert.py:72-76"""

    def __init__(self, config: DuoVanillaBertConfig):
        super().__init__(config.into())


class DuoVanillaBertExplainer(_EngineModule):
    """reference models/duo_vanilla_bert.py:79-150 — returns (raw class logits (B, C), phi (B, C, n))"""

    def __init__(self, config: DuoVanillaBertConfig):
        super().__init__()
        self.config = config
        H, C = config.hidden_size, config.num_labels
        _tree.build_tree(self, _tree.bert_backbone_shapes(config) + [
            ("bert_pooler.dense.weight", (H, H)), ("bert_pooler.dense.bias", (H,)),
            ("classifier.weight", (C, H)), ("classifier.bias", (C,))] + _tree.explainer_extra_shapes(config, False))
        self.bert.embeddings.register_buffer(
            "position_ids", torch.arange(config.max_position_embeddings).expand((1, -1)), persistent=False)

    def forward(self, input_ids: Tensor, attention_mask: MaskLike, token_type_ids: Optional[Tensor],
                surrogate_grand: Optional[Tensor], surrogate_null: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
        _check_token_types(token_type_ids)
        words = pack_token_mask(attention_mask, input_ids.shape[0], engine.n_players_of(self.config))
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .. import training
            phi, logits = training.duo_explainer_forward_train(self, input_ids, words, surrogate_grand, surrogate_null)
        else:
            phi, logits = self._engine(engine.DuoExplainerEngine).duo(input_ids, words, surrogate_grand, surrogate_null)
        return logits, phi


class DuoVanillaBertFinal(nn.Module):
    """reference models/duo_vanilla_bert.py:153-213"""

    def __init__(self, config: DuoVanillaBertConfig):
        super().__init__()
        self.config = config
        self.surrogate = VanillaBertSurrogate(config.into())
        self.surrogate_null = nn.Parameter(torch.zeros((1, config.num_labels)), requires_grad=False)
        self.explainer = DuoVanillaBertExplainer(config)

    def forward(self, input_ids: Tensor, attention_mask: MaskLike, token_type_ids: Optional[Tensor] = None
                ) -> Tuple[Tensor, Tensor]:
        grand = self.surrogate(input_ids, attention_mask, token_type_ids) if self.config.explainer_normalize else None
        logits, phi = self.explainer(input_ids, attention_mask, token_type_ids, grand, self.surrogate_null)
        return logits, phi
