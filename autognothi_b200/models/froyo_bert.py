"""Drop-in for reference models/froyo_bert.py (frozen BERT backbone, trainable heads); see models/froyo_vit.py."""
from __future__ import annotations

from typing import Optional, Tuple

import pydantic
import torch
from torch import Tensor, nn

from .. import engine
from . import _tree
from .shapley import MaskLike
from .vanilla_bert import (VanillaBertClassifier, VanillaBertConfig, VanillaBertExplainer, VanillaBertSurrogate,
                           _check_token_types)
from .vanilla_vit import _EngineModule, pack_token_mask


class FroyoBertConfig(pydantic.BaseModel):
    """reference models/froyo_bert.py:21-66 (identical fields)"""

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


class FroyoBertClassifier(VanillaBertClassifier):
    """reference models/froyo_bert.py:69-80"""

    def __init__(self, config: FroyoBertConfig):
        super().__init__(config.into())

    def train(self, mode: bool = True):
        nn.Module.train(self, mode)
        _tree.freeze_model_parameters(self, "bert")
        _tree.freeze_model_parameters(self, "bert_pooler")
        _tree.freeze_model_parameters(self, "classifier")
        return self


class FroyoBertSurrogate(VanillaBertSurrogate):
    """reference models/froyo_bert.py:83-91 — `bert_pooler.*` and `classifier.*` train"""

    def __init__(self, config: FroyoBertConfig):
        super().__init__(config.into())

    def train(self, mode: bool = True):
        nn.Module.train(self, mode)
        _tree.freeze_model_parameters(self, "bert")
        return self


class FroyoBertExplainer(VanillaBertExplainer):
    """reference models/froyo_bert.py:94-102"""

    def __init__(self, config: FroyoBertConfig):
        super().__init__(config.into())

    def train(self, mode: bool = True):
        nn.Module.train(self, mode)
        _tree.freeze_model_parameters(self, "bert")
        return self


class FroyoBertFinal(_EngineModule):
    """reference models/froyo_bert.py:105-213: one backbone pass, two pooled heads, the explainer tail."""

    def __init__(self, _config: FroyoBertConfig):
        super().__init__()
        config = _config.into()
        self.config = config
        H, C = config.hidden_size, config.num_labels
        _tree.build_tree(self, _tree.bert_backbone_shapes(config) + [
            ("bert_pooler.dense.weight", (H, H)), ("bert_pooler.dense.bias", (H,)),
            ("classifier.weight", (C, H)), ("classifier.bias", (C,)),
            ("srg_bert_pooler.dense.weight", (H, H)), ("srg_bert_pooler.dense.bias", (H,)),
            ("srg_classifier.weight", (C, H)), ("srg_classifier.bias", (C,))]
            + _tree.explainer_extra_shapes(config, False))
        self.surrogate_null = nn.Parameter(torch.zeros((1, config.num_labels)), requires_grad=False)
        self.bert.embeddings.register_buffer(
            "position_ids", torch.arange(config.max_position_embeddings).expand((1, -1)), persistent=False)

    def forward(self, input_ids: Tensor, attention_mask: MaskLike, token_type_ids: Optional[Tensor] = None
                ) -> Tuple[Tensor, Tensor]:
        _check_token_types(token_type_ids)
        words = pack_token_mask(attention_mask, input_ids.shape[0], engine.n_players_of(self.config))
        return self._engine(engine.FroyoFinalEngine).final(input_ids, words)

    def train(self, mode: bool = True):
        super().train(mode)
        _tree.freeze_model_parameters(self, "bert")
        _tree.freeze_model_parameters(self, "bert_pooler")
        _tree.freeze_model_parameters(self, "classifier")
        return self
