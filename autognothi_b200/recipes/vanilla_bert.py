"""Drop-in for reference recipes/vanilla_bert.py (the ModelRecipe of the vanilla BERT pipeline)."""
from __future__ import annotations

import dataclasses
import re
from typing import Any, Callable, List, Optional, Tuple

import torch
from torch import Tensor, nn

from ..models.shapley import MaskLike, PackedMasks
from ..models.vanilla_bert import (VanillaBertClassifier, VanillaBertConfig, VanillaBertExplainer, VanillaBertFinal,
                                   VanillaBertSurrogate)
from ._common import copy_matching, resolve_masks
from .types import ModelRecipe, ModelRecipe_Measurements, ModelRecipe_Training


@dataclasses.dataclass
class VanillaBertMisc:
    tokenizer: Any = None


def _n_players(cfg: VanillaBertConfig) -> int:
    return cfg.max_position_embeddings - 1  # reference recipes/vanilla_bert.py:55


def vanilla_bert_recipe() -> ModelRecipe:
    return ModelRecipe(
        id="vanilla_bert",
        version="beta.1.01",
        t_config=VanillaBertConfig,
        t_classifier=VanillaBertClassifier,
        t_surrogate=VanillaBertSurrogate,
        t_explainer=VanillaBertExplainer,
        t_final=VanillaBertFinal,
        load_misc=_load_misc,
        conv_pretrained_classifier=_conv_pretrained_classifier,
        conv_classifier_surrogate=_conv_classifier_surrogate,
        conv_surrogate_explainer=_conv_surrogate_explainer,
        conv_explainer_final=_conv_explainer_final,
        n_players=_n_players,
        gen_input=lambda cfg, misc, device: _gen_input(cfg.max_position_embeddings, misc.tokenizer, device),
        gen_null=lambda cfg, misc, device: _gen_null(cfg.max_position_embeddings, misc.tokenizer, device),
        training=ModelRecipe_Training(True, True, True, False, False),
        fw_classifier=_fw_classifier,
        fw_surrogate=_fw_surrogate,
        fw_explainer=_fw_explainer,
        fw_final=_fw_final,
        measurements=ModelRecipe_Measurements(True, True, True, True, True, True, True, True, False, True),
    )


def _load_misc(m_path, cfg) -> VanillaBertMisc:
    """reference recipes/vanilla_bert.py:92-96 loads the HF tokenizer stored next to the base model."""
    from transformers import AutoTokenizer  # host-side text preprocessing only
    return VanillaBertMisc(tokenizer=AutoTokenizer.from_pretrained(m_path))


_HF_RULES = [  # HF BertForSequenceClassification -> ours (reference recipes/vanilla_bert.py:105-131)
    (r"^bert\.encoder\.layer\.(\d+)\.(.+)$", r"bert.encoder.layers.\1.\2"),
    (r"^bert\.pooler\.dense\.(weight|bias)$", r"bert_pooler.dense.\1"),
]


def pre_conv_bert(cfg: VanillaBertConfig, model: Any) -> VanillaBertClassifier:
    classifier = VanillaBertClassifier(cfg)
    sd = model.state_dict() if isinstance(model, nn.Module) else dict(model)
    if any(k.startswith("bert.encoder.layers.") for k in sd):
        copy_matching(sd, classifier, ("bert.", "bert_pooler.", "classifier."))
        return classifier
    renamed = {}
    for k, v in sd.items():
        if k.endswith("position_ids"):
            continue
        for pat, rep in _HF_RULES:
            if re.match(pat, k):
                k = re.sub(pat, rep, k)
                break
        renamed[k] = v
    copy_matching(renamed, classifier, ("bert.", "bert_pooler.", "classifier."))
    return classifier


def _conv_pretrained_classifier(cfg, model) -> VanillaBertClassifier:
    return pre_conv_bert(cfg, model)


def _conv_classifier_surrogate(cfg, _misc, classifier) -> VanillaBertSurrogate:
    surrogate = VanillaBertSurrogate(cfg).to(next(classifier.parameters()).device)
    copy_matching(classifier.state_dict(), surrogate, ("bert.", "bert_pooler.", "classifier."))
    return surrogate


def _conv_surrogate_explainer(cfg, _misc, surrogate) -> VanillaBertExplainer:
    explainer = VanillaBertExplainer(cfg).to(next(surrogate.parameters()).device)
    copy_matching(surrogate.state_dict(), explainer, ("bert.",))
    return explainer


def _conv_explainer_final(cfg, misc, classifier, surrogate, explainer) -> VanillaBertFinal:
    device = next(classifier.parameters()).device
    n_players = _n_players(cfg)
    nil_xs = _gen_null(cfg.max_position_embeddings, misc.tokenizer, device)
    surrogate.eval()
    with torch.no_grad():
        surrogate_null, _ = _fw_surrogate(surrogate, nil_xs, PackedMasks.ones(1, n_players, device))
    final = VanillaBertFinal(cfg).to(device)
    copy_matching(classifier.state_dict(), final, ("",), "classifier.")
    copy_matching(surrogate.state_dict(), final, ("",), "surrogate.")
    copy_matching(explainer.state_dict(), final, ("",), "explainer.")
    with torch.no_grad():
        final.surrogate_null.copy_(surrogate_null)
    return final


def _gen_input(max_position_embeddings: int, tokenizer, device) -> Callable[[Any, Any], Tuple[Tensor, Tensor]]:
    """tokenise + pad to T; the tokenizer's own attention_mask is discarded, [PAD]/[SEP] are ordinary
    players (reference recipes/vanilla_bert.py:226-262)."""
    max_length = max_position_embeddings

    def mask_input(raw_xs: List[str], raw_ys: List[int]):
        rows = []
        for raw_x in raw_xs:
            enc = tokenizer(raw_x, return_tensors="pt", padding="max_length", max_length=max_length)
            rows.append(enc["input_ids"][:, :max_length])
        xs = torch.cat(rows, dim=0).to(device, non_blocking=True)
        ys = torch.tensor(raw_ys).to(device, non_blocking=True)
        return xs, ys

    return mask_input


def _gen_null(max_position_embeddings: int, tokenizer, device) -> Tensor:
    """tokenised "" padded to T (reference recipes/vanilla_bert.py:265-278).  Without a tokenizer (synthetic
    benchmarks) the BERT special ids are used directly: [CLS]=101 [SEP]=102 [PAD]=0."""
    if tokenizer is None:
        ids = torch.zeros((1, max_position_embeddings), dtype=torch.int64)
        ids[0, 0], ids[0, 1] = 101, 102
        return ids.to(device)
    enc = tokenizer("", return_tensors="pt", padding="max_length", max_length=max_position_embeddings)
    return enc["input_ids"].to(device)


def _fw_classifier(model, xs: Tensor, mask: MaskLike) -> Tuple[Tensor, Tensor]:
    pm, S = resolve_masks(xs, mask, _n_players(model.config))
    probs = model(xs, pm, None, n_mask_samples=S)
    return probs, probs


def _fw_surrogate(model, xs: Tensor, mask: MaskLike) -> Tuple[Tensor, Optional[Tensor]]:
    pm, S = resolve_masks(xs, mask, _n_players(model.config))
    return model(xs, pm, None, n_mask_samples=S), None


def _fw_explainer(model, xs: Tensor, mask: MaskLike, surrogate_grand: Tensor, surrogate_null: Tensor
                  ) -> Tuple[Tensor, Optional[Tensor]]:
    pm, S = resolve_masks(xs, mask, _n_players(model.config))
    assert S == 1, "the explainer takes one mask row per input"
    return model(xs, pm, None, surrogate_grand, surrogate_null), None


def _fw_final(model, xs: Tensor) -> Tuple[Tensor, Tensor]:
    pm = PackedMasks.ones(xs.shape[0], _n_players(model.config), xs.device)
    return model(xs, pm, None)
