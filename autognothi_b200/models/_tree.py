"""Parameter containers that reproduce the reference's state-dict ABI.

The reference's checkpoints are `state_dict()`s of deeply nested nn.Modules; a drop-in must load them
with `load_state_dict(strict=True)`.  Only the *names and shapes* are ABI — the arithmetic is done by
the CUDA engine — so the tree is built from a (key -> shape) table instead of mirroring the
reference's module classes.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Iterator, Sequence, Tuple

import torch
from torch import nn


class ParamNode(nn.Module):
    """A named container; numeric children make it behave like an nn.ModuleList."""

    def __getitem__(self, idx: int) -> nn.Module:
        return self._modules[str(idx)]

    def __len__(self) -> int:
        return len(self._modules)

    def __iter__(self) -> Iterator[nn.Module]:
        return iter(self._modules.values())


def _init(name: str, shape: Sequence[int]) -> torch.Tensor:
    """Same families of initial distributions as the reference's constructors (nn.Linear / nn.Conv2d
    kaiming-uniform, nn.LayerNorm ones/zeros, nn.Embedding N(0,1), randn cls/pos tokens)."""
    leaf = name.rsplit(".", 1)[-1]
    lname = name.lower()
    if "layernorm" in lname or (name.startswith(("explainer_mlp.0.", "s_explainer_mlp.0.")) and len(shape) == 1):
        return torch.ones(shape) if leaf == "weight" else torch.zeros(shape)
    if leaf in ("cls_token", "position_embeddings") or "embeddings.weight" in name or name.endswith("_embeddings.weight"):
        return torch.randn(shape)
    fan_in = int(torch.tensor(shape[1:]).prod()) if len(shape) > 1 else None
    if leaf == "bias":
        return None  # filled after the matching weight (needs fan_in)
    bound = 1.0 / math.sqrt(fan_in)
    return torch.empty(shape).uniform_(-bound, bound)


def build_tree(root: nn.Module, shapes: Iterable[Tuple[str, Tuple[int, ...]]]) -> None:
    shapes = list(shapes)
    fan_in: Dict[str, int] = {}
    for name, shape in shapes:
        if name.endswith(".weight") and len(shape) > 1:
            fan_in[name[: -len(".weight")]] = int(torch.tensor(shape[1:]).prod())
    for name, shape in shapes:
        parts = name.split(".")
        node = root
        for p in parts[:-1]:
            if p not in node._modules:
                node.add_module(p, ParamNode())
            node = node._modules[p]
        value = _init(name, shape)
        if value is None:
            base = name[: -len(".bias")]
            bound = 1.0 / math.sqrt(fan_in[base]) if base in fan_in else 0.0
            value = torch.empty(shape).uniform_(-bound, bound) if bound > 0 else torch.zeros(shape)
        node.register_parameter(parts[-1], nn.Parameter(value))


def freeze_model_parameters(on: nn.Module, *item_names, requires_grad: bool = False) -> None:
    """reference utils/nnmodel.py:48-60"""
    if len(item_names) == 1 and item_names[0] is ...:
        for param in on.parameters():
            param.requires_grad = requires_grad
    else:
        for name, param in on.named_parameters():
            if any(name.startswith(f"{n}.") for n in item_names):
                param.requires_grad = requires_grad


def state_signature(module: nn.Module) -> Tuple:
    """Cheap change detector for the packed-weight cache: (data_ptr, version) of every tensor."""
    return tuple((p.data_ptr(), p._version) for p in list(module.parameters()) + list(module.buffers()))


# ------------------------------------------------------------------------------------------------
# key tables (mirrors of oracle/synth.py, kept separate on purpose: the product never imports oracle/)
# ------------------------------------------------------------------------------------------------
def _layer(prefix: str, H: int, I: int, vit: bool, ln1: bool = True, ln2: bool = True):
    out = []
    for nm in ("query", "key", "value"):
        out += [(f"{prefix}.attention.self.{nm}.weight", (H, H)), (f"{prefix}.attention.self.{nm}.bias", (H,))]
    out += [(f"{prefix}.attention.output.dense.weight", (H, H)), (f"{prefix}.attention.output.dense.bias", (H,))]
    if not vit and ln1:
        out += [(f"{prefix}.attention.output.LayerNorm.weight", (H,)), (f"{prefix}.attention.output.LayerNorm.bias", (H,))]
    out += [(f"{prefix}.intermediate.dense.weight", (I, H)), (f"{prefix}.intermediate.dense.bias", (I,))]
    out += [(f"{prefix}.output.dense.weight", (H, I)), (f"{prefix}.output.dense.bias", (H,))]
    if vit:
        if ln1:
            out += [(f"{prefix}.layernorm_before.weight", (H,)), (f"{prefix}.layernorm_before.bias", (H,))]
        if ln2:
            out += [(f"{prefix}.layernorm_after.weight", (H,)), (f"{prefix}.layernorm_after.bias", (H,))]
    elif ln2:
        out += [(f"{prefix}.output.LayerNorm.weight", (H,)), (f"{prefix}.output.LayerNorm.bias", (H,))]
    return out


def vit_backbone_shapes(cfg):
    H, I = cfg.hidden_size, cfg.intermediate_size
    T = (cfg.img_px_size // cfg.img_patch_size) ** 2 + 1
    P = cfg.img_patch_size
    out = [("vit.embeddings.cls_token", (1, 1, H)), ("vit.embeddings.position_embeddings", (1, T, H)),
           ("vit.embeddings.patch_embeddings.projection.weight", (H, cfg.img_channels, P, P)),
           ("vit.embeddings.patch_embeddings.projection.bias", (H,))]
    for i in range(cfg.num_hidden_layers):
        out += _layer(f"vit.encoder.layers.{i}", H, I, True)
    out += [("vit.layernorm.weight", (H,)), ("vit.layernorm.bias", (H,))]
    return out


def bert_backbone_shapes(cfg):
    H, I = cfg.hidden_size, cfg.intermediate_size
    out = [("bert.embeddings.word_embeddings.weight", (cfg.vocab_size, H)),
           ("bert.embeddings.position_embeddings.weight", (cfg.max_position_embeddings, H)),
           ("bert.embeddings.token_type_embeddings.weight", (cfg.type_vocab_size, H)),
           ("bert.embeddings.LayerNorm.weight", (H,)), ("bert.embeddings.LayerNorm.bias", (H,))]
    for i in range(cfg.num_hidden_layers):
        out += _layer(f"bert.encoder.layers.{i}", H, I, False)
    return out


def explainer_extra_shapes(cfg, vit: bool):
    H, I, E, C = cfg.hidden_size, cfg.intermediate_size, int(cfg.explainer_head_hidden_size), cfg.num_labels
    out = []
    for i in range(cfg.explainer_attn_num_layers):
        out += _layer(f"explainer_attn.{i}", H, I, vit, ln1=(i != 0), ln2=True)
    if vit:
        out += [("explainer_mlp.0.weight", (H,)), ("explainer_mlp.0.bias", (H,)),
                ("explainer_mlp.1.weight", (E, H)), ("explainer_mlp.1.bias", (E,)),
                ("explainer_mlp.3.weight", (E, E)), ("explainer_mlp.3.bias", (E,)),
                ("explainer_mlp.5.weight", (C, E)), ("explainer_mlp.5.bias", (C,))]
    else:
        out += [("explainer_mlp.0.weight", (E, H)), ("explainer_mlp.0.bias", (E,)),
                ("explainer_mlp.2.weight", (E, E)), ("explainer_mlp.2.bias", (E,)),
                ("explainer_mlp.4.weight", (C, E)), ("explainer_mlp.4.bias", (C,))]
    return out


def ltt_shapes(cfg, vit: bool, kind: str):
    """Key/shape table of the LTT classes (reference models/ltt_vit.py:55-340, models/ltt_bert.py:66-400).
    kind: "surrogate" (ladder 0 + side classifier), "explainer" (ladder 0 + side explainer), "final" (ladders 0, 1 + both)."""
    H, C = cfg.hidden_size, cfg.num_labels
    Hs, Is, E = cfg.s_attn_hidden_size, cfg.s_attn_intermediate_size, int(cfg.explainer_s_head_hidden_size)
    root = "vit" if vit else "bert"
    out = (vit_backbone_shapes(cfg) if vit else bert_backbone_shapes(cfg))
    for b in range(2 if kind == "final" else 1):
        for i in range(cfg.num_hidden_layers):
            out += [(f"{root}.encoder.s_attn_maps.{b}_{i}.weight", (Hs, H)), (f"{root}.encoder.s_attn_maps.{b}_{i}.bias", (Hs,))]
            out += _layer(f"{root}.encoder.s_attn_layers.{b}_{i}", Hs, Is, vit)
        if vit:
            out += [(f"vit.s_attn_layernorm.{b}.weight", (Hs,)), (f"vit.s_attn_layernorm.{b}.bias", (Hs,))]
    if not vit:
        out += [("bert_pooler.dense.weight", (H, H)), ("bert_pooler.dense.bias", (H,))]
    out += [("classifier.weight", (C, H)), ("classifier.bias", (C,))]
    if kind in ("surrogate", "final"):
        if not vit:
            out += [("bert_s_attn_pooler.dense.weight", (Hs, Hs)), ("bert_s_attn_pooler.dense.bias", (Hs,))]
        out += [("s_attn_classifier.weight", (C, Hs)), ("s_attn_classifier.bias", (C,))]
    if kind in ("explainer", "final"):
        attn = "s_explainer_attn" if vit else "s_attn_attention_layers"
        mlp = "s_explainer_mlp" if vit else "s_attn_explainer"
        for i in range(cfg.explainer_s_attn_num_layers):
            out += _layer(f"{attn}.{i}", Hs, Is, vit, ln1=(i != 0), ln2=True)
        if vit:
            out += [(f"{mlp}.0.weight", (Hs,)), (f"{mlp}.0.bias", (Hs,)),
                    (f"{mlp}.1.weight", (E, Hs)), (f"{mlp}.1.bias", (E,)),
                    (f"{mlp}.3.weight", (E, E)), (f"{mlp}.3.bias", (E,)),
                    (f"{mlp}.5.weight", (C, E)), (f"{mlp}.5.bias", (C,))]
        else:
            out += [(f"{mlp}.0.weight", (E, Hs)), (f"{mlp}.0.bias", (E,)),
                    (f"{mlp}.2.weight", (E, E)), (f"{mlp}.2.bias", (E,)),
                    (f"{mlp}.4.weight", (C, E)), (f"{mlp}.4.bias", (C,))]
    return out
