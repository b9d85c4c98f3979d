"""Model configurations used by the oracle, the golden fixtures, the parity tests and bench.py.

Field names are the reference's pydantic config fields (reference models/vanilla_vit.py:14-32,
models/vanilla_bert.py:16-39, models/kernel_shap_bert.py:15-36); values of the full-size entries
are the checked-in `net.params` of reference experiments/<name>/.hparams.json.  Test infrastructure.
"""
from __future__ import annotations

import copy
from typing import Any, Dict


def _vit(hidden, heads, layers, inter, head_hidden, labels=10, px=224, patch=16, attn_layers=1):
    return dict(
        attention_probs_dropout_prob=0.1,
        explainer_attn_num_layers=attn_layers,
        explainer_head_hidden_size=head_hidden,
        explainer_normalize=True,
        hidden_dropout_prob=0.1,
        hidden_size=hidden,
        intermediate_size=inter,
        layer_norm_eps=1e-12,
        num_attention_heads=heads,
        num_hidden_layers=layers,
        num_labels=labels,
        img_channels=3,
        img_px_size=px,
        img_patch_size=patch,
    )


def _bert(hidden, heads, layers, inter, head_hidden, max_pos, vocab=30522, labels=2, attn_layers=1):
    return dict(
        attention_probs_dropout_prob=0.1,
        explainer_attn_num_layers=attn_layers,
        explainer_head_hidden_size=head_hidden,
        explainer_normalize=True,
        hidden_dropout_prob=0.1,
        hidden_size=hidden,
        intermediate_size=inter,
        layer_norm_eps=1e-12,
        max_position_embeddings=max_pos,
        num_attention_heads=heads,
        num_hidden_layers=layers,
        num_labels=labels,
        pad_token_id=0,
        type_vocab_size=2,
        vocab_size=vocab,
    )


def _ltt(base: Dict[str, Any], s_hidden: int, s_inter: int, s_head: int, attn_layers: int = 1) -> Dict[str, Any]:
    """LTT configuration (reference models/ltt_vit.py:14-32, models/ltt_bert.py:20-39) on top of a vanilla one: the explainer_*
    fields are renamed to explainer_s_* and the side ladder's widths are added."""
    out = {k: v for k, v in base.items() if k not in ("explainer_attn_num_layers", "explainer_head_hidden_size")}
    out.update(explainer_s_attn_num_layers=attn_layers, explainer_s_head_hidden_size=s_head, s_attn_hidden_size=s_hidden,
               s_attn_intermediate_size=s_inter)
    return out


CONFIGS: Dict[str, Dict[str, Any]] = {
    # reduced shapes for fast CPU parity (head dim stays 64, the only size the tensor-core attention tiles)
    "vit_mini": _vit(128, 2, 2, 256, 192),
    "vit_mini_px64": _vit(128, 2, 2, 256, 192, px=64),  # T = 17: ragged/small-sequence edge case
    # reference experiments/vit_tiny_imagenette_vanilla/.hparams.json:20-35
    "vit_tiny": _vit(192, 3, 12, 768, 768),
    # reference experiments/vit_base_imagenette_vanilla/.hparams.json:20-35
    "vit_base": _vit(768, 12, 12, 3072, 3072),
    # reference experiments/vit_large_imagenette_vanilla/.hparams.json:20-35
    "vit_large": _vit(1024, 16, 24, 4096, 4096),
    "bert_mini": _bert(128, 2, 2, 256, 192, max_pos=32, vocab=1000),
    # long-sequence edge case: the reference's checked-in BERT configuration has 512 positions
    "bert_mini_512": _bert(128, 2, 2, 256, 192, max_pos=512, vocab=1000),
    # reference experiments/bert_base_tayp_vanilla/.hparams.json:14-30 with max_position_embeddings=128
    # (BASELINE.json "128-token" configuration)
    "bert_base_128": _bert(768, 12, 12, 3072, 3072, max_pos=128),
    "bert_base_512": _bert(768, 12, 12, 3072, 3072, max_pos=512),
}
# the same ViT-Base/16 configuration under a second fixture name: 4 inputs x 32 coalitions = 128 rows (M = 25 216 token rows),
# the smallest shape at which the bench's own kernel variants engage (CTA-pair GEMMs, LayerNorm-folded chain, first-block
# sharing, CLS-only last block) — tests/golden/model_vit_base_b4s32.npz, train_vit_base_b4s32.npz
CONFIGS["vit_base_b4s32"] = CONFIGS["vit_base"]
# LTT (ladder side tuning) variants; side head dim = s_attn_hidden_size / num_attention_heads
CONFIGS.update({
    "ltt_vit_mini": _ltt(CONFIGS["vit_mini"], 32, 64, 48),                 # side head dim 16
    "ltt_vit_tiny": _ltt(CONFIGS["vit_tiny"], 48, 192, 192),               # side head dim 16, 12 ladder rungs
    "ltt_bert_mini": _ltt(CONFIGS["bert_mini"], 16, 64, 48),               # side head dim 8
    # reference experiments/bert_base_tayp_ltt/.hparams.json:14-32 with max_position_embeddings = 128 (side head dim 8)
    "ltt_bert_base_128": _ltt(CONFIGS["bert_base_128"], 96, 384, 3072),
})


def get_config(name: str) -> Dict[str, Any]:
    return copy.deepcopy(CONFIGS[name])


def is_vit(cfg: Dict[str, Any]) -> bool:
    return "img_px_size" in cfg


def n_players(cfg: Dict[str, Any]) -> int:
    """reference recipes/vanilla_vit.py:49 and recipes/vanilla_bert.py:55"""
    if is_vit(cfg):
        return (cfg["img_px_size"] // cfg["img_patch_size"]) ** 2
    return cfg["max_position_embeddings"] - 1
