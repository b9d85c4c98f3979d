"""Shared recipe helpers: mask normalisation at the boundary and state-dict conversions."""
from __future__ import annotations

from typing import Dict, Tuple

import torch
from torch import Tensor, nn

from ..models.shapley import MaskLike, PackedMasks


def resolve_masks(xs: Tensor, mask: MaskLike, n_players: int) -> Tuple[PackedMasks, int]:
    """-> (packed masks with B*S rows, S).  Accepts the reference's (N, n) int64 player mask (S = 1,
    one row per input row), a (B, S, n) tensor, or PackedMasks whose row count is a multiple of B.
    This is where the reference prepends the always-on CLS column (recipes/vanilla_vit.py:219-224)."""
    B = xs.shape[0]
    if isinstance(mask, PackedMasks):
        assert mask.n_players == n_players and mask.rows % B == 0, "mask rows must be a multiple of the batch"
        return mask, mask.rows // B
    assert mask.is_cuda, "masks must be CUDA tensors (no CPU path)"
    assert mask.shape[-1] == n_players, f"mask last dim must be n_players={n_players}"
    if mask.dim() == 3:
        assert mask.shape[0] == B
        return PackedMasks.from_dense(mask), mask.shape[1]
    assert mask.dim() == 2 and mask.shape[0] == B, "mask must be (batch, n_players)"
    return PackedMasks.from_dense(mask), 1


@torch.no_grad()
def copy_matching(src: Dict[str, Tensor], dst: nn.Module, prefixes: Tuple[str, ...], dst_prefix: str = "") -> None:
    """Copy every tensor of `src` whose key starts with one of `prefixes` into `dst` (same key, optionally
    re-rooted under dst_prefix).  The conv_* rules of the reference (recipes/vanilla_vit.py:116-194) reduce
    to this on the vanilla path: 'keep' rules copy, 'New()' rules leave the fresh initialisation."""
    own = dst.state_dict()
    for k, v in src.items():
        if any(k.startswith(p) for p in prefixes):
            kk = dst_prefix + k
            if kk not in own:
                raise KeyError(f"conversion target has no key {kk!r}")
            if own[kk].shape != v.shape:
                raise ValueError(f"shape mismatch for {kk}: {tuple(own[kk].shape)} vs {tuple(v.shape)}")
            own[kk].copy_(v)
