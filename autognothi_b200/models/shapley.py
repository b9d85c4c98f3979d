"""Drop-in for reference models/shapley.py: same function names, argument order and return layouts,
computed by the CUDA kernels behind include/autognothi_b200.h.

Differences a caller can see (all additive):
  * samplers return CUDA tensors (the reference returns CPU tensors that every caller immediately
    moves with `.to(device)`; that call becomes a no-op);
  * `rng="torch"` (default) consumes the global torch CPU generator exactly like the reference, so the
    masks are bit-identical to the reference's under the same `torch.manual_seed`; `rng="philox"`
    draws on the device (no host RNG, no H2D copy) — the throughput path;
  * `packed=True` returns the packed-bitmask form the kernels consume (bit 0 = CLS, bit j+1 = player j).
"""
from __future__ import annotations

from typing import Optional, Tuple, Union

import torch
from torch import Tensor

from .. import ops


class PackedMasks:
    """Packed coalition masks: `words` is (rows, ceil((n_players+1)/32)) int32 on the device."""

    __slots__ = ("words", "n_players")

    def __init__(self, words: Tensor, n_players: int):
        assert words.dim() == 2 and words.dtype == torch.int32 and words.is_cuda
        assert words.shape[1] * 32 >= n_players + 1
        self.words = words.contiguous()
        self.n_players = n_players

    @property
    def rows(self) -> int:
        return self.words.shape[0]

    def dense(self) -> Tensor:
        """(rows, n_players) int64 — the reference's layout."""
        return ops.unpack_masks(self.words, self.n_players, skip=1)

    def reshape_rows(self, rows: int) -> "PackedMasks":
        return PackedMasks(self.words.reshape(rows, -1), self.n_players)

    @staticmethod
    def from_dense(mask: Tensor) -> "PackedMasks":
        assert mask.is_cuda, "masks must live on the GPU (no CPU path)"
        m2 = mask.reshape(-1, mask.shape[-1]).to(torch.int64)
        return PackedMasks(ops.pack_masks(m2, prepend_cls=True), mask.shape[-1])

    @staticmethod
    def ones(rows: int, n_players: int, device) -> "PackedMasks":
        T = n_players + 1
        W = ops.mask_words(T)
        w = torch.full((rows, W), -1, dtype=torch.int32, device=device)
        tail = T - 32 * (W - 1)
        if tail < 32:
            w[:, W - 1] = (1 << tail) - 1
        return PackedMasks(w, n_players)


MaskLike = Union[Tensor, PackedMasks]


def as_packed(mask: MaskLike, rows: Optional[int] = None) -> PackedMasks:
    pm = mask if isinstance(mask, PackedMasks) else PackedMasks.from_dense(mask)
    if rows is not None:
        assert pm.rows == rows, f"expected {rows} mask rows, got {pm.rows}"
    return pm


def _default_device(device) -> torch.device:
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise RuntimeError("autognothi_b200 needs a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


_PREFIX_CACHE = {}


def _shapley_prefix(n_players: int, device: torch.device) -> Tensor:
    """The inverse-CDF table, built with the SAME torch CPU ops as reference models/shapley.py:65-67,132
    (a 195-entry host-side constant; bit-identical to the reference's by construction), cached on device."""
    key = (n_players, str(device))
    if key not in _PREFIX_CACHE:
        probs = torch.arange(1, n_players) * (n_players - torch.arange(1, n_players))
        probs = 1 / probs
        probs = probs / probs.sum()
        prefix = torch.cumsum(probs, dim=0) - probs
        _PREFIX_CACHE[key] = prefix.to(device=device, dtype=torch.float32).contiguous()
    return _PREFIX_CACHE[key]


def mask_shapley_new(n_mask_samples: int, n_players: int, *, device=None, rng: str = "torch", seed: int = 0,
                     offset: int = 0, packed: bool = False):
    """Paired Shapley-kernel coalition sampler (reference models/shapley.py:56-79).
    Returns (n_mask_samples, n_players) int64 {0,1}; rows 2i / 2i+1 are complements."""
    assert n_mask_samples % 2 == 0  # reference l.62
    assert n_players >= 2
    dev = _default_device(device)
    pairs = n_mask_samples // 2
    prefix = _shapley_prefix(n_players, dev)
    if rng == "torch":
        u_players = torch.rand(pairs, n_players)        # reference l.69 (drawn first)
        u_size = torch.rand((pairs, 1)).reshape(-1)      # reference l.133 (inside _torch_choice)
        words, dense = ops.shapley_masks(prefix, pairs, n_players, u_players=u_players.to(dev), u_size=u_size.to(dev),
                                         want_dense=not packed)
    elif rng == "philox":
        words, dense = ops.shapley_masks(prefix, pairs, n_players, seed=seed, offset=offset, want_dense=not packed)
    else:
        raise ValueError(f"unknown rng {rng!r}")
    return PackedMasks(words, n_players) if packed else dense


def loss_logits_kl_divergence(ref: Tensor, current: Tensor) -> Tensor:
    """reference models/shapley.py:96-106 — the surrogate-training objective: batch-mean KL divergence with
    log_softmax(ref) as the input and softmax(current) as the target (both arguments are the models' outputs, which
    are already probabilities there).  (B, C) tensors: stays in torch autograd."""
    import torch.nn.functional as F
    return F.kl_div(input=F.log_softmax(ref, dim=-1), target=F.softmax(current, dim=-1), reduction="batchmean")


def mask_purely_uniform(batch_size: int, n_features: int, *, device=None, rng: str = "torch", seed: int = 0,
                        offset: int = 0, packed: bool = False):
    """reference models/shapley.py:109-115"""
    dev = _default_device(device)
    if rng == "torch":
        u_players = torch.rand((batch_size, n_features))
        u_row = torch.rand((batch_size, 1))
        words, dense = ops.uniform_masks(batch_size, n_features, dev, u_players=u_players.to(dev), u_row=u_row.to(dev),
                                         want_dense=not packed)
    else:
        words, dense = ops.uniform_masks(batch_size, n_features, dev, seed=seed, offset=offset, want_dense=not packed)
    return PackedMasks(words, n_features) if packed else dense


def mask_uniform_selective(batch_size: int, n_features: int, n_masked: int, *, device=None, rng: str = "python",
                           seed: int = 0, offset: int = 0, packed: bool = False):
    """reference models/shapley.py:118-128: masks with exactly `n_masked` players masked out.
    rng="python" replays the reference's own host algorithm (random.shuffle -> identical masks under the same
    random.seed) and uploads it; rng="philox" draws random keys on the device and masks the n_masked top-ranked ones
    (same uniform distribution over subsets, no host loop)."""
    dev = _default_device(device)
    if rng == "python":
        import random
        ret = []
        for _ in range(batch_size):
            ids = list(range(n_features))
            random.shuffle(ids)
            chosen = set(ids[:n_masked])
            ret.append([0 if i in chosen else 1 for i in range(n_features)])
        dense = torch.tensor(ret, dtype=torch.long).reshape(batch_size, n_features).to(dev)
        return PackedMasks(ops.pack_masks(dense, prepend_cls=True), n_features) if packed else dense
    stops = torch.tensor([n_masked], dtype=torch.int32)
    words, dense = ops.rank_masks(None, stops, n_features, 1, rows=batch_size, device=dev, seed=seed, offset=offset,
                                  want_dense=not packed)
    return PackedMasks(words, n_features) if packed else dense


class _NormalizeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, grand, null):
        ctx.T = pred.shape[1]
        ctx.null_shape = null.shape
        return ops.normalize_shapley(pred, grand, null)

    @staticmethod
    def backward(ctx, g):
        # out = pred + ((grand - null) - sum_t pred)/T  ->  dpred = g - mean_t(g); dgrand = mean-sum; dnull = -that
        gs = g.sum(dim=1)
        dnull = (-(gs.sum(dim=0, keepdim=True)) / ctx.T).reshape(ctx.null_shape)     # null may be (1, C) or (C,)
        return g - gs.unsqueeze(1) / ctx.T, gs / ctx.T, dnull


def normalize_shapley_explanation(pred: Tensor, grand: Tensor, null: Tensor) -> Tensor:
    """Additive efficiency normalisation (reference models/shapley.py:82-93): the divisor is
    pred.shape[1] — callers pass the un-sliced (B, T, C) tensor, CLS included."""
    assert pred.is_cuda and pred.dim() == 3
    if torch.is_grad_enabled() and (pred.requires_grad or grand.requires_grad or null.requires_grad):
        return _NormalizeFn.apply(pred, grand, null)
    return ops.normalize_shapley(pred, grand, null)


class _ShapleyLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, phi, words, v_0, v_s, B, S, n):
        loss, resid = ops.shapley_loss_fwd(words, v_0, v_s, phi, B, S, n)
        ctx.save_for_backward(words, resid)
        ctx.dims = (B, S, n, phi.shape[1])
        return loss

    @staticmethod
    def backward(ctx, g):
        words, resid = ctx.saved_tensors
        B, S, n, C = ctx.dims
        return ops.shapley_loss_bwd(words, resid, g, B, S, n, C), None, None, None, None, None, None


def loss_shapley_new(batch_size: int, n_mask_samples: int, n_players: int, mask: MaskLike, v_0: Tensor, v_s: Tensor,
                     v_1: Tensor, phi: Tensor) -> Tensor:
    """Surrogate-vs-explainer loss (reference models/shapley.py:9-53):
    n_players * mean((v_0 + mask @ phi^T - v_s)^2); `v_1` is accepted and unused, as in the reference.
    mask: (B, S, n) int64 or PackedMasks with B*S rows (row order b*S+s)."""
    _ = v_1
    pm = as_packed(mask, batch_size * n_mask_samples)
    assert pm.n_players == n_players and phi.shape == (batch_size, phi.shape[1], n_players)
    if phi.dtype != torch.float32:
        phi = phi.float()          # the loss is computed in fp32 (the reference's `mask.float() @ phi` needs matching dtypes)
    return _ShapleyLossFn.apply(phi, pm.words, v_0, v_s, batch_size, n_mask_samples, n_players)
