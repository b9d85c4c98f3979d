"""Callers of the masked-surrogate hot path with their mask generators moved on device.

Mirrors, with the same names and return shapes,
  * reference scripts/measure_faithfulness.py:182-251  (`_get_perturbed_samples`, the `_infer` closure): insertion /
    deletion curves — rank the players by attribution, flip the top stops[i] of them in a base mask, evaluate the
    surrogate on every perturbed mask and read the class probability;
  * reference scripts/measure_accuracy.py:82-110       (`_measure_surrogate_epoch`): accuracy of the surrogate when a
    fixed number of uniformly chosen players is masked out.
The reference builds the masks in numpy / python loops, replicates the input once per mask (`repeat_interleave`) and
reads results back with one `.item()` per row.  Here the ranking and the packed bitmasks come from one kernel
(`agb_rank_masks`), the input is embedded once and broadcast to its masks inside the embedding kernel, all classes and
all stops go through ONE batched surrogate call, and there is one device->host copy at the end.
"""
from __future__ import annotations

from typing import Any, Dict, Iterable, Optional, Tuple

import numpy as np
import torch
from torch import Tensor

from . import ops
from .models.shapley import PackedMasks, mask_uniform_selective

CurvePoint = Dict[int, Dict[int, float]]


def perturbation_stops(n_players: int, steps: int) -> np.ndarray:
    """stops = np.linspace(0, n_players, min(n_players, steps), dtype=int64)  (reference l.241)"""
    return np.linspace(0, n_players, min(n_players, steps), dtype=np.int64)


def get_perturbed_samples(explanations: Tensor, n_players: int, steps: int, mask_base: int, *, packed: bool = False
                          ) -> Tuple[Tensor, Any]:
    """reference scripts/measure_faithfulness.py:225-251.
    explanations: (n_players,) or (R, n_players) attributions on the GPU.
    -> (stops (steps',) int64, masks): masks (steps', n) int64 for one row of attributions (the reference's shape),
       (R * steps', n) for R rows (row r * steps' + i), or PackedMasks when packed=True."""
    assert explanations.is_cuda, "attributions must live on the GPU (no CPU path)"
    attr = explanations.reshape(-1, n_players).float()
    stops_np = perturbation_stops(n_players, steps)
    stops = torch.from_numpy(stops_np)
    words, dense = ops.rank_masks(attr, stops, n_players, mask_base, want_dense=not packed)
    return stops.to(explanations.device), (PackedMasks(words, n_players) if packed else dense)


def faithfulness_infer(recipe, surrogate, Xs: Tensor, explanation: Tensor, steps: int, mask_base: int,
                       batch_size: Optional[int] = None) -> CurvePoint:
    """The `_infer` closure of reference scripts/measure_faithfulness.py:182-220 for one input.
    Xs (1, ...) on the GPU, explanation (1, C, n_players) -> {class: {stop: surrogate probability of that class}}.
    `batch_size` is accepted for signature compatibility; the surrogate engine chunks rows itself."""
    assert Xs.shape[0] == 1 and explanation.dim() == 3 and explanation.shape[0] == 1
    n_classes, n_players = explanation.shape[1], explanation.shape[2]
    stops, pm = get_perturbed_samples(explanation[0], n_players, steps, mask_base, packed=True)   # rows c * steps' + i
    nst = stops.numel()
    with torch.no_grad():
        probs, _ = recipe.fw_surrogate(surrogate, Xs, PackedMasks(pm.words.reshape(n_classes * nst, -1), n_players))
    picked = probs.reshape(n_classes, nst, -1)[torch.arange(n_classes, device=probs.device), :,
                                               torch.arange(n_classes, device=probs.device)]   # (C, steps')
    host, stops_h = picked.cpu().numpy(), stops.cpu().numpy()
    result: CurvePoint = {}
    for c in range(n_classes):
        ret: Dict[int, float] = {}
        for i in range(nst):                     # duplicate stops overwrite, as in the reference's dict
            ret[int(stops_h[i])] = float(host[c, i])
        result[c] = ret
    return result


def measure_surrogate_accuracy(recipe, surrogate, batches: Iterable[Tuple[Tensor, Tensor]], n_players: int,
                               n_masked_players: int, *, rng: str = "philox", seed: int = 0) -> float:
    """reference scripts/measure_accuracy.py:82-110 (`_measure_surrogate_epoch`): fraction of inputs whose arg-max
    surrogate class equals the label when `n_masked_players` uniformly chosen players are masked out.
    batches yields (Xs, Zs) already on the GPU (the output of recipe.gen_input)."""
    correct = torch.zeros((), dtype=torch.int64)
    total = 0
    offset = 0
    for Xs, Zs in batches:
        B = Xs.shape[0]
        masks = mask_uniform_selective(B, n_players, n_masked_players, device=Xs.device, rng=rng, seed=seed, offset=offset,
                                       packed=True)
        offset += B
        with torch.no_grad():
            ys, _ = recipe.fw_surrogate(surrogate, Xs, masks)
        if correct.device != ys.device:
            correct = correct.to(ys.device)
        correct += (ys.argmax(dim=1) == Zs.to(ys.device)).sum()
        total += B
    return float(correct.item()) / max(total, 1)
