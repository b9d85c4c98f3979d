#!/usr/bin/env python
"""A/B on one box: ViT-Base/16 masked evaluation (32 images x 32 coalitions) with engine.KEPT_FIRST_ORDER on / off, alternating."""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import bench  # noqa: E402
from autognothi_b200 import engine  # noqa: E402
from autognothi_b200.models import shapley as ash  # noqa: E402
from autognothi_b200.recipes.vanilla_vit import vanilla_vit_recipe  # noqa: E402

dev = torch.device("cuda:0")
rec = vanilla_vit_recipe()
cfg = rec.t_config(**dict(bench.VIT_BASE))
n = rec.n_players(cfg)
torch.manual_seed(3407)
srg = rec.t_surrogate(cfg).to(dev).eval()
srg.agb_precision = "bf16"
B, S = 32, 32
xs = torch.randn(B, 3, 224, 224, device=dev)


def run(flag, iters=10):
    engine.KEPT_FIRST_ORDER = flag
    with torch.no_grad():
        for i in range(3):
            pm = ash.mask_shapley_new(B * S, n, device=dev, rng="philox", seed=1, offset=i * B * S, packed=True)
            rec.fw_surrogate(srg, xs, pm)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            pm = ash.mask_shapley_new(B * S, n, device=dev, rng="philox", seed=1, offset=(3 + i) * B * S, packed=True)
            rec.fw_surrogate(srg, xs, pm)
        e1.record()
        torch.cuda.synchronize()
    return B * S * iters / (e0.elapsed_time(e1) * 1e-3)


for rep in range(3):
    print(f"rep {rep}: kept-first order off {run(False):9.0f} evals/s   on {run(True):9.0f} evals/s")
