#!/usr/bin/env python
"""Masked-evaluation throughput vs batch size (the reference's scripts call fw_surrogate with 2-8 inputs x 4-32 coalitions):
wall-clock per call, to see where the host side (Python + ctypes launches) instead of the GPU sets the pace."""
import sys
import time

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import bench  # noqa: E402
from autognothi_b200.models import shapley as ash  # noqa: E402
from autognothi_b200.recipes.vanilla_vit import vanilla_vit_recipe  # noqa: E402

dev = torch.device("cuda:0")
rec = vanilla_vit_recipe()
cfg = rec.t_config(**dict(bench.VIT_BASE))
n = rec.n_players(cfg)
torch.manual_seed(3407)
srg = rec.t_surrogate(cfg).to(dev).eval()
srg.agb_precision = "bf16"
for B, S in ((1, 4), (2, 32), (8, 32), (32, 32)):
    xs = torch.randn(B, 3, 224, 224, device=dev)
    pm = ash.mask_shapley_new(B * S, n, device=dev, rng="philox", seed=1, packed=True)
    with torch.no_grad():
        for _ in range(3):
            rec.fw_surrogate(srg, xs, pm)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        it = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(it):
            rec.fw_surrogate(srg, xs, pm)
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / it * 1e3
        gpu = e0.elapsed_time(e1) / it
    print(f"B={B:3d} S={S:3d} rows={B * S:5d}: wall {wall:7.3f} ms/call  gpu-span {gpu:7.3f} ms  {B * S / wall * 1e3:9.0f} evals/s")
