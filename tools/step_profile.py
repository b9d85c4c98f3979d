#!/usr/bin/env python
"""Per-launch timeline of one masked-evaluation step (CUDA events around every C-ABI call, in-step clocks).
Test infrastructure; run under gpurun:  python tools/step_profile.py [images]"""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import bench  # noqa: E402
from autognothi_b200 import _native as nat  # noqa: E402
from autognothi_b200.models import shapley as ash  # noqa: E402
from autognothi_b200.recipes.vanilla_vit import vanilla_vit_recipe  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    S = bench.S_COALITIONS
    dev = torch.device("cuda:0")
    rec = vanilla_vit_recipe()
    cfg = rec.t_config(**dict(bench.VIT_BASE))
    n = rec.n_players(cfg)
    torch.manual_seed(3407)
    surrogate = rec.t_surrogate(cfg).to(dev).eval()
    surrogate.agb_precision = "bf16"
    xs = torch.randn((B, 3, 224, 224), device=dev)

    def step(i):
        pm = ash.mask_shapley_new(B * S, n, device=dev, rng="philox", seed=3407, offset=i * B * S, packed=True)
        return rec.fw_surrogate(surrogate, xs, pm)[0]

    with torch.no_grad():
        for i in range(4):
            step(i)
        torch.cuda.synchronize()
        nat.PROFILE = []
        for i in range(4, 8):
            step(i)
        torch.cuda.synchronize()
        prof, nat.PROFILE = nat.PROFILE, None
    per_step = len(prof) // 4
    last = prof[-per_step:]
    tot = sum(a.elapsed_time(b) for _, _, a, b, *_ in last)
    print(f"{per_step} launches per step, sum of kernel times {tot:.2f} ms (last of 4 profiled steps)")
    # layer 5 window: find the 6th attention call
    att = [i for i, (nm, *_) in enumerate(last) if nm == "agb_masked_attention_bf16"]
    lo, hi = att[5] - 1, att[6] - 1
    for nm, meta, a, b, *_ in last[lo:hi]:
        t = a.elapsed_time(b) * 1e3
        tf = f"{meta / t * 1e-6:8.1f} TFLOP/s" if meta else ""
        print(f"  {nm:32s} {t:9.1f} us {tf}")
    agg = {}
    for nm, meta, a, b, *_ in last:
        acc = agg.setdefault(nm, [0.0, 0.0, 0])
        acc[0] += a.elapsed_time(b); acc[1] += meta or 0.0; acc[2] += 1
    for nm, (t, fl, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{nm:32s} x{c:<3d} {t:8.3f} ms  {100 * t / tot:5.1f} %  {fl / t * 1e-9 if fl else 0:8.1f} TFLOP/s")


if __name__ == "__main__":
    main()
