#!/usr/bin/env python
"""Kernel-level breakdown of the explainer training step (torch.profiler / CUPTI).  Test infrastructure;
run under gpurun:  python tools/train_profile.py [images]"""
import sys
import time

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import bench  # noqa: E402
from autognothi_b200.dist import GradAllReducer  # noqa: E402
from autognothi_b200.models import shapley as ash  # noqa: E402
from autognothi_b200.recipes.vanilla_vit import vanilla_vit_recipe  # noqa: E402


def main():
    Bt = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    S = bench.S_COALITIONS
    dev = torch.device("cuda:0")
    rec = vanilla_vit_recipe()
    cfg = rec.t_config(**dict(bench.VIT_BASE))
    n = rec.n_players(cfg)
    torch.manual_seed(3407)
    surrogate = rec.t_surrogate(cfg).to(dev).eval()
    surrogate.agb_precision = "bf16"
    explainer = rec.conv_surrogate_explainer(cfg, None, surrogate).train()
    explainer.agb_precision = "bf16"
    opt = torch.optim.AdamW(explainer.parameters(), lr=5e-5, fused=True)
    reducer = GradAllReducer(explainer.parameters(), bucket_mb=64.0)
    ones = ash.PackedMasks.ones(Bt, n, dev)
    xs = torch.randn((Bt, 3, 224, 224), device=dev)
    with torch.no_grad():
        null, _ = rec.fw_surrogate(surrogate, rec.gen_null(cfg, None, dev), ash.PackedMasks.ones(1, n, dev))

    def evals(i):
        pm = ash.mask_shapley_new(Bt * S, n, device=dev, rng="philox", seed=99, offset=i * Bt * S, packed=True)
        with torch.no_grad():
            v_s, _ = rec.fw_surrogate(surrogate, xs, pm)
            grand, _ = rec.fw_surrogate(surrogate, xs, ones)
        return pm, v_s, grand

    def explain(pm, v_s, grand):
        phi, _ = rec.fw_explainer(explainer, xs, ones, grand, null)
        loss = ash.loss_shapley_new(Bt, S, n, pm, null, v_s, grand, phi)
        loss.backward()
        reducer.allreduce()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    for i in range(2):
        explain(*evals(i))
    torch.cuda.synchronize()
    # wall/device split of the two halves
    for name in ("evals", "explainer fwd+bwd+opt"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        args = evals(0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        for i in range(3):
            if name == "evals":
                evals(i)
            else:
                explain(*args)
        e1.record()
        t_host = (time.perf_counter() - t0) / 3 * 1e3
        torch.cuda.synchronize()
        print(f"{name:24s}: device {e0.elapsed_time(e1) / 3:8.2f} ms/step   host-side issue time {t_host:8.2f} ms/step")
    args = evals(0)
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for i in range(3):
            explain(*args)
        torch.cuda.synchronize()
    rows = [(k.key, k.device_time_total / 3e3, k.count // 3) for k in prof.key_averages() if k.device_time_total > 0]
    rows.sort(key=lambda r: -r[1])
    tot = sum(r[1] for r in rows if not r[0].startswith(("aten::", "cuda", "Memcpy", "Memset", "autograd", "Optimizer", "_")))
    print(f"explainer step kernels (per step, 3-step average):")
    for k, ms, c in rows[:45]:
        print(f"  {ms:8.3f} ms  x{c:<4d} {k[:110]}")


if __name__ == "__main__":
    main()
