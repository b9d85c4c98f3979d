#!/usr/bin/env python
"""Kernel table of one LTT masked-evaluation step at the bench shape (ViT-Base/16 frozen + ladder 96, 32 images x 32 coalitions):
where the ladder's time on top of the backbone goes.  torch.profiler (CUPTI) kernel times, 3 profiled steps after warm-up.
    python tools/ltt_eval_profile.py > gpurun_out/ltt_eval_profile.txt
"""
import sys
from collections import OrderedDict

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import bench  # noqa: E402
from autognothi_b200.models import shapley as ash  # noqa: E402
from autognothi_b200.recipes.ltt_vit import ltt_vit_recipe  # noqa: E402

dev = torch.device("cuda:0")
cfgd = dict(bench.VIT_BASE)
lcfgd = {k: v for k, v in cfgd.items() if k not in ("explainer_attn_num_layers", "explainer_head_hidden_size")}
lcfgd.update(explainer_s_attn_num_layers=1, explainer_s_head_hidden_size=3072, s_attn_hidden_size=96, s_attn_intermediate_size=384)
lrec = ltt_vit_recipe()
lcfg = lrec.t_config(**lcfgd)
torch.manual_seed(3407)
lsrg = lrec.t_surrogate(lcfg).to(dev).eval()
lsrg.agb_precision = "bf16"
B, S = 32, 32
n = lrec.n_players(lcfg)
xs = torch.randn(B, 3, 224, 224, device=dev)
pm = ash.mask_shapley_new(B * S, n, device=dev, rng="philox", seed=1, packed=True)


def step():
    with torch.no_grad():
        lrec.fw_surrogate(lsrg, xs, pm)


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    step()
e1.record()
torch.cuda.synchronize()
print(f"LTT eval step: {e0.elapsed_time(e1) / 5:.2f} ms for {B * S} masked evaluations")
steps = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
agg = OrderedDict()
for ev in prof.events():
    if ev.device_type.name != "CUDA":
        continue
    a = agg.setdefault(ev.name[:150], [0, 0.0])
    a[0] += 1
    a[1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
total = sum(a[1] for a in agg.values())
print(f"sum of kernel times {total / steps / 1e3:.2f} ms/step")
for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{us / steps / 1e3:9.3f} ms  x{cnt // steps:<4d} avg {us / cnt:8.1f} us  {name}")
