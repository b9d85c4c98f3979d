#!/usr/bin/env python
"""One LTT masked evaluation (ViT-Base/16 + ladder 96, 256 rows) and one dropout training step, for ncu captures of the
narrow-head attention and the DROP instantiations:  ncu --set full -k regex:'narrow|bwd_small|dropout|Lb1' python tools/ltt_ncu_case.py"""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import bench  # noqa: E402
from autognothi_b200.models import shapley as ash  # noqa: E402
from autognothi_b200.recipes.ltt_vit import ltt_vit_recipe  # noqa: E402
from autognothi_b200.recipes.vanilla_vit import vanilla_vit_recipe  # noqa: E402

dev = torch.device("cuda:0")
cfgd = dict(bench.VIT_BASE)
lcfgd = {k: v for k, v in cfgd.items() if k not in ("explainer_attn_num_layers", "explainer_head_hidden_size")}
lcfgd.update(explainer_s_attn_num_layers=1, explainer_s_head_hidden_size=3072, s_attn_hidden_size=96, s_attn_intermediate_size=384)
lrec = ltt_vit_recipe()
lcfg = lrec.t_config(**lcfgd)
torch.manual_seed(3407)
lsrg = lrec.t_surrogate(lcfg).to(dev).eval()
lsrg.agb_precision = "bf16"
B, S = 8, 32
n = lrec.n_players(lcfg)
xs = torch.randn(B, 3, 224, 224, device=dev)
pm = ash.mask_shapley_new(B * S, n, device=dev, rng="philox", seed=1, packed=True)
with torch.no_grad():
    lrec.fw_surrogate(lsrg, xs, pm)
# vanilla explainer training step with dropout (train() mode)
rec = vanilla_vit_recipe()
cfg = rec.t_config(**cfgd)
exp = rec.t_explainer(cfg).to(dev).train()
exp.agb_precision = "bf16"
ones = ash.PackedMasks.ones(B, n, dev)
C = cfgd["num_labels"]
grand, null = torch.full((B, C), 1.0 / C, device=dev), torch.full((1, C), 1.0 / C, device=dev)
v_s = torch.full((B * S, C), 1.0 / C, device=dev)
phi, _ = rec.fw_explainer(exp, xs, ones, grand, null)
loss = ash.loss_shapley_new(B, S, n, pm, null, v_s, grand, phi)
loss.backward()
torch.cuda.synchronize()
