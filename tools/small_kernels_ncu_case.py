#!/usr/bin/env python
"""The HBM-bound small kernels of the training step at bench shapes (32 images x 32 coalitions, ViT-Base/16): coalition
sampling, explainer head + efficiency normalisation (forward / adjoint), Shapley loss (forward / adjoint), rank masks.
  ncu --set full -k regex:'shapley|explainer_head|rank_masks|normalize' python tools/small_kernels_ncu_case.py"""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from autognothi_b200 import ops  # noqa: E402
from autognothi_b200.models import shapley as ash  # noqa: E402

dev = torch.device("cuda:0")
B, S, n, C, E = 32, 32, 196, 10, 3072
T = n + 1
for _ in range(2):
    pm = ash.mask_shapley_new(B * S, n, device=dev, rng="philox", seed=1, packed=True)
    h = torch.randn(B * T, E, device=dev).to(torch.bfloat16)
    w_c, b_c = torch.randn(C, E, device=dev) * 0.02, torch.zeros(C, device=dev)
    grand, null = torch.rand(B, C, device=dev), torch.rand(1, C, device=dev)
    phi = ops.explainer_head_fwd(h, B, T, w_c, b_c, grand, null, True)
    v_s = torch.rand(B * S, C, device=dev)
    phi_g = phi.clone().requires_grad_(True)
    loss = ash.loss_shapley_new(B, S, n, pm, null, v_s, grand, phi_g)
    loss.backward()
    dW, db = torch.zeros_like(w_c), torch.zeros_like(b_c)
    ops.explainer_head_bwd(phi_g.grad, h, B, T, w_c, True, dW, db)
    attr = torch.randn(B * C, n, device=dev)
    ops.rank_masks(attr, torch.linspace(0, n, 25).long(), n, 1, want_dense=False)
torch.cuda.synchronize()
