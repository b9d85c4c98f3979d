#!/usr/bin/env python
"""Explainer head + efficiency normalisation (forward / adjoint) at the bench shape, CUDA-event timing, warm.
Algorithmic bytes: h once (B*T*E*2 B bf16) [+ dh once for the adjoint]; roofline = measured HBM bandwidth."""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from autognothi_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def timed(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for B, T, E, C in ((32, 197, 3072, 10), (128, 197, 3072, 10), (32, 128, 3072, 2), (256, 128, 3072, 2)):
    h = torch.randn(B * T, E, device=dev).to(torch.bfloat16)
    w_c, b_c = torch.randn(C, E, device=dev) * 0.02, torch.zeros(C, device=dev)
    grand, null = torch.rand(B, C, device=dev), torch.rand(1, C, device=dev)
    dphi = torch.randn(B, C, T - 1, device=dev)
    dW, db = torch.zeros_like(w_c), torch.zeros_like(b_c)
    big = torch.empty(64 << 20, device=dev)       # 256 MB: flushes L2 between calls

    def fwd():
        big.zero_()
        return ops.explainer_head_fwd(h, B, T, w_c, b_c, grand, null, True)

    def bwd():
        big.zero_()
        return ops.explainer_head_bwd(dphi, h, B, T, w_c, True, dW, db)

    def flush():
        big.zero_()

    t0 = timed(flush)
    tf, tb = timed(fwd) - t0, timed(bwd) - t0
    bytes_f = B * T * E * 2
    print(f"head B={B} T={T} E={E} C={C}: fwd {tf:7.1f} us ({bytes_f / tf * 1e-3:7.1f} GB/s)   bwd {tb:7.1f} us "
          f"({2 * bytes_f / tb * 1e-3:7.1f} GB/s, h read + dh written)")
