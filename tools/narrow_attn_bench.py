"""Narrow-head (LTT side ladder) attention: accuracy against the fp32 CUDA-core kernel and per-launch time at the bench shape.

    python tools/narrow_attn_bench.py            # warp-level tensor-core kernel (default path)
    AGB_NARROW_SIMT=1 python tools/narrow_attn_bench.py    # the CUDA-core kernel it replaced, for the A/B

Shapes: ViT-Base ladder (T = 197, 12 heads x 8, masked logit := 0) and BERT-base ladder (T = 128, 12 heads x 8, masked keys
absent), 1024 coalition rows with masks from the Shapley sampler; also head dims 16 / 32.  Never imports oracle/.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autognothi_b200 import ops  # noqa: E402
from autognothi_b200.models import shapley  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    which = "cuda-core" if os.environ.get("AGB_NARROW_SIMT", "0") not in ("", "0") else "mma.sync"
    print(f"kernel: {which}")
    for T, heads, d, mode in [(197, 12, 8, 0), (128, 12, 8, 1), (197, 12, 16, 0), (197, 12, 32, 0), (512, 12, 8, 1)]:
        rows, H = 1024, heads * d
        qkv = (torch.randn(rows * T, 3 * H, device=dev) * 1.5).to(torch.bfloat16)
        dense = shapley.mask_shapley_new(rows, T - 1, device=dev)
        masks = ops.pack_masks(dense, prepend_cls=True)
        ref = ops.masked_attention(qkv[:32 * T].float().contiguous(), masks[:32].contiguous(), T, heads, mode)
        got = ops.masked_attention(qkv, masks, T, heads, mode)
        err = (got[:32 * T].float() - ref).abs().max().item()
        rel = ((got[:32 * T].float() - ref).norm() / ref.norm()).item()
        for _ in range(3):
            ops.masked_attention(qkv, masks, T, heads, mode)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize()
        ev[0].record()
        iters = 20
        for _ in range(iters):
            ops.masked_attention(qkv, masks, T, heads, mode)
        ev[1].record()
        torch.cuda.synchronize()
        us = ev[0].elapsed_time(ev[1]) * 1e3 / iters
        kept = float(dense.float().mean()) * (T - 1) + 1
        print(f"T={T:4d} heads={heads} d={d:2d} mode={mode} rows={rows}: {us:8.1f} us/launch   max|err|={err:.2e} relL2={rel:.2e}"
              f"   mean kept keys {kept:.0f}")


if __name__ == "__main__":
    main()
