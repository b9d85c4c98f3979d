#!/usr/bin/env python
"""Cross-check the tcgen05 attention adjoint against the CUDA-core one (bf16 and fp32-exact) and time both.
Test infrastructure; run under gpurun:  python tools/attn_bwd_check.py [rows]"""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from autognothi_b200 import _native as nat, ops  # noqa: E402


def timed(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / iters * 1e3


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    ok = True
    for name, T, mode in (("vit", 197, ops.MASK_MUL0), ("bert", 128, ops.MASK_NEGINF), ("bert250", 250, ops.MASK_NEGINF),
                          ("vit70", 70, ops.MASK_MUL0), ("vit130", 130, ops.MASK_MUL0)):
        heads, H = 12, 768
        qkv = (torch.randn(rows * T, 3 * H, device=dev)).to(torch.bfloat16)
        dctx = (torch.randn(rows * T, H, device=dev) * 0.1).to(torch.bfloat16)
        dense = (torch.rand(rows, T, device=dev) > 0.5).to(torch.int64)
        dense[:, 0] = 1
        dense[0, 1:] = 0
        if rows > 1:
            dense[1, :] = 1
        masks = ops.pack_masks(dense[:, 1:].contiguous(), prepend_cls=True)
        nat.lib.agb_attention_bwd_set_variant(1)
        simt, us1 = timed(lambda: ops.masked_attention_bwd(qkv, dctx, masks, T, heads, mode))
        # fp32 CUDA-core reference (its smem staging limits it to T <= 200; beyond that compare with the bf16 CUDA-core kernel)
        ref = ops.masked_attention_bwd(qkv.float(), dctx.float(), masks, T, heads, mode) if T <= 200 else simt.float()
        nat.lib.agb_attention_bwd_set_variant(0)
        tc, us0 = timed(lambda: ops.masked_attention_bwd(qkv, dctx, masks, T, heads, mode))
        for part, sl in (("dQ", slice(0, H)), ("dK", slice(H, 2 * H)), ("dV", slice(2 * H, 3 * H))):
            r = ref[:, sl]
            e_tc = (tc[:, sl].float() - r).norm() / r.norm()
            e_si = (simt[:, sl].float() - r).norm() / r.norm()
            good = bool(torch.isfinite(tc.float()).all()) and e_tc.item() < 2e-2
            ok &= good
            print(f"{name:8s} T={T:3d} {part}: rel-L2 vs fp32  tc {e_tc.item():.3e}  simt-bf16 {e_si.item():.3e}  {'OK' if good else 'BAD'}")
        print(f"{name:8s} rows={rows}: simt {us1:9.1f} us   tcgen05 {us0:9.1f} us   speed-up {us1 / us0:5.1f}x")
    print("ALL OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
