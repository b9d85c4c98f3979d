#!/usr/bin/env python
"""Time agb_masked_attention_bf16 (both kernel generations) in isolation and cross-check them against the
fp32 CUDA-core kernel.  Test infrastructure; run under gpurun:  python tools/attn_bench.py [rows]"""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from autognothi_b200 import _native as nat, ops  # noqa: E402


def run(rows, T, heads, mode, variant, qkv, masks, iters=10):
    nat.lib.agb_attention_set_variant(variant)
    out = ops.masked_attention(qkv, masks, T, heads, mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        ops.masked_attention(qkv, masks, T, heads, mode)
    e0.record()
    for _ in range(iters):
        ops.masked_attention(qkv, masks, T, heads, mode)
    e1.record()
    torch.cuda.synchronize()
    nat.lib.agb_attention_set_variant(0)
    return out, e0.elapsed_time(e1) / iters * 1e3


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    for name, T, mode in (("vit", 197, ops.MASK_MUL0), ("bert", 128, ops.MASK_NEGINF), ("bert250", 250, ops.MASK_NEGINF), ("bert512", 512, ops.MASK_NEGINF), ("vit300", 300, ops.MASK_MUL0),
                          ("vit70", 70, ops.MASK_MUL0)):
        heads, H = 12, 768
        qkv = (torch.randn(rows * T, 3 * H, device=dev) * 1.5).to(torch.bfloat16)
        dense = (torch.rand(rows, T, device=dev) > 0.5).to(torch.int64)
        dense[:, 0] = 1
        dense[0, 1:] = 0          # only CLS kept
        dense[1, :] = 1           # everything kept
        masks = ops.pack_masks(dense[:, 1:].contiguous(), prepend_cls=True)
        ref = ops.masked_attention(qkv[:64 * T].float(), masks[:64], T, heads, mode).float()
        flops = 4.0 * rows * T * T * H
        for v in ((1, 0) if T <= 256 else (0,)):
            out, us = run(rows, T, heads, mode, v, qkv, masks)
            err = (out[:64 * T].float() - ref).abs().max().item()
            print(f"{name:8s} T={T:3d} rows={rows} variant={v}: {us:9.1f} us  {flops / us * 1e-6:7.1f} TFLOP/s  "
                  f"max|err| vs fp32 simt (first 64 rows) {err:.3e}  finite={bool(torch.isfinite(out.float()).all())}")




def trace(T=197, mode=None, rows=1024):
    """Pipeline timeline of CTA 0 (clock64 deltas, cycles): python tools/attn_bench.py trace [T]"""
    import ctypes
    dev = torch.device("cuda:0")
    heads, H = 12, 768
    mode = ops.MASK_MUL0 if mode is None else mode
    qkv = (torch.randn(rows * T, 3 * H, device=dev) * 1.5).to(torch.bfloat16)
    dense = (torch.rand(rows, T, device=dev) > 0.5).to(torch.int64)
    masks = ops.pack_masks(dense[:, 1:].contiguous(), prepend_cls=True)
    buf = torch.zeros((64, 8), dtype=torch.int64, device=dev)
    for _ in range(2):
        ops.masked_attention(qkv, masks, T, heads, mode)
    nat.lib.agb_attention_set_trace(ctypes.c_void_p(buf.data_ptr()))
    ops.masked_attention(qkv, masks, T, heads, mode)
    torch.cuda.synchronize()
    nat.lib.agb_attention_set_trace(None)
    t = buf.cpu()
    t0 = int(t[t > 0].min())
    print("item  kv_load  S_issued  P_seen  PV_issued | s_full_seen  p_arrive  o_full_seen  o_free   (cycles since start)")
    for k in range(40):
        row = [int(v) - t0 if int(v) > 0 else -1 for v in t[k]]
        print(f"{k:4d} {row[0]:8d} {row[1]:9d} {row[2]:7d} {row[3]:10d} | {row[4]:11d} {row[5]:9d} {row[6]:12d} {row[7]:7d}")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "trace":
    trace(int(sys.argv[2]) if len(sys.argv) > 2 else 197,
          ops.MASK_NEGINF if (len(sys.argv) > 3 and sys.argv[3] == "bert") else ops.MASK_MUL0)
elif __name__ == "__main__":
    main()
