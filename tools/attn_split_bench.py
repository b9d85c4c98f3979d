#!/usr/bin/env python
"""Third-generation (split-softmax) attention kernel against the pipelined one at the bench shape: ViT-Base/16, 1024 rows x 12
heads, T = 197, coalition masks from the Shapley-kernel sampler.  Test infrastructure; run under gpurun:
    python tools/attn_split_bench.py [rows]          timings + max error against the fp32 CUDA-core kernel
    python tools/attn_split_bench.py trace [prefix]  pipeline timeline of CTA 0 (clock64 deltas)"""
import ctypes
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from autognothi_b200 import _native as nat, ops  # noqa: E402
from autognothi_b200.models import shapley as ash  # noqa: E402

dev = torch.device("cuda:0")
T, heads, H = 197, 12, 768


def timeit(fn, it=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3


def setup(rows):
    torch.manual_seed(0)
    qkv = (torch.randn(rows * T, 3 * H, device=dev) * 1.5).to(torch.bfloat16)
    pm = ash.mask_shapley_new(rows, T - 1, device=dev, rng="philox", seed=1, packed=True)
    order, pos, nkeep, prefix = ops.kept_first_order(pm.words, T)
    r = torch.arange(rows, device=dev)
    qkv_p = qkv.reshape(rows, T, 3 * H)[r[:, None], order.long()].reshape(rows * T, 3 * H).contiguous()
    return qkv, pm.words, qkv_p, nkeep, order


def main(rows):
    qkv, masks, qkv_p, nkeep, order = setup(rows)
    nref = min(rows, 64)
    ref = ops.masked_attention(qkv[:nref * T].float(), masks[:nref], T, heads, ops.MASK_MUL0).float()
    r = torch.arange(nref, device=dev)
    ref_p = ref.reshape(nref, T, H)[r[:, None], order[:nref].long()].reshape(nref * T, H)
    print(f"rows={rows} T={T} heads={heads}; mean kept tokens {float(nkeep.float().mean()):.1f}; "
          f"rows with <= 16 kept {float((nkeep <= 16).float().mean()):.2f}, >= 192 kept {float((nkeep >= 192).float().mean()):.2f}")
    for variant, name in ((2, "pipelined (gen 2)"), (3, "split softmax (gen 3)")):
        prev = nat.lib.agb_attention_set_variant(variant)
        out = ops.masked_attention(qkv, masks, T, heads, ops.MASK_MUL0)
        outp = ops.attention_prefix(qkv_p, nkeep, T, heads)
        e1 = float((out[:nref * T].float() - ref).abs().max())
        e2 = float((outp[:nref * T].float() - ref_p).abs().max())
        t1 = timeit(lambda: ops.masked_attention(qkv, masks, T, heads, ops.MASK_MUL0))
        t2 = timeit(lambda: ops.attention_prefix(qkv_p, nkeep, T, heads))
        nat.lib.agb_attention_set_variant(prev)
        print(f"{name:24s} token order {t1:7.1f} us (max|err| {e1:.3e})   kept-first order {t2:7.1f} us (max|err| {e2:.3e})   "
              f"finite={bool(torch.isfinite(out.float()).all() and torch.isfinite(outp.float()).all())}")


def trace(prefix, rows=1024):
    qkv, masks, qkv_p, nkeep, order = setup(rows)
    buf = torch.zeros((64, 16), dtype=torch.int64, device=dev)
    fn = (lambda: ops.attention_prefix(qkv_p, nkeep, T, heads)) if prefix else (lambda: ops.masked_attention(qkv, masks, T, heads, ops.MASK_MUL0))
    nat.lib.agb_attention_set_variant(3)
    for _ in range(2):
        fn()
    nat.lib.agb_attention_set_trace(ctypes.c_void_p(buf.data_ptr()))
    fn()
    torch.cuda.synchronize()
    nat.lib.agb_attention_set_trace(None)
    t = buf.cpu()
    t0 = int(t[t > 0].min())
    print(f"# split-softmax kernel, {'kept-first order' if prefix else 'token order'}; rows of CTA 0: "
          + " ".join(str(int(nkeep[(k * 148) // heads])) for k in range(20)) + " kept tokens per unit")
    print("item  kv_load  S_issued  P_seen  PV_issued | s_full_seen  p_arrive  o_full_seen  o_free | softmax warp (quarter 0, half 0): "
          "first_chunk_landed  max_exchanged  chunks_done  halves_reconciled  stores_done   (cycles since start)")
    for k in range(40):
        row = [int(v) - t0 if int(v) > 0 else -1 for v in t[k]]
        print(f"{k:4d} {row[0]:8d} {row[1]:9d} {row[2]:7d} {row[3]:10d} | {row[4]:11d} {row[5]:9d} {row[6]:12d} {row[7]:7d} | "
              f"{row[8]:8d} {row[9]:8d} {row[10]:8d} {row[11]:8d} {row[12]:8d} | upper half: s_full_seen {row[13]:8d} arrive {row[15]:8d}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "trace":
        trace(len(sys.argv) > 2 and sys.argv[2] == "prefix")
    else:
        main(int(sys.argv[1]) if len(sys.argv) > 1 else 1024)
