"""Diagnostic probes for the tcgen05 attention kernel (test infrastructure)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from autognothi_b200 import ops
torch.manual_seed(0)
DEV = "cuda"

def run(T, heads, rows, q, k, v, mask, mode, tag):
    H = heads * 64
    qkv = torch.cat([q, k, v], dim=-1).to(DEV).bfloat16().reshape(rows * T, 3 * H).contiguous()
    words = (T + 31) // 32
    pm = ops.pack_masks(mask.to(DEV).to(torch.int64), prepend_cls=False)
    out = ops.masked_attention(qkv, pm, T, heads, mode).float().reshape(rows, T, H)
    ref = ops.masked_attention(qkv, pm, T, heads, mode, force_simt=True).float().reshape(rows, T, H)
    err = (out - ref).abs()
    print(f"[{tag}] T={T} heads={heads} rows={rows} mode={mode}: max err {err.max().item():.4f} mean {err.mean().item():.5f} finite={torch.isfinite(out).all().item()}")
    return out.cpu(), ref.cpu(), err.cpu()

for T in (128, 197, 17):
    rows, heads = 1, 1
    H = 64
    ones = torch.ones(rows, T, dtype=torch.int64)
    z = torch.zeros(rows, T, H)
    # A: Q=0 -> uniform P; V random -> ctx = mean over keys
    v = torch.randn(rows, T, H)
    out, ref, err = run(T, heads, rows, z, torch.randn(rows, T, H), v, ones, 0, "A q=0")
    print("   out[0,0,:6]", out[0, 0, :6].numpy(), " ref", ref[0, 0, :6].numpy())
    print("   err by d-col max:", err[0].max(dim=0).values[:16].numpy())
    # B: V[j,c] = j/T
    vj = (torch.arange(T).float() / T).reshape(1, T, 1).expand(rows, T, H).contiguous()
    out, ref, err = run(T, heads, rows, z, z, vj, ones, 0, "B v=j/T")
    print("   out[0,:4,0]", out[0, :4, 0].numpy(), " ref", ref[0, :4, 0].numpy())
    # C: one-hot V reveals P for key blocks
    q, k = torch.randn(rows, T, H), torch.randn(rows, T, H)
    for blk in range((T + 63) // 64):
        voh = torch.zeros(rows, T, H)
        for c in range(64):
            j = blk * 64 + c
            if j < T:
                voh[0, j, c] = 1.0
        out, ref, err = run(T, heads, rows, q, k, voh, ones, 0, f"C onehot blk{blk}")
        bad = (err[0] > 0.01).nonzero()
        print("   #bad", len(bad), " first bad (query,keycol):", bad[:8].tolist())
        print("   row0 out", out[0, 0, :8].numpy(), "\n   row0 ref", ref[0, 0, :8].numpy())
        print("   err per query row (max) first 8:", err[0].max(dim=1).values[:8].numpy(), " last 4:", err[0].max(dim=1).values[-4:].numpy())
    # D: masked
    mask = (torch.rand(rows, T) > 0.5).long(); mask[:, 0] = 1
    for mode in (0, 1):
        out, ref, err = run(T, heads, rows, q, k, torch.randn(rows, T, H), mask, mode, f"D masked mode{mode}")
