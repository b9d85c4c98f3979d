import sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
from autognothi_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
for T, heads, rows in ((197, 12, 24), (197, 2, 7), (128, 3, 5), (65, 1, 4), (256, 2, 3)):
    H = heads * 64
    qkv = (torch.randn(rows * T, 3 * H, device=dev) * 1.0).to(torch.bfloat16)
    nk = torch.randint(1, T + 1, (rows,), device=dev, dtype=torch.int32)
    nk[0] = T
    nk[1] = 1
    if rows > 2: nk[2] = T - 1
    if rows > 3: nk[3] = 2
    dense = (torch.arange(T - 1, device=dev)[None, :] < (nk[:, None] - 1)).to(torch.int64)
    masks = ops.pack_masks(dense, prepend_cls=True)
    ref32 = ops.masked_attention(qkv.float().contiguous(), masks, T, heads, ops.MASK_MUL0)
    ref16 = ops.masked_attention(qkv, masks, T, heads, ops.MASK_MUL0)
    got = ops.attention_prefix(qkv, nk.contiguous(), T, heads)
    torch.cuda.synchronize()
    e_new = float((got.float() - ref32).abs().max()); e_old = float((ref16.float() - ref32).abs().max())
    print(f"T={T} heads={heads} rows={rows}: max|prefix - fp32| = {e_new:.4f}   max|tcgen05 mul0 - fp32| = {e_old:.4f}   scale {float(ref32.abs().max()):.3f}")
# timing at bench shape
T, heads, rows = 197, 12, 1024
H = 768
qkv = torch.randn(rows * T, 3 * H, device=dev).to(torch.bfloat16)
from autognothi_b200.models import shapley as ash
pm = ash.mask_shapley_new(rows, T - 1, device=dev, rng="philox", seed=1, packed=True)
dense = pm.dense()
nk = (dense.sum(1) + 1).to(torch.int32).contiguous()
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / it * 1e3
print("mul0  ", t(lambda: ops.masked_attention(qkv, pm.words, T, heads, ops.MASK_MUL0)), "us")
print("prefix", t(lambda: ops.attention_prefix(qkv, nk, T, heads)), "us   mean kept", float(nk.float().mean()))
