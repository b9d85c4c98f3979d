import sys, torch
sys.path.insert(0, "/root/repo")
from autognothi_b200 import ops
dev = torch.device("cuda:0")
d, S, C, B = 128, 2048, 2, 592
dense = (torch.rand(B * S, d - 1, device=dev) > 0.5).to(torch.int64)
Zp = ops.pack_masks(dense, prepend_cls=True).reshape(B, S, -1)
w = torch.rand(B, S, device=dev) + 0.1
probs = torch.rand(B, S, C, device=dev) * 0.9 + 0.05
fx = torch.rand(B, C, device=dev) * 0.9 + 0.05
fnull = torch.rand(C, device=dev) * 0.9 + 0.05
for _ in range(2):
    ops.kernelshap_solve(Zp, w, probs, fx, fnull, d)
torch.cuda.synchronize()
