#!/usr/bin/env python
"""ncu case for the attention kernels at the bench shape (test infrastructure):
    ncu --set full --clock-control none --import-source on -k regex:attention_ -s 4 -c 4 -o gpurun_out/attn python tools/attn_ncu_case.py
launches, after one warm-up of each: pipelined token order, pipelined kept-first, split token order, split kept-first."""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from autognothi_b200 import _native as nat, ops  # noqa: E402
from tools.attn_split_bench import T, heads, setup  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
qkv, masks, qkv_p, nkeep, order = setup(rows)
for rep in range(2):
    for variant in (2, 3):
        prev = nat.lib.agb_attention_set_variant(variant)
        ops.masked_attention(qkv, masks, T, heads, ops.MASK_MUL0)
        ops.attention_prefix(qkv_p, nkeep, T, heads)
        nat.lib.agb_attention_set_variant(prev)
torch.cuda.synchronize()
