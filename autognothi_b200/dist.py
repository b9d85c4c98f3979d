"""Multi-GPU partitioning of the hot path: one process per GPU, `torch.distributed` for the
plumbing.  Masked evaluation shards inputs (an input's S coalitions stay on one rank) and needs only a final
gather; explainer training is data-parallel with one bucketed gradient all-reduce (NCCL over NVLink on the
GPU box; the same code runs over gloo on CPU tensors in the tests)."""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced partition [lo, hi) of n_items; the first (n_items % world) ranks get one extra."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_rows(local: Tensor, counts: Sequence[int]) -> Tensor:
    """All ranks contribute `local` (counts[rank] rows); returns the concatenation in rank order on every rank.
    Ragged shards are padded to the largest count for the collective and trimmed afterwards."""
    rank, ws = world()
    if ws == 1:
        return local
    assert local.shape[0] == counts[rank]
    m = max(counts)
    buf = local
    if local.shape[0] < m:
        buf = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        buf[: local.shape[0]] = local
    outs = [torch.empty_like(buf) for _ in range(ws)]
    dist.all_gather(outs, buf.contiguous())
    return torch.cat([o[:c] for o, c in zip(outs, counts)], dim=0)


class GradAllReducer:
    """Averages `.grad` of the given parameters across ranks in fixed-size flat buckets.
    Each bucket is one all-reduce; buckets are issued back to back (async) so NCCL pipelines them over NVLink,
    then unpacked.  With world size 1 this is a no-op."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_mb: float = 64.0):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.bucket_elems = max(1, int(bucket_mb * 1024 * 1024 / 4))
        self.buckets: List[List[torch.nn.Parameter]] = []
        cur, n = [], 0
        for p in self.params:
            if cur and n + p.numel() > self.bucket_elems:
                self.buckets.append(cur)
                cur, n = [], 0
            cur.append(p)
            n += p.numel()
        if cur:
            self.buckets.append(cur)

    def allreduce(self) -> None:
        rank, ws = world()
        if ws == 1:
            return
        pending = []
        for bucket in self.buckets:
            grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in bucket]
            flat = torch.cat([g.reshape(-1).float() for g in grads])
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)
            pending.append((work, flat, bucket, grads))
        for work, flat, bucket, grads in pending:
            work.wait()
            flat.div_(ws)
            off = 0
            for p, g in zip(bucket, grads):
                n = g.numel()
                if p.grad is None:
                    p.grad = flat[off:off + n].reshape(p.shape).clone()
                else:
                    p.grad.copy_(flat[off:off + n].reshape(p.shape))
                off += n
