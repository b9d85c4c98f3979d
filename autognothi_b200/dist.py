"""Multi-GPU partitioning of the hot path: one process per GPU, `torch.distributed` for the
plumbing.  Masked evaluation shards inputs (an input's S coalitions stay on one rank) and needs only a final
gather; explainer training is data-parallel with one bucketed gradient all-reduce (NCCL over NVLink on the
GPU box; the same code runs over gloo on CPU tensors in the tests)."""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Tuple

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


def broadcast_parameters(params: Iterable[Tensor], src: int = 0) -> None:
    """Make every rank start from rank `src`'s parameter values (layers the recipes add with New() are randomly initialised
    per process; without this, data-parallel replicas diverge silently unless every rank seeds identically)."""
    rank, ws = world()
    if ws == 1:
        return
    for p in params:
        dist.broadcast(p.data if isinstance(p, torch.nn.Parameter) else p, src=src)


def parameter_checksum(params: Iterable[Tensor]) -> float:
    """Order-dependent fp64 checksum of the parameter values (cheap replica-consistency check)."""
    acc = 0.0
    for i, p in enumerate(params):
        acc += float(p.detach().double().sum()) * (1.0 + 1e-3 * (i % 97))
    return acc


def assert_replicas_identical(params: Iterable[Tensor]) -> None:
    rank, ws = world()
    if ws == 1:
        return
    params = list(params)
    dev = params[0].device if params else torch.device("cpu")
    mine = torch.tensor([parameter_checksum(params)], dtype=torch.float64, device=dev)
    allv = [torch.empty_like(mine) for _ in range(ws)]
    dist.all_gather(allv, mine)
    vals = [float(v) for v in allv]
    if any(v != vals[0] for v in vals):
        raise RuntimeError(f"data-parallel replicas hold different parameters (checksums per rank: {vals}); "
                           "call broadcast_parameters() after building the model")


class OverlappedGradReducer:
    """Gradient averaging that runs WHILE the backward pass is still going (reference loop: scripts/train_explainer.py:197-198,
    `loss.backward(); optimizer.step()`).  The hand-written adjoint (training.backward_train) hands over the
    gradients of every block the moment they exist — last block first — through push(); they are copied into pre-allocated
    flat buckets, and a bucket that is full goes out as ONE asynchronous all-reduce (NCCL over NVLink on its own stream)
    while the main stream continues with the next block's adjoint.  finish() flushes the last bucket, waits for the
    outstanding work and returns the averaged gradients as views of the buckets (no torch.cat, no copy back).

    wire_dtype = torch.bfloat16 halves the bytes on the wire (the sum is formed by NCCL in bf16: ~3 significant digits per
    gradient element, averaged over ranks); torch.float32 (default) is exact averaging.
    The bucket layout is fixed by the order of the first backward pass, which is the same code path on every rank."""

    def __init__(self, bucket_mb: float = 32.0, wire_dtype: torch.dtype = torch.float32):
        assert wire_dtype in (torch.float32, torch.bfloat16)
        self.bucket_bytes = max(16, int(bucket_mb * 1024 * 1024))
        self.wire_dtype = wire_dtype
        self.layout: Optional[List[Tuple[str, int, int, int, torch.Size]]] = None   # (name, bucket, offset, numel, shape)
        self.bucket_elems: List[int] = []
        self.flat: List[Tensor] = []
        self.recording: List[Tuple[str, int, torch.Size]] = []
        self._reset()
        self.stats = {"buckets": 0, "bytes": 0}

    def _reset(self) -> None:
        self.cursor = 0                  # index into layout of the next expected gradient
        self.launched = -1               # last bucket already handed to the collective
        self.work: List = []
        self.pushed: Dict[str, Tensor] = {}

    # -- layout ------------------------------------------------------------------------------------------------
    def _build_layout(self, device) -> None:
        esize = 4 if self.wire_dtype == torch.float32 else 2
        cap = max(1, self.bucket_bytes // esize)
        layout, sizes, b, off = [], [], 0, 0
        for name, numel, shape in self.recording:
            if off > 0 and off + numel > cap:
                sizes.append(off)
                b, off = b + 1, 0
            layout.append((name, b, off, numel, shape))
            off += (numel + 3) // 4 * 4       # keep every view 16-byte aligned
        sizes.append(off)
        self.layout, self.bucket_elems = layout, sizes
        self.flat = [torch.zeros((n,), dtype=self.wire_dtype, device=device) for n in sizes]
        self.index = {name: i for i, (name, *_rest) in enumerate(layout)}

    def _launch(self, b: int) -> None:
        rank, ws = world()
        if ws == 1:
            return
        buf = self.flat[b]
        if dist.get_backend() == "nccl":
            self.work.append((dist.all_reduce(buf, op=dist.ReduceOp.AVG, async_op=True), b, False))
        else:                              # gloo (CPU tests): no AVG, no bf16 reduction
            if buf.dtype == torch.bfloat16:
                tmp = buf.float()
                self.work.append((dist.all_reduce(tmp, op=dist.ReduceOp.SUM, async_op=True), b, tmp))
            else:
                self.work.append((dist.all_reduce(buf, op=dist.ReduceOp.SUM, async_op=True), b, True))
        self.stats["buckets"] += 1
        self.stats["bytes"] += buf.numel() * buf.element_size()

    # -- called by the adjoint -----------------------------------------------------------------------------------
    def push(self, grads: Dict[str, Tensor], names: Sequence[str]) -> None:
        """The gradients `names` of this backward pass now exist in `grads` (fp32 tensors)."""
        if self.layout is None:            # first pass: record the order, reduce at finish()
            for n in names:
                self.recording.append((n, grads[n].numel(), grads[n].shape))
                self.pushed[n] = grads[n]
            return
        for n in names:
            i = self.index[n]
            assert i == self.cursor, f"gradient order changed between steps: got {n}, expected {self.layout[self.cursor][0]}"
            _, b, off, numel, _shape = self.layout[i]
            self.flat[b][off:off + numel].copy_(grads[n].reshape(-1))
            self.cursor += 1
            # every gradient of the buckets before b is in place: send them
            while self.launched < b - 1:
                self.launched += 1
                self._launch(self.launched)

    def finish(self, grads: Dict[str, Tensor]) -> Dict[str, Tensor]:
        """-> {name: averaged gradient}, fp32, shaped like the parameter; call once per backward pass."""
        rank, ws = world()
        if self.layout is None:
            dev = next(iter(self.pushed.values())).device
            self._build_layout(dev)
            pushed, self.pushed = self.pushed, {}
            self.cursor = 0
            self.push(pushed, [n for n, *_ in self.layout])
        assert self.cursor == len(self.layout), "backward pass ended before every gradient was pushed"
        while self.launched < len(self.flat) - 1:
            self.launched += 1
            self._launch(self.launched)
        for work, b, aux in self.work:
            work.wait()
            if aux is True:
                self.flat[b].div_(ws)
            elif aux is not False:
                self.flat[b].copy_(aux.div_(ws))
        out = dict(grads)
        for name, b, off, numel, shape in self.layout:
            v = self.flat[b][off:off + numel].reshape(shape)
            out[name] = v if v.dtype == torch.float32 else v.float()
        self._reset()
        return out


class GradAllReducer:
    """Averages `.grad` of the given parameters across ranks in fixed-size flat buckets.
    Each bucket is one all-reduce; buckets are issued back to back (async) so NCCL pipelines them over NVLink,
    then unpacked.  With world size 1 this is a no-op.
    attach(model): the model's hand-written adjoint averages its gradients DURING backward through an
    OverlappedGradReducer instead (model.agb_grad_reducer); allreduce() then only covers parameters whose gradients did not
    come out of that adjoint."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_mb: float = 64.0, broadcast: bool = True):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.overlapped: Optional[OverlappedGradReducer] = None
        if broadcast:
            broadcast_parameters(self.params, src=0)
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

    def attach(self, model, bucket_mb: float = 32.0, wire_dtype: torch.dtype = torch.float32) -> "OverlappedGradReducer":
        self.overlapped = OverlappedGradReducer(bucket_mb=bucket_mb, wire_dtype=wire_dtype)
        model.agb_grad_reducer = self.overlapped
        return self.overlapped

    def allreduce(self) -> None:
        rank, ws = world()
        if ws == 1 or self.overlapped is not None:
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
