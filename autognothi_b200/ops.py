"""Tensor-level wrappers over the C-ABI (device memory and streams come from PyTorch; the arithmetic
is ours).  Every function launches on torch's current CUDA stream and checks shapes/dtypes before
handing raw pointers across the boundary."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _native as nat

MASK_MUL0 = 0
MASK_NEGINF = 1
ACT_NONE = 0
ACT_GELU = 1


def _c(t: Tensor) -> Tensor:
    return t if t.is_contiguous() else t.contiguous()


def mask_words(n_tokens: int) -> int:
    return (n_tokens + 31) // 32


# ------------------------------------------------------------------------------------------------
# masks
# ------------------------------------------------------------------------------------------------
def pack_masks(mask: Tensor, prepend_cls: bool = True) -> Tensor:
    """(rows, n) int64 {0,1} player masks -> (rows, words) packed int32 (bit 0 = CLS)."""
    assert mask.dim() == 2 and mask.dtype == torch.int64, "mask must be (rows, n_players) int64"
    mask = _c(mask)
    rows, n = mask.shape
    words = mask_words(n + (1 if prepend_cls else 0))
    out = torch.empty((rows, words), dtype=torch.int32, device=mask.device)
    nat.call("agb_pack_masks_i64", nat.ptr(mask), rows, n, 1 if prepend_cls else 0, nat.ptr(out), words, nat.stream())
    return out


def unpack_masks(packed: Tensor, n: int, skip: int = 1) -> Tensor:
    packed = _c(packed)
    rows, words = packed.shape
    out = torch.empty((rows, n), dtype=torch.int64, device=packed.device)
    nat.call("agb_unpack_masks_i64", nat.ptr(packed), rows, n, skip, words, nat.ptr(out), nat.stream())
    return out


def shapley_masks(prefix: Tensor, pairs: int, n_players: int, *, u_players: Optional[Tensor] = None,
                  u_size: Optional[Tensor] = None, seed: int = 0, offset: int = 0,
                  want_dense: bool = False) -> Tuple[Tensor, Optional[Tensor]]:
    """Paired Shapley-kernel sampler -> (packed (2*pairs, words), dense (2*pairs, n) int64 | None)."""
    dev = prefix.device
    words = mask_words(n_players + 1)
    packed = torch.empty((2 * pairs, words), dtype=torch.int32, device=dev)
    dense = torch.empty((2 * pairs, n_players), dtype=torch.int64, device=dev) if want_dense else None
    use_philox = u_players is None
    if not use_philox:
        assert u_players.shape == (pairs, n_players) and u_size.numel() == pairs
        u_players, u_size = _c(u_players.float()), _c(u_size.float().reshape(-1))
    nat.call("agb_shapley_masks", nat.ptr(u_players), nat.ptr(u_size), nat.ptr(_c(prefix)), 1 if use_philox else 0,
             seed, offset, pairs, n_players, nat.ptr(packed), words, nat.ptr(dense), nat.stream())
    return packed, dense


def uniform_masks(rows: int, n_players: int, device, *, u_players: Optional[Tensor] = None,
                  u_row: Optional[Tensor] = None, seed: int = 0, offset: int = 0,
                  want_dense: bool = False) -> Tuple[Tensor, Optional[Tensor]]:
    words = mask_words(n_players + 1)
    packed = torch.empty((rows, words), dtype=torch.int32, device=device)
    dense = torch.empty((rows, n_players), dtype=torch.int64, device=device) if want_dense else None
    use_philox = u_players is None
    if not use_philox:
        u_players, u_row = _c(u_players.float()), _c(u_row.float().reshape(-1))
    nat.call("agb_uniform_masks", nat.ptr(u_players), nat.ptr(u_row), 1 if use_philox else 0, seed, offset, rows,
             n_players, nat.ptr(packed), words, nat.ptr(dense), nat.stream())
    return packed, dense


def rank_masks(scores: Optional[Tensor], stops: Tensor, n_players: int, mask_base: int, *, rows: Optional[int] = None,
               device=None, seed: int = 0, offset: int = 0, want_dense: bool = False) -> Tuple[Tensor, Optional[Tensor]]:
    """scores (rows, n) fp32 (None -> Philox keys on device) + stops int32 (nstops,) -> packed (rows*nstops, words)
    [, dense int64 (rows*nstops, n)]: row r*nstops+i = mask_base with the stops[i] top-ranked players flipped."""
    use_philox = scores is None
    if not use_philox:
        scores = _c(scores.float())
        assert scores.dim() == 2 and scores.shape[1] == n_players
        rows, device = scores.shape[0], scores.device
    assert rows is not None and device is not None
    stops = _c(stops.to(device=device, dtype=torch.int32))
    nstops = stops.numel()
    words = mask_words(n_players + 1)
    packed = torch.empty((rows * nstops, words), dtype=torch.int32, device=device)
    dense = torch.empty((rows * nstops, n_players), dtype=torch.int64, device=device) if want_dense else None
    nat.call("agb_rank_masks", nat.ptr(scores), 1 if use_philox else 0, seed, offset, rows, n_players, nat.ptr(stops), nstops,
             1 if mask_base else 0, nat.ptr(packed), words, nat.ptr(dense), nat.stream())
    return packed, dense


# ------------------------------------------------------------------------------------------------
# dense
# ------------------------------------------------------------------------------------------------
def gemm_bf16(a: Tensor, w: Tensor, bias: Optional[Tensor] = None, *, act: int = ACT_NONE,
              residual: Optional[Tensor] = None, out_dtype=torch.bfloat16, out: Optional[Tensor] = None,
              a_mn: bool = False, w_mn: bool = False, alpha: float = 1.0,
              res_group: int = 0, res_rows: int = 0) -> Tensor:
    """out[M,N] = act(alpha * A @ W^T + bias) + residual.   K-major: a [M,K], w [N,K] (nn.Linear layout).
    MN-major (a_mn / w_mn): the operand is stored transposed, a [K,M] / w [K,N]."""
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.dim() == 2 and w.dim() == 2
    assert a.stride(1) == 1 and w.stride(1) == 1
    M, K = (a.shape[1], a.shape[0]) if a_mn else a.shape
    N, Kw = (w.shape[1], w.shape[0]) if w_mn else w.shape
    assert K == Kw, f"inner dims differ: {K} vs {Kw}"
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    assert out.shape == (M, N) and out.stride(1) == 1
    res_bf16 = residual if (residual is not None and residual.dtype == torch.bfloat16) else None
    res_f32 = residual if (residual is not None and residual.dtype == torch.float32) else None
    if residual is not None:
        assert residual.stride(-1) == 1 and (res_bf16 is not None or res_f32 is not None)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
    ldr = residual.stride(0) if residual is not None else 0
    nat.NEXT_META = 2.0 * M * N * K
    nat.NEXT_INFO = (f"M{M} N{N} K{K}" + (" +res" if residual is not None else "") + (" gelu" if act else ""),
                     2.0 * M * K + 2.0 * N * K + M * N * out.element_size() + (M * N * residual.element_size() if residual is not None else 0))
    nat.call("agb_gemm_bf16", nat.ptr(a), a.stride(0), 1 if a_mn else 0, nat.ptr(w), w.stride(0), 1 if w_mn else 0,
             M, N, K, float(alpha), nat.ptr(bias), act, nat.ptr(res_bf16), nat.ptr(res_f32), ldr, res_group, res_rows,
             nat.ptr(out), out.stride(0), 1 if out.dtype == torch.float32 else 0, nat.stream())
    return out


def gemm_stats_parts(N: int) -> int:
    return 2 * ((N + 255) // 256)


def gemm_bf16_fused(a: Tensor, w: Tensor, bias: Optional[Tensor], *, act: int = ACT_NONE, residual: Optional[Tensor] = None,
                    out: Optional[Tensor] = None, out_dtype=torch.bfloat16,
                    ln: Optional[Tuple[Tensor, Tensor, float]] = None, emit_copy_stats: bool = False
                    ) -> Tuple[Tensor, Optional[Tensor], Optional[Tensor]]:
    """LayerNorm-folded GEMM chain (agb_gemm_bf16_fused).  ln = (row stats [M, parts, 2], colsum [N], eps) applies
    Linear(LayerNorm(x)) to the UN-normalised bf16 rows `a` (w pre-scaled by gamma, bias pre-shifted by W beta).
    emit_copy_stats (fp32 residual epilogue) also returns a bf16 copy of the output and its row statistics."""
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    copy16 = stats = None
    if emit_copy_stats:
        assert residual is not None and out.dtype == torch.float32
        copy16 = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
        stats = torch.empty((M, gemm_stats_parts(N), 2), dtype=torch.float32, device=a.device)
    ln_stats = ln_colsum = None
    ln_parts, ln_eps = 0, 0.0
    if ln is not None:
        ln_stats, ln_colsum, ln_eps = ln
        assert ln_stats.dtype == torch.float32 and ln_stats.is_contiguous() and ln_stats.shape[0] == M and ln_stats.shape[2] == 2
        assert ln_colsum.dtype == torch.float32 and ln_colsum.numel() == N
        ln_parts = ln_stats.shape[1]
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.stride(1) == 1
    nat.NEXT_META = 2.0 * M * N * K
    nat.NEXT_INFO = (f"M{M} N{N} K{K}" + (" +res" if residual is not None else "") + (" gelu" if act else "") +
                     (" ln-folded" if ln is not None else "") + (" +copy+stats" if emit_copy_stats else ""),
                     2.0 * M * K + 2.0 * N * K + M * N * out.element_size() + (4.0 * M * N if residual is not None else 0) +
                     (2.0 * M * N + 8.0 * M * gemm_stats_parts(N) if emit_copy_stats else 0) + (8.0 * M * ln_parts if ln is not None else 0))
    nat.call("agb_gemm_bf16_fused", nat.ptr(a), a.stride(0), nat.ptr(w), w.stride(0), M, N, K, nat.ptr(bias), act,
             nat.ptr(residual), residual.stride(0) if residual is not None else 0, nat.ptr(out), out.stride(0),
             1 if out.dtype == torch.float32 else 0, nat.ptr(ln_stats), ln_parts, nat.ptr(ln_colsum), float(ln_eps),
             nat.ptr(copy16), N if copy16 is not None else 0, nat.ptr(stats), nat.stream())
    return out, copy16, stats


def gemm_bf16_hilo(a: Tensor, w: Tensor, bias: Optional[Tensor], x_hi: Tensor, x_lo: Tensor) -> Tensor:
    """Residual GEMM on a hi/lo residual stream (agb_gemm_bf16_hilo): (x_hi + x_lo) += a @ w.T + bias, both bf16 planes
    updated in place; -> per-row partial statistics [M, parts, 2] of the new stream for the LayerNorm-folded consumer."""
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and x_hi.shape == (M, N) and x_lo.shape == (M, N)
    assert x_hi.dtype == torch.bfloat16 and x_lo.dtype == torch.bfloat16 and x_hi.is_contiguous() and x_lo.is_contiguous()
    stats = torch.empty((M, gemm_stats_parts(N), 2), dtype=torch.float32, device=a.device)
    nat.NEXT_META = 2.0 * M * N * K
    nat.NEXT_INFO = (f"M{M} N{N} K{K} +res hi/lo +stats",
                     2.0 * M * K + 2.0 * N * K + 8.0 * M * N + 8.0 * M * gemm_stats_parts(N))
    nat.call("agb_gemm_bf16_hilo", nat.ptr(a), a.stride(0), nat.ptr(w), w.stride(0), M, N, K, nat.ptr(bias), nat.ptr(x_hi),
             nat.ptr(x_lo), N, nat.ptr(stats), nat.stream())
    return stats


def split_hilo(x: Tensor) -> Tuple[Tensor, Tensor]:
    """fp32 -> (hi, lo) bf16 planes with hi + lo == x to 16 significant bits (agb_split_hilo)."""
    x = _c(x)
    assert x.dtype == torch.float32 and x.numel() % 8 == 0
    hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    nat.call("agb_split_hilo", nat.ptr(x), x.numel(), nat.ptr(hi), nat.ptr(lo), nat.stream())
    return hi, lo


def rowstats_cast(x: Tensor) -> Tuple[Tensor, Tensor]:
    """fp32 (rows, H) -> (bf16 copy, stats (rows, 1, 2) = per-row (sum, sum of squares))."""
    assert x.dim() == 2 and x.dtype == torch.float32 and x.stride(1) == 1
    rows, H = x.shape
    out = torch.empty((rows, H), dtype=torch.bfloat16, device=x.device)
    stats = torch.empty((rows, 1, 2), dtype=torch.float32, device=x.device)
    nat.call("agb_rowstats_cast", nat.ptr(x), x.stride(0), rows, H, nat.ptr(out), H, nat.ptr(stats), nat.stream())
    return out, stats


def gemm_f32(a: Tensor, w: Tensor, bias: Optional[Tensor] = None, *, act: int = ACT_NONE,
             residual: Optional[Tensor] = None, out: Optional[Tensor] = None, alpha: float = 1.0,
             a_mn: bool = False, w_mn: bool = False) -> Tensor:
    assert a.dtype == torch.float32 and w.dtype == torch.float32 and a.stride(1) == 1 and w.stride(1) == 1
    M, K = (a.shape[1], a.shape[0]) if a_mn else a.shape
    N, Kw = (w.shape[1], w.shape[0]) if w_mn else w.shape
    assert K == Kw
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    step = 65535 * 64
    for m0 in range(0, M, step):
        m1 = min(M, m0 + step)
        r = residual[m0:m1] if residual is not None else None
        a_sub = a[:, m0:m1] if a_mn else a[m0:m1]
        nat.call("agb_gemm_f32", nat.ptr(a_sub), a.stride(0), 1 if a_mn else 0, nat.ptr(w), w.stride(0),
                 1 if w_mn else 0, m1 - m0, N, K, float(alpha), nat.ptr(bias), act, nat.ptr(r),
                 r.stride(0) if r is not None else 0, nat.ptr(out[m0:m1]), out.stride(0), nat.stream())
    return out


def layernorm(x: Tensor, gamma: Tensor, beta: Tensor, eps: float, *, want_bf16: bool, want_f32: bool
              ) -> Tuple[Optional[Tensor], Optional[Tensor]]:
    """LayerNorm over the last dim of a 2-D tensor; returns (bf16 copy | None, fp32 copy | None)."""
    assert x.dim() == 2 and x.stride(1) == 1 and x.dtype in (torch.float32, torch.bfloat16)
    rows, H = x.shape
    ob = torch.empty((rows, H), dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    of = torch.empty((rows, H), dtype=torch.float32, device=x.device) if want_f32 else None
    nat.call("agb_layernorm", nat.ptr(x), 1 if x.dtype == torch.bfloat16 else 0, x.stride(0), rows, H, nat.ptr(gamma),
             nat.ptr(beta), float(eps), nat.ptr(ob), nat.ptr(of), H, nat.stream())
    return ob, of


def to_bf16(x: Tensor) -> Tensor:
    x = _c(x)
    assert x.dtype == torch.float32
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    nat.call("agb_cast_f32_to_bf16", nat.ptr(x), nat.ptr(out), x.numel(), nat.stream())
    return out


def vit_im2col(images: Tensor, patch: int, out_dtype) -> Tensor:
    images = _c(images)
    assert images.dtype == torch.float32 and images.dim() == 4
    B, C, px, px2 = images.shape
    assert px == px2
    g = px // patch
    out = torch.empty((B * g * g, C * patch * patch), dtype=out_dtype, device=images.device)
    nat.call("agb_vit_im2col", nat.ptr(images), B, C, px, patch, nat.ptr(out), 1 if out_dtype == torch.bfloat16 else 0,
             nat.stream())
    return out


def vit_assemble(patch_emb: Tensor, cls_token: Tensor, pos_emb: Tensor, B: int, S: int, T: int, H: int) -> Tensor:
    x = torch.empty((B * S, T, H), dtype=torch.float32, device=patch_emb.device)
    nat.call("agb_vit_assemble", nat.ptr(_c(patch_emb)), nat.ptr(_c(cls_token)), nat.ptr(_c(pos_emb)), B, S, T, H,
             nat.ptr(x), nat.stream())
    return x


def repeat_rows(x: Tensor, S: int) -> Tensor:
    """(B, ...) -> (B*S, ...): S consecutive copies of every leading-dim slice (torch.repeat_interleave(x, S, 0) as one
    vectorised kernel: agb_repeat_rows)."""
    x = _c(x)
    B = x.shape[0]
    row_bytes = (x.numel() // B) * x.element_size() if B else 16
    if S == 1:
        return x.clone()
    if row_bytes % 16 != 0 or x.data_ptr() % 16 != 0:
        return x.repeat_interleave(S, dim=0)
    out = torch.empty((B * S,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    nat.call("agb_repeat_rows", nat.ptr(x), B, row_bytes, S, nat.ptr(out), nat.stream())
    return out


def bert_embed(ids: Tensor, word: Tensor, pos: Tensor, type_emb: Tensor, gamma: Tensor, beta: Tensor, eps: float,
               S: int) -> Tensor:
    ids = _c(ids)
    assert ids.dtype == torch.int64 and ids.dim() == 2
    B, T = ids.shape
    H = word.shape[1]
    x = torch.empty((B * S, T, H), dtype=torch.float32, device=ids.device)
    nat.call("agb_bert_embed", nat.ptr(ids), nat.ptr(word), nat.ptr(pos), nat.ptr(type_emb), nat.ptr(gamma),
             nat.ptr(beta), float(eps), B, S, T, H, word.shape[0], nat.ptr(x), nat.stream())
    return x


def cls_head(x: Tensor, mode: int, w_cls: Tensor, b_cls: Tensor, *, ln: Optional[Tuple[Tensor, Tensor, float]] = None,
             pool: Optional[Tuple[Tensor, Tensor]] = None, want_logits: bool = False):
    """x (rows, T, H) fp32 -> probabilities (rows, C) [, logits]."""
    assert x.dtype == torch.float32 and x.dim() == 3 and x.is_contiguous()
    rows, T, H = x.shape
    C = w_cls.shape[0]
    probs = torch.empty((rows, C), dtype=torch.float32, device=x.device)
    logits = torch.empty((rows, C), dtype=torch.float32, device=x.device) if want_logits else None
    g, b, eps = ln if ln is not None else (None, None, 0.0)
    wp, bp = pool if pool is not None else (None, None)
    nat.call("agb_cls_head", nat.ptr(x), T * H, rows, H, C, mode, nat.ptr(g), nat.ptr(b), float(eps), nat.ptr(wp),
             nat.ptr(bp), nat.ptr(w_cls), nat.ptr(b_cls), nat.ptr(probs), nat.ptr(logits), nat.stream())
    return (probs, logits) if want_logits else probs


def masked_attention(qkv: Tensor, packed_mask: Tensor, T: int, heads: int, mode: int, *, force_simt: bool = False,
                     share: int = 1) -> Tensor:
    """qkv (rows*T, 3H) fused projections -> ctx (rows*T, H), same dtype.  bf16 + head dim 64 + T <= 512 runs
    the tcgen05 kernel; fp32 (the exact mode) and the remaining shapes run the CUDA-core kernel.
    share > 1 (tcgen05 kernel only): qkv is (rows/share*T, 3H) and `share` consecutive mask rows read the same
    projections (the coalitions of one input in the first block)."""
    assert qkv.is_contiguous() and packed_mask.is_contiguous() and packed_mask.dtype == torch.int32
    rows = packed_mask.shape[0]
    assert rows % share == 0 and qkv.shape[0] == (rows // share) * T, "one qkv row block per `share` mask rows"
    H = qkv.shape[1] // 3
    words = packed_mask.shape[1]
    ctx = torch.empty((rows * T, H), dtype=qkv.dtype, device=qkv.device)
    tc_ok = qkv.dtype == torch.bfloat16 and H == heads * 64 and T <= 512 and not force_simt
    if tc_ok:
        nat.NEXT_META = 4.0 * rows * T * T * H
        if share == 1:
            nat.call("agb_masked_attention_bf16", nat.ptr(qkv), nat.ptr(packed_mask), words, rows, T, H, heads, mode,
                     nat.ptr(ctx), nat.stream())
        else:
            nat.call("agb_masked_attention_bf16_shared", nat.ptr(qkv), nat.ptr(packed_mask), words, rows, share, T, H, heads,
                     mode, nat.ptr(ctx), nat.stream())
    else:
        assert share == 1, "shared projections need the tcgen05 attention kernel"
        step = 65535
        for r0 in range(0, rows, step):
            r1 = min(rows, r0 + step)
            nat.call("agb_masked_attention_simt", nat.ptr(qkv[r0 * T:r1 * T]), 1 if qkv.dtype == torch.bfloat16 else 0,
                     nat.ptr(packed_mask[r0:r1]), words, r1 - r0, T, H, heads, mode, nat.ptr(ctx[r0 * T:r1 * T]),
                     nat.stream())
    return ctx


def cls_attention(q: Tensor, kv: Tensor, k_off: int, v_off: int, packed_mask: Tensor, T: int, heads: int, mode: int) -> Tensor:
    """CLS-query attention of the last block: q (rows, H), kv (rows*T, ld) with keys at columns [k_off, k_off+H) and
    values at [v_off, v_off+H) -> ctx (rows, H), same dtype (fp32 math)."""
    assert q.dim() == 2 and kv.dim() == 2 and q.dtype == kv.dtype and q.stride(1) == 1 and kv.stride(1) == 1
    rows, H = q.shape
    assert kv.shape[0] == rows * T and packed_mask.shape[0] == rows and packed_mask.dtype == torch.int32
    ctx = torch.empty((rows, H), dtype=q.dtype, device=q.device)
    step = 65535
    for r0 in range(0, rows, step):
        r1 = min(rows, r0 + step)
        nat.call("agb_cls_attention", nat.ptr(q[r0:r1]), q.stride(0), nat.ptr(kv[r0 * T:r1 * T]), kv.stride(0), k_off, v_off,
                 _is_bf16(q), nat.ptr(_c(packed_mask[r0:r1])), packed_mask.shape[1], r1 - r0, T, H, heads, mode,
                 nat.ptr(ctx[r0:r1]), H, nat.stream())
    return ctx


def pack_kept_tokens(packed_mask: Tensor, T: int, S: int) -> Tuple[Tensor, Tensor, int]:
    """Masked-token dropping (additive masks): -> (cu (rows+1,) int32 exclusive prefix of the kept-token counts,
    src (total,) int64 = row of the per-input (B*T, H) embedding each packed token comes from, total kept tokens).
    One device->host read of the total (buffer sizes and GEMM shapes depend on it)."""
    assert packed_mask.dtype == torch.int32 and packed_mask.is_contiguous()
    rows, words = packed_mask.shape
    counts = torch.empty((rows,), dtype=torch.int32, device=packed_mask.device)
    nat.call("agb_mask_counts", nat.ptr(packed_mask), rows, words, T, nat.ptr(counts), nat.stream())
    cu = torch.zeros((rows + 1,), dtype=torch.int32, device=packed_mask.device)
    torch.cumsum(counts, 0, out=cu[1:])
    total = int(cu[-1].item())
    src = torch.empty((total,), dtype=torch.int64, device=packed_mask.device)
    nat.call("agb_packed_token_index", nat.ptr(packed_mask), rows, words, T, S, nat.ptr(cu), nat.ptr(src), nat.stream())
    return cu, src, total


def attention_varlen(qkv: Tensor, cu: Tensor, max_len: int, heads: int) -> Tensor:
    """qkv (total_tokens, 3H) bf16 packed rows, row r = tokens [cu[r], cu[r+1]) -> ctx (total_tokens, H)."""
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and cu.dtype == torch.int32
    total, H3 = qkv.shape
    H = H3 // 3
    ctx = torch.empty((total, H), dtype=torch.bfloat16, device=qkv.device)
    nat.NEXT_META = None
    nat.call("agb_attention_bf16_varlen", nat.ptr(qkv), nat.ptr(cu), cu.numel() - 1, max_len, total, H, heads, nat.ptr(ctx),
             nat.stream())
    return ctx


def attention_prefix(qkv: Tensor, nkeep: Tensor, T: int, heads: int) -> Tensor:
    """ViT-masked attention for kept-first token order: qkv (rows*T, 3H) bf16, nkeep (rows,) int32 kept tokens per row (CLS
    included) -> ctx (rows*T, H).  The masked keys are folded into one virtual key (agb_attention_bf16_prefix)."""
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and nkeep.dtype == torch.int32 and nkeep.is_contiguous()
    rows = nkeep.shape[0]
    H = qkv.shape[1] // 3
    assert qkv.shape[0] == rows * T
    ctx = torch.empty((rows * T, H), dtype=torch.bfloat16, device=qkv.device)
    nat.NEXT_META = 4.0 * rows * T * T * H       # dense-equivalent FLOPs (the kernel executes fewer: trimmed key range)
    nat.call("agb_attention_bf16_prefix", nat.ptr(qkv), nat.ptr(nkeep), rows, T, H, heads, nat.ptr(ctx), nat.stream())
    return ctx


def kept_first_order(packed_mask: Tensor, T: int) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Stable kept-first partition of every row's tokens (agb_kept_first_order): packed (rows, words) int32 ->
    (order (rows, T) uint8 [position -> token], pos (rows, T) uint8 [token -> position], nkeep (rows,) int32,
    prefix masks (rows, words) int32 of the permuted rows)."""
    assert packed_mask.dtype == torch.int32 and packed_mask.is_contiguous() and T <= 256
    rows, words = packed_mask.shape
    dev = packed_mask.device
    order = torch.empty((rows, T), dtype=torch.uint8, device=dev)
    pos = torch.empty((rows, T), dtype=torch.uint8, device=dev)
    nkeep = torch.empty((rows,), dtype=torch.int32, device=dev)
    prefix = torch.empty((rows, words), dtype=torch.int32, device=dev)
    nat.call("agb_kept_first_order", nat.ptr(packed_mask), rows, words, T, nat.ptr(order), nat.ptr(pos), nat.ptr(nkeep),
             nat.ptr(prefix), nat.stream())
    return order, pos, nkeep, prefix


def gather_token_rows(x: Tensor, order: Tensor, S: int) -> Tensor:
    """x (B, T, H) -> (B*S, T, H) with out[r, q] = x[r // S, order[r, q]] (agb_gather_token_rows)."""
    x = _c(x)
    B, T, H = x.shape
    rows = order.shape[0]
    assert rows == B * S and order.shape[1] == T and order.dtype == torch.uint8 and order.is_contiguous()
    out = torch.empty((rows, T, H), dtype=x.dtype, device=x.device)
    nat.call("agb_gather_token_rows", nat.ptr(x), nat.ptr(order), rows, T, S, H * x.element_size(), nat.ptr(out), nat.stream())
    return out


def gather_token_rows_hilo(x: Tensor, order: Tensor, S: int) -> Tuple[Tensor, Tensor]:
    """gather_token_rows of fp32 rows, written as the (hi, lo) bf16 planes of the hi/lo residual stream."""
    x = _c(x)
    B, T, H = x.shape
    rows = order.shape[0]
    assert x.dtype == torch.float32 and H % 8 == 0
    assert rows == B * S and order.shape[1] == T and order.dtype == torch.uint8 and order.is_contiguous()
    hi = torch.empty((rows, T, H), dtype=torch.bfloat16, device=x.device)
    lo = torch.empty((rows, T, H), dtype=torch.bfloat16, device=x.device)
    nat.call("agb_gather_token_rows_hilo", nat.ptr(x), nat.ptr(order), rows, T, S, H, nat.ptr(hi), nat.ptr(lo), nat.stream())
    return hi, lo


def masked_attention_scatter(qkv: Tensor, packed_mask: Tensor, T: int, heads: int, share: int, pos: Tensor) -> Tensor:
    """masked_attention (ViT masks, bf16 tcgen05 kernel, `share` mask rows per qkv row block) whose query token t of row r
    lands at token position pos[r, t] of ctx (rows*T, H)."""
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and packed_mask.dtype == torch.int32 and packed_mask.is_contiguous()
    rows = packed_mask.shape[0]
    H = qkv.shape[1] // 3
    assert rows % share == 0 and qkv.shape[0] == (rows // share) * T and pos.shape == (rows, T) and pos.dtype == torch.uint8
    ctx = torch.empty((rows * T, H), dtype=torch.bfloat16, device=qkv.device)
    nat.NEXT_META = 4.0 * rows * T * T * H
    nat.call("agb_masked_attention_bf16_scatter", nat.ptr(qkv), nat.ptr(packed_mask), packed_mask.shape[1], rows, share, T, H,
             heads, nat.ptr(pos), nat.ptr(ctx), nat.stream())
    return ctx


def cls_attention_varlen(q: Tensor, kv: Tensor, k_off: int, v_off: int, cu: Tensor, max_len: int, heads: int) -> Tensor:
    """CLS-query attention over packed rows: q (rows, H), kv (total_tokens, ld) -> ctx (rows, H) bf16."""
    assert q.dtype == torch.bfloat16 and kv.dtype == torch.bfloat16 and q.stride(1) == 1 and kv.stride(1) == 1
    rows, H = q.shape
    ctx = torch.empty((rows, H), dtype=torch.bfloat16, device=q.device)
    step = 65535
    for r0 in range(0, rows, step):
        r1 = min(rows, r0 + step)
        nat.call("agb_cls_attention_varlen", nat.ptr(q[r0:r1]), q.stride(0), nat.ptr(kv), kv.stride(0), k_off, v_off, 1,
                 nat.ptr(cu[r0:r1 + 1]), r1 - r0, max_len, H, heads, nat.ptr(ctx[r0:r1]), H, nat.stream())
    return ctx


# ------------------------------------------------------------------------------------------------
# explainer head / loss
# ------------------------------------------------------------------------------------------------
def explainer_head_fwd(h: Tensor, B: int, T: int, W: Tensor, bias: Tensor, grand: Optional[Tensor],
                       null: Optional[Tensor], normalize: bool, want_pred: bool = False):
    assert h.dim() == 2 and h.is_contiguous() and h.shape[0] == B * T
    E, C = h.shape[1], W.shape[0]
    phi = torch.empty((B, C, T - 1), dtype=torch.float32, device=h.device)
    pred = torch.empty((B, T, C), dtype=torch.float32, device=h.device) if want_pred else None
    if normalize:
        grand = _c(grand.float())
        null = _c(null.float().reshape(-1))
        assert grand.shape == (B, C) and null.numel() == C
    nat.call("agb_explainer_head_fwd", nat.ptr(h), 1 if h.dtype == torch.bfloat16 else 0, B, T, E, C, nat.ptr(_c(W)),
             nat.ptr(_c(bias)), nat.ptr(grand) if normalize else None, nat.ptr(null) if normalize else None,
             1 if normalize else 0, nat.ptr(phi), nat.ptr(pred), nat.stream())
    return (phi, pred) if want_pred else phi


def explainer_head_bwd(dphi: Tensor, h: Tensor, B: int, T: int, W: Tensor, normalize: bool,
                       dW: Optional[Tensor], db: Optional[Tensor], want_dh: bool = True) -> Optional[Tensor]:
    E, C = h.shape[1], W.shape[0]
    dphi = _c(dphi.float())
    dh = torch.empty_like(h) if want_dh else None
    nat.call("agb_explainer_head_bwd", nat.ptr(dphi), nat.ptr(h), 1 if h.dtype == torch.bfloat16 else 0, B, T, E, C,
             nat.ptr(_c(W)), 1 if normalize else 0, nat.ptr(dh), nat.ptr(dW), nat.ptr(db), nat.stream())
    return dh


def normalize_shapley(pred: Tensor, grand: Tensor, null: Tensor) -> Tensor:
    pred = _c(pred.float())
    B, T, C = pred.shape
    out = torch.empty_like(pred)
    nat.call("agb_normalize_shapley", nat.ptr(pred), nat.ptr(_c(grand.float())), nat.ptr(_c(null.float().reshape(-1))),
             B, T, C, nat.ptr(out), nat.stream())
    return out


def shapley_loss_fwd(packed: Tensor, v0: Tensor, v_s: Tensor, phi: Tensor, B: int, S: int, n: int):
    C = phi.shape[1]
    packed, v_s, phi = _c(packed), _c(v_s.float()), _c(phi.float())
    v0 = _c(v0.float().reshape(-1))
    assert packed.shape[0] == B * S and v_s.shape == (B * S, C) and phi.shape == (B, C, n)
    resid = torch.empty((B * S, C), dtype=torch.float32, device=phi.device)
    partial = torch.empty((B,), dtype=torch.float32, device=phi.device)
    loss = torch.empty((1,), dtype=torch.float32, device=phi.device)
    nat.call("agb_shapley_loss_fwd", nat.ptr(packed), packed.shape[1], nat.ptr(v0), nat.ptr(v_s), nat.ptr(phi), B, S, n,
             C, nat.ptr(resid), nat.ptr(partial), nat.ptr(loss), nat.stream())
    return loss.reshape(()), resid


def shapley_loss_bwd(packed: Tensor, resid: Tensor, grad_out: Optional[Tensor], B: int, S: int, n: int, C: int) -> Tensor:
    dphi = torch.empty((B, C, n), dtype=torch.float32, device=resid.device)
    g = _c(grad_out.float().reshape(1)) if grad_out is not None else None
    nat.call("agb_shapley_loss_bwd", nat.ptr(_c(packed)), packed.shape[1], nat.ptr(resid), nat.ptr(g), B, S, n, C,
             nat.ptr(dphi), nat.stream())
    return dphi


# ------------------------------------------------------------------------------------------------
# training adjoints
# ------------------------------------------------------------------------------------------------
def _is_bf16(t: Tensor) -> int:
    assert t.dtype in (torch.float32, torch.bfloat16)
    return 1 if t.dtype == torch.bfloat16 else 0


def gelu_fwd(z: Tensor) -> Tensor:
    z = _c(z)
    out = torch.empty_like(z)
    nat.call("agb_gelu_fwd", nat.ptr(z), nat.ptr(out), z.numel(), _is_bf16(z), nat.stream())
    return out


def gelu_bwd(dy: Tensor, z: Tensor) -> Tensor:
    dy, z = _c(dy), _c(z)
    assert dy.dtype == z.dtype and dy.shape == z.shape
    dz = torch.empty_like(z)
    nat.call("agb_gelu_bwd", nat.ptr(dy), nat.ptr(z), nat.ptr(dz), z.numel(), _is_bf16(z), nat.stream())
    return dz


def colsum_into(y: Tensor, out: Tensor) -> None:
    """out[n] += sum_m y[m, n]   (bias gradients)"""
    assert y.dim() == 2 and y.stride(1) == 1 and out.dtype == torch.float32 and out.numel() == y.shape[1]
    nat.call("agb_colsum", nat.ptr(y), _is_bf16(y), y.stride(0), y.shape[0], y.shape[1], nat.ptr(out), nat.stream())


def layernorm_bwd(x: Tensor, dy: Tensor, gamma: Tensor, eps: float, dres: Optional[Tensor], dgamma: Optional[Tensor],
                  dbeta: Optional[Tensor]) -> Tensor:
    x, dy = _c(x), _c(dy)
    rows, H = x.shape
    dx = torch.empty((rows, H), dtype=torch.float32, device=x.device)
    if dres is not None:
        assert dres.dtype == torch.float32 and dres.is_contiguous()
    nat.call("agb_layernorm_bwd", nat.ptr(x), _is_bf16(x), nat.ptr(dy), _is_bf16(dy), nat.ptr(gamma), nat.ptr(dres), rows, H,
             float(eps), nat.ptr(dx), nat.ptr(dgamma), nat.ptr(dbeta), nat.stream())
    return dx


def masked_attention_bwd(qkv: Tensor, dctx: Tensor, packed_mask: Tensor, T: int, heads: int, mode: int) -> Tensor:
    qkv, dctx = _c(qkv), _c(dctx)
    assert qkv.dtype == dctx.dtype
    rows = qkv.shape[0] // T
    H = qkv.shape[1] // 3
    dqkv = torch.empty_like(qkv)
    nat.call("agb_masked_attention_bwd", nat.ptr(qkv), nat.ptr(dctx), _is_bf16(qkv), nat.ptr(packed_mask),
             packed_mask.shape[1], rows, T, H, heads, mode, nat.ptr(dqkv), nat.stream())
    return dqkv


# ---- dropout (training mode; masks are regenerated from a counter hash, never stored) --------------------------------
def dropout_thr(p: float) -> int:
    """16-bit threshold of a drop probability: an element is dropped iff its 16 hash bits < round(p * 65536)."""
    assert 0.0 <= p < 1.0
    return int(round(p * 65536.0))


def dropout(y: Tensor, thr16: int, seed: int, tag: int, *, residual: Optional[Tensor] = None, out_dtype=None) -> Tensor:
    """out = residual + keep * y / (1 - p)   (agb_dropout); the adjoint is the same call on the gradient."""
    y = _c(y)
    out = torch.empty(y.shape, dtype=out_dtype or y.dtype, device=y.device)
    if residual is not None:
        residual = _c(residual)
        assert residual.dtype == torch.float32 and residual.shape == y.shape
    nat.call("agb_dropout", nat.ptr(y), _is_bf16(y), nat.ptr(residual), nat.ptr(out), _is_bf16(out), y.numel(), thr16,
             seed & 0xFFFFFFFFFFFFFFFF, tag, nat.stream())
    return out


def gemm_bf16_gelu_dual(a: Tensor, w: Tensor, bias: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    """-> (z = a @ w.T + bias, GELU(z)), both bf16, from one GEMM (agb_gemm_bf16_gelu_dual; N >= 192)."""
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and N >= 192 and N % 8 == 0
    z = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    f = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    nat.NEXT_META = 2.0 * M * N * K
    nat.call("agb_gemm_bf16_gelu_dual", nat.ptr(a), a.stride(0), nat.ptr(w), w.stride(0), M, N, K, nat.ptr(bias), nat.ptr(z),
             nat.ptr(f), N, nat.stream())
    return z, f


def gemm_bf16_gelu_bwd(dy: Tensor, w: Tensor, z: Tensor) -> Tensor:
    """dz = (dy @ w) * GELU'(z): dy (M, K), w (K, N) the forward weight of the layer AFTER the GELU, z (M, N) the stored
    pre-activation (agb_gemm_bf16_gelu_bwd; N >= 192)."""
    assert dy.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and z.dtype == torch.bfloat16
    assert dy.stride(1) == 1 and w.stride(1) == 1 and z.stride(1) == 1
    M, K = dy.shape
    N = w.shape[1]
    assert w.shape[0] == K and z.shape == (M, N) and N >= 192 and N % 8 == 0
    dz = torch.empty((M, N), dtype=torch.bfloat16, device=dy.device)
    nat.NEXT_META = 2.0 * M * N * K
    nat.call("agb_gemm_bf16_gelu_bwd", nat.ptr(dy), dy.stride(0), nat.ptr(w), w.stride(0), M, N, K, nat.ptr(z), z.stride(0),
             nat.ptr(dz), N, nat.stream())
    return dz


def gemm_dropout_residual_supported(M: int, N: int) -> bool:
    return N >= 192 and N % 4 == 0 and M * N < (1 << 32)


def gemm_bf16_dropout_residual(a: Tensor, w: Tensor, bias: Optional[Tensor], residual: Tensor, thr16: int, seed: int, tag: int) -> Tensor:
    """residual + dropout(a @ w.T + bias) in one tcgen05 GEMM (agb_gemm_bf16_dropout_residual); the keep mask is the stream of
    dropout(seed, tag) over the (M, N) output, so `dropout` on the gradient is its adjoint."""
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and residual.shape == (M, N) and residual.dtype == torch.float32 and residual.stride(1) == 1
    assert gemm_dropout_residual_supported(M, N) and 0 < thr16 < 65536
    out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    nat.NEXT_META = 2.0 * M * N * K
    nat.call("agb_gemm_bf16_dropout_residual", nat.ptr(a), a.stride(0), nat.ptr(w), w.stride(0), M, N, K, nat.ptr(bias),
             nat.ptr(residual), residual.stride(0), nat.ptr(out), thr16, seed & 0xFFFFFFFFFFFFFFFF, tag, nat.stream())
    return out


def masked_attention_dropout(qkv: Tensor, packed_mask: Tensor, T: int, heads: int, mode: int, thr16: int, seed: int) -> Tensor:
    """masked_attention with dropout on the probabilities (bf16: head dim 64 with T <= 256 on the tcgen05 kernels, or head
    dims 8/16/32; fp32: CUDA-core kernels)."""
    qkv = _c(qkv)
    rows = packed_mask.shape[0]
    H = qkv.shape[1] // 3
    ctx = torch.empty((rows * T, H), dtype=qkv.dtype, device=qkv.device)
    nat.call("agb_masked_attention_dropout_fwd", nat.ptr(qkv), _is_bf16(qkv), nat.ptr(packed_mask), packed_mask.shape[1], rows,
             T, H, heads, mode, nat.ptr(ctx), thr16, seed & 0xFFFFFFFFFFFFFFFF, nat.stream())
    return ctx


def masked_attention_dropout_bwd(qkv: Tensor, dctx: Tensor, packed_mask: Tensor, T: int, heads: int, mode: int, thr16: int,
                                 seed: int) -> Tensor:
    qkv, dctx = _c(qkv), _c(dctx)
    assert qkv.dtype == dctx.dtype
    rows = qkv.shape[0] // T
    H = qkv.shape[1] // 3
    dqkv = torch.empty_like(qkv)
    nat.call("agb_masked_attention_dropout_bwd", nat.ptr(qkv), nat.ptr(dctx), _is_bf16(qkv), nat.ptr(packed_mask),
             packed_mask.shape[1], rows, T, H, heads, mode, nat.ptr(dqkv), thr16, seed & 0xFFFFFFFFFFFFFFFF, nat.stream())
    return dqkv


def attention_dropout_mask(rows: int, heads: int, T: int, thr16: int, seed: int, device) -> Tensor:
    """(rows, heads, T, T) uint8 keep mask the attention kernels regenerate (tests / diagnostics)."""
    keep = torch.empty((rows, heads, T, T), dtype=torch.uint8, device=device)
    nat.call("agb_attention_dropout_mask", nat.ptr(keep), rows, heads, T, thr16, seed & 0xFFFFFFFFFFFFFFFF, nat.stream())
    return keep


def vit_embed_bwd(dx: Tensor, B: int, T: int, H: int, dpos: Tensor, dcls: Tensor, patch_dtype) -> Tensor:
    dpatch = torch.empty((B * (T - 1), H), dtype=patch_dtype, device=dx.device)
    nat.call("agb_vit_embed_bwd", nat.ptr(_c(dx)), B, T, H, nat.ptr(dpos), nat.ptr(dcls), nat.ptr(dpatch),
             1 if patch_dtype == torch.bfloat16 else 0, nat.stream())
    return dpatch


def bert_embed_sum(ids: Tensor, word: Tensor, pos: Tensor, type0: Tensor) -> Tensor:
    ids = _c(ids)
    B, T = ids.shape
    H = word.shape[1]
    out = torch.empty((B * T, H), dtype=torch.float32, device=ids.device)
    nat.call("agb_bert_embed_sum", nat.ptr(ids), nat.ptr(word), nat.ptr(pos), nat.ptr(type0), B * T, T, H, word.shape[0],
             nat.ptr(out), nat.stream())
    return out


def bert_embed_scatter(ids: Tensor, dsum: Tensor, pad_id: int, dword: Tensor, dpos: Tensor, dtype0: Tensor) -> None:
    ids = _c(ids)
    B, T = ids.shape
    H = dword.shape[1]
    nat.call("agb_bert_embed_scatter", nat.ptr(ids), nat.ptr(_c(dsum)), B * T, T, H, dword.shape[0], pad_id, nat.ptr(dword),
             nat.ptr(dpos), nat.ptr(dtype0), nat.stream())


# ------------------------------------------------------------------------------------------------
# KernelSHAP
# ------------------------------------------------------------------------------------------------
def pack_feature_masks(Z: Tensor) -> Tensor:
    """(rows, d) {0,1} -> packed (rows, ceil(d/32)) int32 with bit j = feature j (no CLS offset)."""
    return pack_masks(Z.to(torch.int64), prepend_cls=False)


def kernelshap_solve(Zp: Tensor, weights: Tensor, probs: Tensor, f_x: Tensor, f_null: Tensor, d: int,
                     link_logit: bool = True):
    """Zp (B,S,words) packed, weights (B,S), probs (B,S,C), f_x (B,C), f_null (C) -> (phi (B,C,d) float64, info (B))"""
    B, S, words = Zp.shape
    C = probs.shape[2]
    dev = Zp.device
    f64 = lambda t: _c(t.to(torch.float64))  # noqa: E731
    weights, probs, f_x, f_null = f64(weights), f64(probs), f64(f_x), f64(f_null.reshape(-1))
    n = d - 1
    A = torch.empty((B, n * n), dtype=torch.float64, device=dev)
    R = torch.empty((B, n, C), dtype=torch.float64, device=dev)
    phi = torch.empty((B, C, d), dtype=torch.float64, device=dev)
    info = torch.empty((B,), dtype=torch.int32, device=dev)
    nat.call("agb_kernelshap_solve", nat.ptr(_c(Zp)), words, nat.ptr(weights), nat.ptr(probs), nat.ptr(f_x), nat.ptr(f_null),
             B, S, d, C, 1 if link_logit else 0, nat.ptr(A), nat.ptr(R), nat.ptr(phi), nat.ptr(info), nat.stream())
    return phi, info
