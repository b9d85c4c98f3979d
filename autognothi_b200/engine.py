"""Forward engine: strings the CUDA kernels into the masked surrogate / explainer passes.

Two precision policies share one code path:
  * "bf16"  — the throughput path: bf16 operands on tcgen05 tensor cores, fp32 accumulation, an fp32
              residual stream, fp32 LayerNorm statistics, fp32 heads / normalisation / loss;
  * "fp32"  — the exact path (CUDA-core fp32 GEMM + attention) used to verify against the reference
              at rtol 1e-4 (BASELINE.json north_star).
Inputs are (B, ...) tensors plus packed coalition masks of shape (B*S, words): each input is embedded
ONCE and broadcast to its S coalition rows inside the embedding kernel, so the reference's Xs_EXT
replication (scripts/train_explainer.py:159-163) never exists.  S = 1 gives the reference-shaped call.
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional, Tuple

import torch
from torch import Tensor

from . import ops

State = Dict[str, Tensor]

# bf16 ViT backbone: fold every LayerNorm into the neighbouring tcgen05 GEMMs (no LayerNorm kernels, the fp32
# residual stream is read once per GEMM instead of once more per LayerNorm).  Off = the LayerNorm-kernel path.
FUSE_LAYERNORM = True


def _is_vit(cfg) -> bool:
    return hasattr(cfg, "img_px_size")


def n_players_of(cfg) -> int:
    if _is_vit(cfg):
        return (cfg.img_px_size // cfg.img_patch_size) ** 2
    return cfg.max_position_embeddings - 1


class _Policy:
    def __init__(self, precision: str):
        assert precision in ("bf16", "fp32"), precision
        self.precision = precision
        self.bf16 = precision == "bf16"
        self.act_dtype = torch.bfloat16 if self.bf16 else torch.float32

    def weight(self, w: Tensor) -> Tensor:
        w = w.detach().reshape(w.shape[0], -1).contiguous().float()
        return ops.to_bf16(w) if self.bf16 else w

    def linear(self, a: Tensor, w: Tensor, b: Optional[Tensor], *, act: int = 0, residual: Optional[Tensor] = None,
               out_f32: bool = False, out: Optional[Tensor] = None, res_group: int = 0, res_rows: int = 0) -> Tensor:
        if self.bf16:
            return ops.gemm_bf16(a, w, b, act=act, residual=residual, out=out,
                                 out_dtype=torch.float32 if out_f32 else torch.bfloat16,
                                 res_group=res_group, res_rows=res_rows)
        assert res_group == 0
        return ops.gemm_f32(a, w, b, act=act, residual=residual, out=out)

    def ln(self, x: Tensor, g: Tensor, b: Tensor, eps: float, *, want_f32: bool = False) -> Tuple[Tensor, Optional[Tensor]]:
        """-> (activation-dtype copy, fp32 copy | None)"""
        if self.bf16:
            ob, of = ops.layernorm(x, g, b, eps, want_bf16=True, want_f32=want_f32)
            return ob, of
        _, of = ops.layernorm(x, g, b, eps, want_bf16=False, want_f32=True)
        return of, of

    def act(self, x_f32: Tensor) -> Tensor:
        return ops.to_bf16(x_f32) if self.bf16 else x_f32


def _f32(t: Tensor) -> Tensor:
    return t.detach().float().contiguous()


class LayerWeights:
    """One transformer block in kernel-ready form (fused QKV, operand dtype per policy)."""

    def __init__(self, sd: State, prefix: str, pol: _Policy, vit: bool):
        sa = prefix + ".attention.self."
        self.wqkv = pol.weight(torch.cat([sd[sa + "query.weight"], sd[sa + "key.weight"], sd[sa + "value.weight"]], 0))
        self.bqkv = _f32(torch.cat([sd[sa + "query.bias"], sd[sa + "key.bias"], sd[sa + "value.bias"]], 0))
        self.wo = pol.weight(sd[prefix + ".attention.output.dense.weight"])
        self.bo = _f32(sd[prefix + ".attention.output.dense.bias"])
        self.w1 = pol.weight(sd[prefix + ".intermediate.dense.weight"])
        self.b1 = _f32(sd[prefix + ".intermediate.dense.bias"])
        self.w2 = pol.weight(sd[prefix + ".output.dense.weight"])
        self.b2 = _f32(sd[prefix + ".output.dense.bias"])
        n1 = prefix + (".layernorm_before" if vit else ".attention.output.LayerNorm")
        n2 = prefix + (".layernorm_after" if vit else ".output.LayerNorm")
        self.ln1 = (_f32(sd[n1 + ".weight"]), _f32(sd[n1 + ".bias"])) if (n1 + ".weight") in sd else None
        self.ln2 = (_f32(sd[n2 + ".weight"]), _f32(sd[n2 + ".bias"])) if (n2 + ".weight") in sd else None
        # pre-LN blocks in bf16 mode: LayerNorm folded into the consuming GEMM (ops.gemm_bf16_fused); built lazily by
        # fold() because only the inference engines use it (the training tape keeps explicit LayerNorm outputs)
        self.folded = None
        self._fold_src = (sd, sa, prefix) if (vit and pol.bf16 and self.ln1 is not None and self.ln2 is not None) else None

    def fold(self) -> Optional[dict]:
        """W' = W * gamma (bf16), b' = b + W beta, colsum_j = sum_k W'_jk of the ROUNDED weights."""
        if self.folded is None and self._fold_src is not None:
            sd, sa, prefix = self._fold_src
            wqkv32 = torch.cat([sd[sa + "query.weight"], sd[sa + "key.weight"], sd[sa + "value.weight"]], 0).detach().float()
            w132 = sd[prefix + ".intermediate.dense.weight"].detach().float()
            g1, b1 = self.ln1
            g2, b2 = self.ln2
            wq = ops.to_bf16((wqkv32 * g1[None, :]).contiguous())
            w1 = ops.to_bf16((w132 * g2[None, :]).contiguous())
            self.folded = {
                "wqkv": wq, "bqkv": (self.bqkv + wqkv32 @ b1).contiguous(), "cqkv": wq.float().sum(1).contiguous(),
                "w1": w1, "b1": (self.b1 + w132 @ b2).contiguous(), "c1": w1.float().sum(1).contiguous(),
            }
            self._fold_src = None
        return self.folded


class BackboneWeights:
    def __init__(self, sd: State, cfg, pol: _Policy):
        self.vit = _is_vit(cfg)
        root = "vit" if self.vit else "bert"
        self.layers = [LayerWeights(sd, f"{root}.encoder.layers.{i}", pol, self.vit) for i in range(cfg.num_hidden_layers)]
        if self.vit:
            e = "vit.embeddings."
            self.cls_token = _f32(sd[e + "cls_token"]).reshape(-1)
            self.pos_emb = _f32(sd[e + "position_embeddings"]).reshape(-1, cfg.hidden_size)
            self.w_patch = pol.weight(sd[e + "patch_embeddings.projection.weight"])
            self.b_patch = _f32(sd[e + "patch_embeddings.projection.bias"])
            self.final_ln = (_f32(sd["vit.layernorm.weight"]), _f32(sd["vit.layernorm.bias"]))
        else:
            e = "bert.embeddings."
            self.word = _f32(sd[e + "word_embeddings.weight"])
            self.pos = _f32(sd[e + "position_embeddings.weight"])
            self.type0 = _f32(sd[e + "token_type_embeddings.weight"])[0].contiguous()
            self.emb_ln = (_f32(sd[e + "LayerNorm.weight"]), _f32(sd[e + "LayerNorm.bias"]))


def embed(bw: BackboneWeights, cfg, pol: _Policy, xs: Tensor, S: int) -> Tensor:
    """-> x (B*S, T, H) fp32 (reference models/vanilla_vit.py:242-253, models/vanilla_bert.py:307-325)"""
    H = cfg.hidden_size
    if bw.vit:
        assert xs.dim() == 4 and xs.shape[1] == cfg.img_channels and xs.shape[2] == cfg.img_px_size == xs.shape[3], \
            "pixel_values must be (B, img_channels, img_px_size, img_px_size)"
        B = xs.shape[0]
        T = n_players_of(cfg) + 1
        patches = ops.vit_im2col(xs.float(), cfg.img_patch_size, pol.act_dtype)
        pe = pol.linear(patches, bw.w_patch, bw.b_patch, out_f32=True)
        return ops.vit_assemble(pe, bw.cls_token, bw.pos_emb, B, S, T, H)
    assert xs.dim() == 2 and xs.dtype == torch.int64, "input_ids must be (B, T) int64"
    assert xs.shape[1] <= cfg.max_position_embeddings
    return ops.bert_embed(xs, bw.word, bw.pos, bw.type0, bw.emb_ln[0], bw.emb_ln[1], cfg.layer_norm_eps, S)


def vit_layer(pol: _Policy, lw: LayerWeights, x: Tensor, masks: Tensor, T: int, heads: int, eps: float,
              ctx: Optional[Tensor] = None) -> Tensor:
    """pre-LN block on the fp32 residual stream x (rows*T, H); updates x in place.
    reference models/vanilla_vit.py:364-377.  ctx: attention output computed by the caller (first-block sharing)."""
    if ctx is None:
        h = pol.ln(x, lw.ln1[0], lw.ln1[1], eps)[0] if lw.ln1 is not None else pol.act(x)
        qkv = pol.linear(h, lw.wqkv, lw.bqkv)
        ctx = ops.masked_attention(qkv, masks, T, heads, ops.MASK_MUL0)
    pol.linear(ctx, lw.wo, lw.bo, residual=x, out_f32=True, out=x)
    h = pol.ln(x, lw.ln2[0], lw.ln2[1], eps)[0]
    f = pol.linear(h, lw.w1, lw.b1, act=ops.ACT_GELU)
    pol.linear(f, lw.w2, lw.b2, residual=x, out_f32=True, out=x)
    return x


def vit_layer_fused(lw: LayerWeights, x: Tensor, x16: Optional[Tensor], stats: Optional[Tensor], masks: Tensor, T: int,
                    heads: int, eps: float, last: bool, ctx: Optional[Tensor] = None, nkeep: Optional[Tensor] = None
                    ) -> Tuple[Tensor, Optional[Tensor], Optional[Tensor]]:
    """Same block as vit_layer (bf16 mode) without LayerNorm kernels: x16 / stats are the bf16 copy and per-row
    (sum, sum of squares) partials of the fp32 residual stream x, produced by the previous residual GEMM's epilogue
    (or ops.rowstats_cast at the entry).  Returns (x, x16, stats) for the next block."""
    f = lw.fold()
    if ctx is None:
        qkv, _, _ = ops.gemm_bf16_fused(x16, f["wqkv"], f["bqkv"], ln=(stats, f["cqkv"], eps))
        # nkeep: the rows are in kept-first token order -> the masked keys fold into one virtual key
        ctx = ops.attention_prefix(qkv, nkeep, T, heads) if nkeep is not None else ops.masked_attention(qkv, masks, T, heads, ops.MASK_MUL0)
    _, y16, ystats = ops.gemm_bf16_fused(ctx, lw.wo, lw.bo, residual=x, out=x, emit_copy_stats=True)
    h = ops.gemm_bf16_fused(y16, f["w1"], f["b1"], act=ops.ACT_GELU, ln=(ystats, f["c1"], eps))[0]
    if last:
        ops.gemm_bf16(h, lw.w2, lw.b2, residual=x, out=x, out_dtype=torch.float32)
        return x, None, None
    _, x16n, statsn = ops.gemm_bf16_fused(h, lw.w2, lw.b2, residual=x, out=x, emit_copy_stats=True)
    return x, x16n, statsn


# Surrogate / classifier evaluation of the bf16 ViT backbone (fused chain, CLS-only last block): the residual stream lives in
# HBM as two bf16 planes, x = hi + lo (16 significant bits; hi = bf16(x)), updated in place by the residual GEMMs
# (agb_gemm_bf16_hilo).  The hi plane IS the bf16 copy the LayerNorm-folded QKV / FC1 GEMMs read, so the separate copy of the
# fp32-stream variant is never written: 10 instead of 12 bytes per element for the HBM-bound output projection.  A plain bf16
# stream (8 bits) pushes the attributions past the 1e-2 tolerance; 16 bits does not move them.
HILO_RESIDUAL = os.environ.get("AGB_HILO_RESIDUAL", "1") != "0"      # the environment switch is for A/B runs of bench.py


def vit_layer_hilo(lw: LayerWeights, xh: Tensor, xl: Tensor, stats: Optional[Tensor], T: int, heads: int, eps: float,
                   masks: Tensor, ctx: Optional[Tensor] = None, nkeep: Optional[Tensor] = None) -> Tensor:
    """vit_layer_fused on the hi/lo residual stream (xh, xl) (rows*T, H) bf16, updated in place; stats: row statistics of
    the incoming stream (unused when the caller supplies the first block's ctx).  -> statistics of the outgoing stream."""
    f = lw.fold()
    if ctx is None:
        qkv, _, _ = ops.gemm_bf16_fused(xh, f["wqkv"], f["bqkv"], ln=(stats, f["cqkv"], eps))
        ctx = ops.attention_prefix(qkv, nkeep, T, heads) if nkeep is not None else ops.masked_attention(qkv, masks, T, heads, ops.MASK_MUL0)
    ystats = ops.gemm_bf16_hilo(ctx, lw.wo, lw.bo, xh, xl)
    h = ops.gemm_bf16_fused(xh, f["w1"], f["b1"], act=ops.ACT_GELU, ln=(ystats, f["c1"], eps))[0]
    return ops.gemm_bf16_hilo(h, lw.w2, lw.b2, xh, xl)


def bert_layer(pol: _Policy, lw: LayerWeights, x: Tensor, xa: Optional[Tensor], masks: Tensor, T: int, heads: int, eps: float,
               ctx: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """post-LN block; x fp32 residual stream, xa its activation-dtype copy.
    reference models/vanilla_bert.py:396-427, 556-560, 600-604.  ctx: attention output computed by the caller."""
    if ctx is None:
        qkv = pol.linear(xa, lw.wqkv, lw.bqkv)
        ctx = ops.masked_attention(qkv, masks, T, heads, ops.MASK_NEGINF)
    a = pol.linear(ctx, lw.wo, lw.bo, residual=x, out_f32=True)
    if lw.ln1 is not None:
        aa, a = pol.ln(a, lw.ln1[0], lw.ln1[1], eps, want_f32=True)
    else:
        aa = pol.act(a)
    f = pol.linear(aa, lw.w1, lw.b1, act=ops.ACT_GELU)
    y = pol.linear(f, lw.w2, lw.b2, residual=a, out_f32=True)
    ya, y = pol.ln(y, lw.ln2[0], lw.ln2[1], eps, want_f32=True)
    return y, ya


# Exact work-skipping for surrogate / classifier evaluation: the heads read only token 0 of the last hidden state
# (reference models/vanilla_vit.py:51-56, models/vanilla_bert.py:615-619), so the LAST block needs keys / values for
# every token but the query, the attention output, the output projection and the whole MLP only for the CLS row.
# (The explainer reads all tokens and never takes this path.)  Off = every block runs on all T tokens.
CLS_ONLY_LAST_BLOCK = True
# Exact work-skipping, first block: LayerNorm + QKV projection once per input instead of once per coalition (bf16 path).
SHARE_FIRST_BLOCK = True
# Exact work-skipping for additive (-inf) masks (BERT surrogate / classifier heads, bf16 path): masked tokens are never
# attended to and the head reads only token 0, so only the kept tokens of every row are carried through the encoder
# (packed back to back; variable-length attention).  SURVEY.md 8f-3.
DROP_MASKED_TOKENS = True
# ViT surrogate / classifier evaluation (multiplicative "logit := 0" masks, bf16, head reads token 0): the tokens of every
# (input, coalition) row are permuted so that the kept ones come first.  Everything between the attentions is token-wise
# and attention is permutation-equivariant, so the CLS output is unchanged; the attention kernel then sees the masked keys
# as one contiguous tail, all with the logit 0, and folds them into ONE virtual key (agb_attention_bf16_prefix).
# Exact (tests/test_gpu_edges.py).  Round 1 built the order with torch.sort + two index_select gathers, which gave back what
# the attention kernel gained; now the order comes from one small kernel (agb_kept_first_order), the residual stream is
# gathered in that order instead of being replicated (agb_gather_token_rows) and the first block's attention scatters its
# output rows (agb_masked_attention_bf16_scatter), so the switch costs nothing extra per step.
KEPT_FIRST_ORDER = True


def last_block_cls_only(pol: _Policy, lw: LayerWeights, vit: bool, x: Optional[Tensor], xa: Optional[Tensor], x16: Optional[Tensor],
                        stats: Optional[Tensor], masks: Tensor, T: int, heads: int, eps: float,
                        x_lo: Optional[Tensor] = None) -> Tuple[Tensor, Optional[Tensor]]:
    """Last encoder block restricted to the CLS query.  x (rows*T, H) fp32 residual stream; ViT fused path passes
    (x16, stats) for the folded LayerNorm, otherwise they are None; BERT passes xa (activation-dtype copy of x).
    hi/lo residual stream: x is None, x16 is the hi plane and x_lo the lo plane.
    -> (x_cls (rows, H) fp32 after the block, activation copy | None)"""
    rows = masks.shape[0]
    if x is None:
        H = x16.shape[1]
        x_cls = x16.view(rows, T, H)[:, 0, :].float() + x_lo.view(rows, T, H)[:, 0, :].float()
    else:
        H = x.shape[1]
        x_cls = x.view(rows, T, H)[:, 0, :].contiguous()
    wq, wkv, bq, bkv = lw.wqkv[:H], lw.wqkv[H:], lw.bqkv[:H], lw.bqkv[H:]
    if vit:
        if x16 is not None:
            f = lw.fold()
            kv = ops.gemm_bf16_fused(x16, f["wqkv"][H:], f["bqkv"][H:], ln=(stats, f["cqkv"][H:], eps))[0]
            x16_cls = x16.view(rows, T, H)[:, 0, :].contiguous()
            st_cls = stats.view(rows, T, stats.shape[1], 2)[:, 0].contiguous()
            q = ops.gemm_bf16_fused(x16_cls, f["wqkv"][:H], f["bqkv"][:H], ln=(st_cls, f["cqkv"][:H], eps))[0]
        else:
            h = pol.ln(x, lw.ln1[0], lw.ln1[1], eps)[0] if lw.ln1 is not None else pol.act(x)
            kv = pol.linear(h, wkv, bkv)
            q = pol.linear(h.view(rows, T, H)[:, 0, :].contiguous(), wq, bq)
        ctx = ops.cls_attention(q, kv, 0, H, masks, T, heads, ops.MASK_MUL0)
        y = pol.linear(ctx, lw.wo, lw.bo, residual=x_cls, out_f32=True)
        h2 = pol.ln(y, lw.ln2[0], lw.ln2[1], eps)[0]
        ff = pol.linear(h2, lw.w1, lw.b1, act=ops.ACT_GELU)
        return pol.linear(ff, lw.w2, lw.b2, residual=y, out_f32=True), None
    kv = pol.linear(xa, wkv, bkv)
    q = pol.linear(xa.view(rows, T, H)[:, 0, :].contiguous(), wq, bq)
    ctx = ops.cls_attention(q, kv, 0, H, masks, T, heads, ops.MASK_NEGINF)
    a = pol.linear(ctx, lw.wo, lw.bo, residual=x_cls, out_f32=True)
    if lw.ln1 is not None:
        aa, a = pol.ln(a, lw.ln1[0], lw.ln1[1], eps, want_f32=True)
    else:
        aa = pol.act(a)
    ff = pol.linear(aa, lw.w1, lw.b1, act=ops.ACT_GELU)
    y = pol.linear(ff, lw.w2, lw.b2, residual=a, out_f32=True)
    ya, y = pol.ln(y, lw.ln2[0], lw.ln2[1], eps, want_f32=True)
    return y, ya


def run_bert_packed(bw: BackboneWeights, cfg, pol: _Policy, xs: Tensor, masks: Tensor, S: int) -> Tuple[Tensor, Tensor]:
    """BERT surrogate / classifier evaluation on the kept tokens only -> (x_cls (rows, H) fp32, activation copy).
    Every block but the last runs on the packed tokens; the last one on the CLS query (see last_block_cls_only)."""
    T = n_players_of(cfg) + 1
    H, heads, eps = cfg.hidden_size, cfg.num_attention_heads, cfg.layer_norm_eps
    rows = masks.shape[0]
    cu, src, total = ops.pack_kept_tokens(masks, T, S)
    x_img = embed(bw, cfg, pol, xs, 1).reshape(-1, H)                       # (B*T, H): embeddings once per input
    x = x_img.index_select(0, src)                                           # (total, H) packed residual stream
    xa = pol.act(x)
    for lw in bw.layers[:-1]:
        qkv = pol.linear(xa, lw.wqkv, lw.bqkv)
        ctx = ops.attention_varlen(qkv, cu, T, heads)
        x, xa = bert_layer(pol, lw, x, xa, masks, T, heads, eps, ctx=ctx)
    lw = bw.layers[-1]
    first = cu[:-1].long()                                                   # packed index of every row's CLS token
    x_cls, xa_cls = x.index_select(0, first), xa.index_select(0, first)
    kv = pol.linear(xa, lw.wqkv[H:], lw.bqkv[H:])
    q = pol.linear(xa_cls, lw.wqkv[:H], lw.bqkv[:H])
    ctx = ops.cls_attention_varlen(q, kv, 0, H, cu, T, heads)
    a = pol.linear(ctx, lw.wo, lw.bo, residual=x_cls, out_f32=True)
    if lw.ln1 is not None:
        aa, a = pol.ln(a, lw.ln1[0], lw.ln1[1], eps, want_f32=True)
    else:
        aa = pol.act(a)
    ff = pol.linear(aa, lw.w1, lw.b1, act=ops.ACT_GELU)
    y = pol.linear(ff, lw.w2, lw.b2, residual=a, out_f32=True)
    ya, y = pol.ln(y, lw.ln2[0], lw.ln2[1], eps, want_f32=True)
    assert y.shape[0] == rows
    return y, ya


def run_backbone(bw: BackboneWeights, cfg, pol: _Policy, xs: Tensor, masks: Tensor, S: int, cls_only: bool = False,
                 layer_hook=None, token0_only: bool = False) -> Tuple[Tensor, Optional[Tensor]]:
    """-> (x (rows*T, H) fp32 after the encoder stack [ViT: BEFORE the final LayerNorm], activation copy | None).
    cls_only (surrogate / classifier heads): the last block runs for the CLS query only and x is (rows, H).
    layer_hook(i, x, x_act, masks): called after every block with its output (fp32 residual stream | None, activation-dtype
    copy | None, the packed masks that go with the rows' token order) — the side ladders of the LTT variants tap the frozen
    backbone here; x is updated in place by the next block, so the hook must consume it on the same stream.
    token0_only (with a hook; the LTT surrogate ladders): everything downstream — the hook and the caller's heads — reads
    token 0 only and is equivariant to a per-row permutation of the tokens given the permuted masks, so the stack may run in
    kept-first order on the hi/lo residual stream; x is then the CLS rows (rows, H) and the hook gets x = None."""
    T = n_players_of(cfg) + 1
    H, heads, eps = cfg.hidden_size, cfg.num_attention_heads, cfg.layer_norm_eps
    cls_only = cls_only and CLS_ONLY_LAST_BLOCK and len(bw.layers) > 0 and layer_hook is None
    if (cls_only and not bw.vit and DROP_MASKED_TOKENS and pol.bf16 and T <= 512 and H == heads * 64
            and all(lw.ln2 is not None for lw in bw.layers)):
        return run_bert_packed(bw, cfg, pol, xs, masks, S)
    full = bw.layers[:-1] if cls_only else bw.layers
    fused = bw.vit and FUSE_LAYERNORM and pol.bf16 and H % 256 == 0 and all(lw.fold() is not None for lw in bw.layers)
    # First-block sharing (exact): before the first attention all S coalitions of an input hold identical activations,
    # so that block's LayerNorm + QKV projection runs once per input and the attention kernel reads the shared rows.
    share = (SHARE_FIRST_BLOCK and S > 1 and pol.bf16 and len(full) > 0 and T <= 512 and H == heads * 64)
    ctx0 = None
    nkeep = None             # set when the rows are switched to kept-first token order (KEPT_FIRST_ORDER)
    free_order = cls_only or (token0_only and layer_hook is not None)      # nobody downstream depends on the token order
    hilo = HILO_RESIDUAL and fused and free_order and len(full) > 0
    xh = xl = None           # hi/lo residual stream (see HILO_RESIDUAL)
    if share:
        x_img = embed(bw, cfg, pol, xs, 1)                                   # (B, T, H) fp32, one row block per input
        assert x_img.shape[1] == T, f"sequence length {x_img.shape[1]} != n_players + 1 = {T}"
        lw0 = full[0]
        xi = x_img.reshape(-1, H)
        if fused:
            xi16, sti = ops.rowstats_cast(xi)
            f0 = lw0.fold()
            qkv0 = ops.gemm_bf16_fused(xi16, f0["wqkv"], f0["bqkv"], ln=(sti, f0["cqkv"], eps))[0]
        elif bw.vit:
            h0 = pol.ln(xi, lw0.ln1[0], lw0.ln1[1], eps)[0] if lw0.ln1 is not None else pol.act(xi)
            qkv0 = pol.linear(h0, lw0.wqkv, lw0.bqkv)
        else:
            qkv0 = pol.linear(pol.act(xi), lw0.wqkv, lw0.bqkv)
        if (KEPT_FIRST_ORDER and free_order and fused and T <= 208 and len(full) > 1):
            # kept-first token order per row (stable: CLS stays first).  The first block's attention (projections shared per
            # input, masks in the original token order) scatters its output rows into that order, the residual stream of
            # every coalition is gathered from the per-input embeddings in that order, and from here on the masks are
            # prefixes of length nkeep: the attention kernels fold the masked tail into one virtual key.
            order, pos, nkeep, masks_p = ops.kept_first_order(masks, T)
            ctx0 = ops.masked_attention_scatter(qkv0, masks, T, heads, S, pos)
            if hilo:
                xh, xl = ops.gather_token_rows_hilo(x_img, order, S)
                x3 = xh
            else:
                x3 = ops.gather_token_rows(x_img, order, S)
            masks = masks_p
        else:
            ctx0 = ops.masked_attention(qkv0, masks, T, heads, ops.MASK_MUL0 if bw.vit else ops.MASK_NEGINF, share=S)
            x3 = ops.repeat_rows(x_img, S)                                   # the residual stream of every coalition
    else:
        x3 = embed(bw, cfg, pol, xs, S)
    assert x3.shape[1] == T, f"sequence length {x3.shape[1]} != n_players + 1 = {T}"
    rows = x3.shape[0]
    assert masks.shape[0] == rows, f"need one packed mask row per (input, coalition): {masks.shape[0]} vs {rows}"
    x = x3.reshape(rows * T, H)
    if hilo:
        stats = None
        if xh is None:
            if ctx0 is None:
                _, stats = ops.rowstats_cast(x)
            xh, xl = ops.split_hilo(x)
        xh, xl = xh.reshape(rows * T, H), xl.reshape(rows * T, H)
        for i, lw in enumerate(full):
            stats = vit_layer_hilo(lw, xh, xl, stats, T, heads, eps, masks, ctx=ctx0 if i == 0 else None, nkeep=nkeep)
            if layer_hook is not None:
                layer_hook(i, None, xh, masks)
        if not cls_only:         # token0_only: the caller's heads read the CLS rows
            return xh.view(rows, T, H)[:, 0, :].float() + xl.view(rows, T, H)[:, 0, :].float(), None
        return last_block_cls_only(pol, bw.layers[-1], True, None, None, xh, stats, masks, T, heads, eps, x_lo=xl)
    if bw.vit:
        if fused:
            x16 = stats = None
            if ctx0 is None:
                x16, stats = ops.rowstats_cast(x)
            for i, lw in enumerate(full):
                x, x16, stats = vit_layer_fused(lw, x, x16, stats, masks, T, heads, eps,
                                                last=(not cls_only and i == len(full) - 1 and layer_hook is None),
                                                ctx=ctx0 if i == 0 else None, nkeep=nkeep)
                if layer_hook is not None:
                    layer_hook(i, x, x16, masks)
            if cls_only:
                if x16 is None:      # single-block model: the CLS-only block is also the first one
                    x16, stats = ops.rowstats_cast(x)
                return last_block_cls_only(pol, bw.layers[-1], True, x, None, x16, stats, masks, T, heads, eps)
            return x, None
        for i, lw in enumerate(full):
            x = vit_layer(pol, lw, x, masks, T, heads, eps, ctx=ctx0 if i == 0 else None)
            if layer_hook is not None:
                layer_hook(i, x, None, masks)
        if cls_only:
            return last_block_cls_only(pol, bw.layers[-1], True, x, None, None, None, masks, T, heads, eps)
        return x, None
    xa = pol.act(x) if ctx0 is None else None
    for i, lw in enumerate(full):
        x, xa = bert_layer(pol, lw, x, xa, masks, T, heads, eps, ctx=ctx0 if i == 0 else None)
        if layer_hook is not None:
            layer_hook(i, x, xa, masks)
    if cls_only:
        if xa is None:
            xa = pol.act(x)
        return last_block_cls_only(pol, bw.layers[-1], False, x, xa, None, None, masks, T, heads, eps)
    return x, xa


# Small calls are launch-bound: the reference's scripts evaluate 2-8 inputs x 4-32 coalitions at a time, where the ~65 kernel
# launches of a ViT-Base evaluation cost more host time (~1.8 ms through Python + ctypes) than the GPU needs.  Calls of at most
# GRAPH_MAX_ROWS (input, coalition) rows are therefore captured once per shape into a CUDA graph and replayed (fixed-layout
# paths only: the packed BERT path sizes its buffers from a device-side count).  0 = off.
GRAPH_MAX_ROWS = 128
# (input, coalition) rows per pass of SurrogateEngine.probs: larger calls are processed in chunks of whole inputs
MAX_ROWS = int(os.environ.get("AGB_MAX_ROWS", "1024"))     # 1024 vs 4096: no measurable difference (A/B on one box)
GRAPH_CACHE_ENTRIES = 4


class _GraphCache:
    """shape-keyed CUDA graphs of a pure function of device tensors (static inputs copied in, static output cloned out)"""

    def __init__(self):
        self.entries: Dict[Any, Any] = {}

    def run(self, key, fn, inputs: Tuple[Tensor, ...]) -> Tensor:
        entry = self.entries.get(key)
        if entry is None:
            static_in = [t.clone() for t in inputs]
            side = torch.cuda.Stream(device=inputs[0].device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):            # warm-up off the capture: lazy allocations, function attributes
                for _ in range(2):
                    fn(*static_in)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = fn(*static_in)
            if len(self.entries) >= GRAPH_CACHE_ENTRIES:
                self.entries.pop(next(iter(self.entries)))
            entry = (graph, static_in, static_out)
            self.entries[key] = entry
        graph, static_in, static_out = entry
        for dst, src in zip(static_in, inputs):
            dst.copy_(src)
        graph.replay()
        return static_out.clone()


class SurrogateEngine:
    """Masked value function v(S): class probabilities for (input, coalition) rows.
    reference models/vanilla_vit.py:51-56, models/vanilla_bert.py:61-77"""

    def __init__(self, sd: State, cfg, precision: str):
        self.cfg = cfg
        self.pol = _Policy(precision)
        self.bw = BackboneWeights(sd, cfg, self.pol)
        self.w_cls, self.b_cls = _f32(sd["classifier.weight"]), _f32(sd["classifier.bias"])
        if not self.bw.vit:
            self.w_pool, self.b_pool = _f32(sd["bert_pooler.dense.weight"]), _f32(sd["bert_pooler.dense.bias"])
        self.graphs = _GraphCache()

    def _packed_path(self) -> bool:
        cfg = self.cfg
        return (not self.bw.vit and DROP_MASKED_TOKENS and self.pol.bf16 and CLS_ONLY_LAST_BLOCK
                and cfg.hidden_size == cfg.num_attention_heads * 64)

    @torch.no_grad()
    def probs(self, xs: Tensor, masks: Tensor, S: int, max_rows: int = 0) -> Tensor:
        """xs (B, ...), masks packed (B*S, words) -> (B*S, C) fp32 probabilities, row order b*S+s."""
        max_rows = max_rows or MAX_ROWS
        cfg, T = self.cfg, n_players_of(self.cfg) + 1
        B = xs.shape[0]
        assert masks.shape[0] == B * S
        if B * S == 0:
            return torch.empty((0, cfg.num_labels), dtype=torch.float32, device=xs.device)
        from . import _native as nat
        if (0 < B * S <= GRAPH_MAX_ROWS and B * S <= max_rows and self.pol.bf16 and not self._packed_path()
                and nat.PROFILE is None and not torch.cuda.is_current_stream_capturing()):
            key = (xs.device, tuple(xs.shape), xs.dtype, tuple(masks.shape), S, FUSE_LAYERNORM, CLS_ONLY_LAST_BLOCK,
                   SHARE_FIRST_BLOCK, KEPT_FIRST_ORDER, HILO_RESIDUAL)
            return self.graphs.run(key, lambda x_, m_: self._probs_eager(x_, m_, S, max_rows), (xs.contiguous(), masks))
        return self._probs_eager(xs, masks, S, max_rows)

    def _probs_eager(self, xs: Tensor, masks: Tensor, S: int, max_rows: int) -> Tensor:
        cfg, T = self.cfg, n_players_of(self.cfg) + 1
        B = xs.shape[0]
        per = max(1, max_rows // S)
        outs: List[Tensor] = []
        for b0 in range(0, B, per):
            b1 = min(B, b0 + per)
            x, _ = run_backbone(self.bw, cfg, self.pol, xs[b0:b1], masks[b0 * S:b1 * S], S, cls_only=True)
            x3 = x.reshape((b1 - b0) * S, -1, cfg.hidden_size)       # (rows, 1, H) with the CLS-only last block, else (rows, T, H)
            if self.bw.vit:
                outs.append(ops.cls_head(x3, 0, self.w_cls, self.b_cls, ln=(self.bw.final_ln[0], self.bw.final_ln[1], cfg.layer_norm_eps)))
            else:
                outs.append(ops.cls_head(x3, 1, self.w_cls, self.b_cls, pool=(self.w_pool, self.b_pool)))
        return outs[0] if len(outs) == 1 else torch.cat(outs, 0)


class ExplainerEngine:
    """Inference-time explainer: backbone -> explainer_attn -> explainer_mlp -> fused head/normalise.
    reference models/vanilla_vit.py:102-130, models/vanilla_bert.py:123-162"""

    def __init__(self, sd: State, cfg, precision: str):
        self.cfg = cfg
        self.pol = pol = _Policy(precision)
        self.bw = BackboneWeights(sd, cfg, pol)
        vit = self.bw.vit
        self.attn = [LayerWeights(sd, f"explainer_attn.{i}", pol, vit) for i in range(cfg.explainer_attn_num_layers)]
        if vit:
            self.mlp_ln = (_f32(sd["explainer_mlp.0.weight"]), _f32(sd["explainer_mlp.0.bias"]))
            names = ("explainer_mlp.1", "explainer_mlp.3", "explainer_mlp.5")
        else:
            self.mlp_ln = None
            names = ("explainer_mlp.0", "explainer_mlp.2", "explainer_mlp.4")
        self.w_a, self.b_a = pol.weight(sd[names[0] + ".weight"]), _f32(sd[names[0] + ".bias"])
        self.w_b, self.b_b = pol.weight(sd[names[1] + ".weight"]), _f32(sd[names[1] + ".bias"])
        self.w_c, self.b_c = _f32(sd[names[2] + ".weight"]), _f32(sd[names[2] + ".bias"])  # fp32 head

    @torch.no_grad()
    def phi(self, xs: Tensor, masks: Tensor, grand: Optional[Tensor], null: Optional[Tensor], want_pred: bool = False):
        """xs (B,...), masks packed (B, words) -> phi (B, C, n) fp32 [, pred (B,T,C)]"""
        if xs.shape[0] == 0:
            n, C = n_players_of(self.cfg), self.cfg.num_labels
            phi = torch.empty((0, C, n), dtype=torch.float32, device=xs.device)
            return (phi, torch.empty((0, n + 1, C), dtype=torch.float32, device=xs.device)) if want_pred else phi
        from . import _native as nat
        B = xs.shape[0]
        if (B <= GRAPH_MAX_ROWS and self.pol.bf16 and not want_pred and grand is not None and null is not None
                and nat.PROFILE is None and not torch.cuda.is_current_stream_capturing()):
            # small explainer calls are launch-bound as well: one CUDA graph per shape (see GRAPH_MAX_ROWS)
            if not hasattr(self, "graphs"):
                self.graphs = _GraphCache()
            key = (xs.device, tuple(xs.shape), xs.dtype, tuple(masks.shape), tuple(grand.shape), tuple(null.shape), FUSE_LAYERNORM)

            def fn(x_, m_, g_, n_):
                h, ha = run_backbone(self.bw, self.cfg, self.pol, x_, m_, 1)
                return self.tail(h, ha, m_, B, g_, n_, False)

            return self.graphs.run(key, fn, (xs.contiguous(), masks, grand.float().contiguous(), null.float().contiguous()))
        x, xa = run_backbone(self.bw, self.cfg, self.pol, xs, masks, 1)
        return self.tail(x, xa, masks, xs.shape[0], grand, null, want_pred)

    @torch.no_grad()
    def tail(self, x: Tensor, xa: Optional[Tensor], masks: Tensor, B: int, grand: Optional[Tensor], null: Optional[Tensor],
             want_pred: bool = False):
        """Everything after the encoder stack: x (B*T, H) fp32 [ViT: before vit.layernorm], xa its activation copy."""
        cfg, pol = self.cfg, self.pol
        T = n_players_of(cfg) + 1
        heads, eps = cfg.num_attention_heads, cfg.layer_norm_eps
        if self.bw.vit:
            _, x = ops.layernorm(x, self.bw.final_ln[0], self.bw.final_ln[1], eps, want_bf16=False, want_f32=True)
            for lw in self.attn:
                x = vit_layer(pol, lw, x, masks, T, heads, eps)
            h = pol.ln(x, self.mlp_ln[0], self.mlp_ln[1], 1e-5)[0]  # nn.LayerNorm default eps (reference l.94)
        else:
            for lw in self.attn:
                x, xa = bert_layer(pol, lw, x, xa, masks, T, heads, eps)
            h = xa
        h = pol.linear(h, self.w_a, self.b_a, act=ops.ACT_GELU)
        h = pol.linear(h, self.w_b, self.b_b, act=ops.ACT_GELU)
        return ops.explainer_head_fwd(h, B, T, self.w_c, self.b_c, grand, null, bool(cfg.explainer_normalize), want_pred)


class DuoExplainerEngine(ExplainerEngine):
    """Duo explainer (reference models/duo_vanilla_vit.py:62-123, duo_vanilla_bert.py:72-148): the explainer's own backbone
    also feeds a classification head.  -> (phi, class output): ViT softmax probabilities, BERT raw logits."""

    def __init__(self, sd: State, cfg, precision: str):
        super().__init__(sd, cfg, precision)
        self.w_cls, self.b_cls = _f32(sd["classifier.weight"]), _f32(sd["classifier.bias"])
        if not self.bw.vit:
            self.pool = (_f32(sd["bert_pooler.dense.weight"]), _f32(sd["bert_pooler.dense.bias"]))

    @torch.no_grad()
    def duo(self, xs: Tensor, masks: Tensor, grand: Optional[Tensor], null: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
        cfg = self.cfg
        B = xs.shape[0]
        if B == 0:
            return _empty_outputs(cfg, xs.device)[::-1]
        x, xa = run_backbone(self.bw, cfg, self.pol, xs, masks, 1)
        x3 = x.reshape(B, -1, cfg.hidden_size)
        if self.bw.vit:
            cls = ops.cls_head(x3, 0, self.w_cls, self.b_cls, ln=(self.bw.final_ln[0], self.bw.final_ln[1], cfg.layer_norm_eps))
        else:
            cls = ops.cls_head(x3, 1, self.w_cls, self.b_cls, pool=self.pool, want_logits=True)[1]
        return self.tail(x, xa, masks, B, grand, null), cls


class FroyoFinalEngine(ExplainerEngine):
    """Froyo bundle (reference models/froyo_vit.py:100-171, models/froyo_bert.py:103-204): ONE frozen backbone pass feeds
    the classifier head, the surrogate head (srg_*) and the explainer tail."""

    def __init__(self, sd: State, cfg, precision: str):
        super().__init__(sd, cfg, precision)
        self.w_cls, self.b_cls = _f32(sd["classifier.weight"]), _f32(sd["classifier.bias"])
        self.w_srg, self.b_srg = _f32(sd["srg_classifier.weight"]), _f32(sd["srg_classifier.bias"])
        if not self.bw.vit:
            self.pool = (_f32(sd["bert_pooler.dense.weight"]), _f32(sd["bert_pooler.dense.bias"]))
            self.srg_pool = (_f32(sd["srg_bert_pooler.dense.weight"]), _f32(sd["srg_bert_pooler.dense.bias"]))
        self.null = _f32(sd["surrogate_null"])

    @torch.no_grad()
    def final(self, xs: Tensor, masks: Tensor) -> Tuple[Tensor, Tensor]:
        cfg = self.cfg
        B = xs.shape[0]
        if B == 0:
            return _empty_outputs(cfg, xs.device)
        x, xa = run_backbone(self.bw, cfg, self.pol, xs, masks, 1)
        x3 = x.reshape(B, -1, cfg.hidden_size)
        if self.bw.vit:
            ln = (self.bw.final_ln[0], self.bw.final_ln[1], cfg.layer_norm_eps)
            cls = ops.cls_head(x3, 0, self.w_cls, self.b_cls, ln=ln)
            grand = ops.cls_head(x3, 0, self.w_srg, self.b_srg, ln=ln) if cfg.explainer_normalize else None
        else:
            cls = ops.cls_head(x3, 1, self.w_cls, self.b_cls, pool=self.pool)
            grand = ops.cls_head(x3, 1, self.w_srg, self.b_srg, pool=self.srg_pool) if cfg.explainer_normalize else None
        phi = self.tail(x, xa, masks, B, grand, self.null if cfg.explainer_normalize else None)
        return cls, phi


# ------------------------------------------------------------------------------------------------------------------
# LTT ("ladder side tuning", the paper's method): the backbone is frozen and every block's output is
# mapped into a narrow side ladder  s <- block_i^side(s + GELU(W_i h_i))  (reference models/ltt_vit.py:407-440,
# models/ltt_bert.py:467-499).  Surrogate and explainer are side ladders over the SAME backbone activations, so the
# bundle evaluates the backbone once.
# ------------------------------------------------------------------------------------------------------------------
def _empty_outputs(cfg, device) -> Tuple[Tensor, Tensor]:
    """(probabilities (0, C), attributions (0, C, n)) of an empty batch"""
    n, C = n_players_of(cfg), cfg.num_labels
    return (torch.empty((0, C), dtype=torch.float32, device=device), torch.empty((0, C, n), dtype=torch.float32, device=device))


class SideBranch:
    """One side ladder in kernel-ready form: H -> Hs maps, Hs-wide blocks, [ViT] its final LayerNorm."""

    def __init__(self, sd: State, cfg, pol: _Policy, vit: bool, b: int):
        root = "vit" if vit else "bert"
        L = cfg.num_hidden_layers
        self.maps = [(pol.weight(sd[f"{root}.encoder.s_attn_maps.{b}_{i}.weight"]),
                      _f32(sd[f"{root}.encoder.s_attn_maps.{b}_{i}.bias"])) for i in range(L)]
        self.layers = [LayerWeights(sd, f"{root}.encoder.s_attn_layers.{b}_{i}", pol, vit) for i in range(L)]
        self.final_ln = (_f32(sd[f"vit.s_attn_layernorm.{b}.weight"]), _f32(sd[f"vit.s_attn_layernorm.{b}.bias"])) if vit else None


def run_ltt_bert_packed(bw: BackboneWeights, branches: List[SideBranch], cfg, pol: _Policy, xs: Tensor, masks: Tensor, S: int,
                        stop: int):
    """LTT surrogate evaluation for additive (-inf) masks on the kept tokens only (SURVEY.md 8f-3 applied to the ladder): a
    masked token is never attended to in the backbone NOR in the ladder (both use the same key mask) and both heads read
    token 0, so neither output can depend on it.  Backbone and ladder run on the packed rows with variable-length
    attention.  -> (backbone CLS rows (rows, H) fp32, [ladder CLS rows (rows, Hs) fp32 per branch])"""
    T = n_players_of(cfg) + 1
    H, heads, eps = cfg.hidden_size, cfg.num_attention_heads, cfg.layer_norm_eps
    cu, src, total = ops.pack_kept_tokens(masks, T, S)
    x = embed(bw, cfg, pol, xs, 1).reshape(-1, H).index_select(0, src)       # (total, H) packed residual stream
    xa = pol.act(x)
    side: List[Optional[Tensor]] = [None] * len(branches)
    for i, lw in enumerate(bw.layers):
        ctx = ops.attention_varlen(pol.linear(xa, lw.wqkv, lw.bqkv), cu, T, heads)
        x, xa = bert_layer(pol, lw, x, xa, masks, T, heads, eps, ctx=ctx)
        if i >= stop:
            continue
        for k, br in enumerate(branches):
            w, b = br.maps[i]
            slw = br.layers[i]
            s = pol.linear(xa, w, b, act=ops.ACT_GELU, residual=side[k], out_f32=True)
            sa = pol.act(s)
            sctx = ops.attention_varlen(pol.linear(sa, slw.wqkv, slw.bqkv), cu, T, heads)
            side[k], _ = bert_layer(pol, slw, s, sa, masks, T, heads, eps, ctx=sctx)
    first = cu[:-1].long()                                                   # packed index of every row's CLS token
    return x.index_select(0, first), [s.index_select(0, first) for s in side]


def run_ltt(bw: BackboneWeights, branches: List[SideBranch], cfg, pol: _Policy, xs: Tensor, masks: Tensor, S: int,
            freeze_layer: Optional[int] = None, token0_only: bool = False):
    """-> (x, xa of the backbone as run_backbone, [(side state (rows*T, Hs) fp32, activation copy | None) per branch])
    freeze_layer: reference `_ltt_freeze_layer` — blocks >= it do not feed the ladders."""
    T = n_players_of(cfg) + 1
    heads, eps = cfg.num_attention_heads, cfg.layer_norm_eps
    L = len(bw.layers)
    stop = L if freeze_layer is None else max(1, min(L, int(freeze_layer)))
    side: List[Optional[Tensor]] = [None] * len(branches)
    side_a: List[Optional[Tensor]] = [None] * len(branches)

    def hook(i: int, x: Optional[Tensor], x_act: Optional[Tensor], m: Tensor) -> None:
        # m: the masks in the token order the backbone runs in (kept-first order permutes the tokens of every row)
        if i >= stop:
            return
        if x_act is None:
            x_act = pol.act(x)
        for k, br in enumerate(branches):
            w, b = br.maps[i]
            s = pol.linear(x_act, w, b, act=ops.ACT_GELU, residual=side[k], out_f32=True)
            if bw.vit:
                side[k] = vit_layer(pol, br.layers[i], s, m, T, heads, eps)
            else:
                side[k], side_a[k] = bert_layer(pol, br.layers[i], s, pol.act(s), m, T, heads, eps)

    x, xa = run_backbone(bw, cfg, pol, xs, masks, S, layer_hook=hook, token0_only=token0_only)
    return x, xa, list(zip(side, side_a))


class LttEngine:
    """LTT surrogate / explainer / bundle on one backbone pass.  `heads` selects what exists in the state dict:
    "surrogate" (side ladder 0 + s_attn_classifier), "explainer" (side ladder 0 + s_explainer_*), "final" (ladder 0 =
    surrogate, ladder 1 = explainer).  reference models/ltt_vit.py:55-287, models/ltt_bert.py:66-327."""

    def __init__(self, sd: State, cfg, precision: str, kind: str):
        assert kind in ("surrogate", "explainer", "final")
        self.cfg, self.kind = cfg, kind
        self.pol = pol = _Policy(precision)
        self.bw = BackboneWeights(sd, cfg, pol)
        vit = self.bw.vit
        self.branches = [SideBranch(sd, cfg, pol, vit, b) for b in range(2 if kind == "final" else 1)]
        self.w_cls, self.b_cls = _f32(sd["classifier.weight"]), _f32(sd["classifier.bias"])
        if not vit:
            self.pool = (_f32(sd["bert_pooler.dense.weight"]), _f32(sd["bert_pooler.dense.bias"]))
        if kind in ("surrogate", "final"):
            self.w_srg, self.b_srg = _f32(sd["s_attn_classifier.weight"]), _f32(sd["s_attn_classifier.bias"])
            if not vit:
                self.srg_pool = (_f32(sd["bert_s_attn_pooler.dense.weight"]), _f32(sd["bert_s_attn_pooler.dense.bias"]))
        if kind in ("explainer", "final"):
            attn = "s_explainer_attn" if vit else "s_attn_attention_layers"     # the two families name them differently
            self.attn = [LayerWeights(sd, f"{attn}.{i}", pol, vit) for i in range(cfg.explainer_s_attn_num_layers)]
            if vit:
                self.mlp_ln = (_f32(sd["s_explainer_mlp.0.weight"]), _f32(sd["s_explainer_mlp.0.bias"]))
                names = ("s_explainer_mlp.1", "s_explainer_mlp.3", "s_explainer_mlp.5")
            else:
                self.mlp_ln = None
                names = ("s_attn_explainer.0", "s_attn_explainer.2", "s_attn_explainer.4")
            self.w_a, self.b_a = pol.weight(sd[names[0] + ".weight"]), _f32(sd[names[0] + ".bias"])
            self.w_b, self.b_b = pol.weight(sd[names[1] + ".weight"]), _f32(sd[names[1] + ".bias"])
            self.w_c, self.b_c = _f32(sd[names[2] + ".weight"]), _f32(sd[names[2] + ".bias"])
        if kind == "final":
            self.null = _f32(sd["surrogate_null"])
        self.freeze_layer: Optional[int] = None

    def _main_probs(self, x: Tensor, rows: int) -> Tensor:
        cfg = self.cfg
        x3 = x.reshape(rows, -1, cfg.hidden_size)
        if self.bw.vit:
            return ops.cls_head(x3, 0, self.w_cls, self.b_cls, ln=(self.bw.final_ln[0], self.bw.final_ln[1], cfg.layer_norm_eps))
        return ops.cls_head(x3, 1, self.w_cls, self.b_cls, pool=self.pool)

    def _side_probs(self, s: Tensor, rows: int) -> Tensor:
        cfg = self.cfg
        s3 = s.reshape(rows, -1, cfg.s_attn_hidden_size)
        if self.bw.vit:
            g, b = self.branches[0].final_ln
            return ops.cls_head(s3, 0, self.w_srg, self.b_srg, ln=(g, b, cfg.layer_norm_eps))
        return ops.cls_head(s3, 1, self.w_srg, self.b_srg, pool=self.srg_pool)

    def _explain(self, br: SideBranch, s: Tensor, sa: Optional[Tensor], masks: Tensor, B: int, grand, null) -> Tensor:
        cfg, pol = self.cfg, self.pol
        T = n_players_of(cfg) + 1
        heads, eps = cfg.num_attention_heads, cfg.layer_norm_eps
        if self.bw.vit:
            _, s = ops.layernorm(s, br.final_ln[0], br.final_ln[1], eps, want_bf16=False, want_f32=True)
            for lw in self.attn:
                s = vit_layer(pol, lw, s, masks, T, heads, eps)
            h = pol.ln(s, self.mlp_ln[0], self.mlp_ln[1], 1e-5)[0]   # nn.LayerNorm default eps (reference ltt_vit.py:124)
        else:
            for lw in self.attn:
                s, sa = bert_layer(pol, lw, s, sa, masks, T, heads, eps)
            h = sa
        h = pol.linear(h, self.w_a, self.b_a, act=ops.ACT_GELU)
        h = pol.linear(h, self.w_b, self.b_b, act=ops.ACT_GELU)
        return ops.explainer_head_fwd(h, B, T, self.w_c, self.b_c, grand, null, bool(cfg.explainer_normalize))

    @torch.no_grad()
    def surrogate(self, xs: Tensor, masks: Tensor, S: int, max_rows: int = 1024) -> Tuple[Tensor, Tensor]:
        """-> (side-ladder probabilities (B*S, C), backbone probabilities (B*S, C)), row order b*S+s."""
        B = xs.shape[0]
        assert masks.shape[0] == B * S
        if B * S == 0:
            e = _empty_outputs(self.cfg, xs.device)[0]
            return e, e.clone()
        per = max(1, max_rows // S)
        srg, cls = [], []
        cfg = self.cfg
        L = len(self.bw.layers)
        stop = L if self.freeze_layer is None else max(1, min(L, int(self.freeze_layer)))
        hs = cfg.s_attn_hidden_size // cfg.num_attention_heads
        packed = (not self.bw.vit and DROP_MASKED_TOKENS and self.pol.bf16 and n_players_of(cfg) + 1 <= 512
                  and cfg.hidden_size == cfg.num_attention_heads * 64 and hs in (8, 16, 32) and cfg.s_attn_hidden_size % 8 == 0)
        for b0 in range(0, B, per):
            b1 = min(B, b0 + per)
            rows = (b1 - b0) * S
            if packed:
                x, sides_cls = run_ltt_bert_packed(self.bw, self.branches[:1], cfg, self.pol, xs[b0:b1], masks[b0 * S:b1 * S], S, stop)
                cls.append(self._main_probs(x, rows))
                srg.append(self._side_probs(sides_cls[0], rows))
                continue
            # both heads read token 0 and the ladder is token-wise + attention: free to run in kept-first order (hi/lo stream)
            x, _, sides = run_ltt(self.bw, self.branches[:1], self.cfg, self.pol, xs[b0:b1], masks[b0 * S:b1 * S], S, self.freeze_layer,
                                  token0_only=True)
            cls.append(self._main_probs(x, rows))
            srg.append(self._side_probs(sides[0][0], rows))
        return (srg[0], cls[0]) if len(srg) == 1 else (torch.cat(srg, 0), torch.cat(cls, 0))

    @torch.no_grad()
    def explainer(self, xs: Tensor, masks: Tensor, grand: Optional[Tensor], null: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
        """-> (phi (B, C, n), backbone probabilities (B, C))"""
        B = xs.shape[0]
        if B == 0:
            return _empty_outputs(self.cfg, xs.device)[::-1]
        x, _, sides = run_ltt(self.bw, self.branches[:1], self.cfg, self.pol, xs, masks, 1, self.freeze_layer)
        cls = self._main_probs(x, B)
        return self._explain(self.branches[0], sides[0][0], sides[0][1], masks, B, grand, null), cls

    @torch.no_grad()
    def final(self, xs: Tensor, masks: Tensor) -> Tuple[Tensor, Tensor]:
        """-> (backbone probabilities (B, C), phi (B, C, n)); the surrogate ladder supplies `grand` when normalising."""
        cfg = self.cfg
        B = xs.shape[0]
        if B == 0:
            return _empty_outputs(cfg, xs.device)
        use = self.branches if cfg.explainer_normalize else self.branches[1:]
        x, _, sides = run_ltt(self.bw, use, cfg, self.pol, xs, masks, 1, self.freeze_layer)
        cls = self._main_probs(x, B)
        grand = self._side_probs(sides[0][0], B) if cfg.explainer_normalize else None
        s, sa = sides[-1]
        return cls, self._explain(self.branches[1], s, sa, masks, B, grand, self.null if cfg.explainer_normalize else None)
