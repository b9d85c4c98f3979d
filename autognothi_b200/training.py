"""Explainer training path: forward with a tape + hand-written backward over the CUDA kernels.

The reference trains the explainer with torch autograd (scripts/train_explainer.py:182-198 over
models/vanilla_vit.py:102-130 / models/vanilla_bert.py:123-162).  Here the whole explainer is ONE
autograd node: `explainer_forward_train` returns phi attached to the graph, and `loss.backward()` runs
the adjoint below — every Linear's dgrad/wgrad on the tcgen05 GEMM (MN-major operand modes, no
transposed copies), attention / LayerNorm / GELU / head adjoints on their own kernels.  So the
reference's loop (`phi = fw_explainer(...); loss = loss_shapley_new(...); loss.backward();
optimizer.step()`) runs unchanged.

Dropout: the reference keeps hidden / attention-probability dropout (p = 0.1 in the checked-in configs) active while
training.  In train() mode this path applies both (`_Drop`; bf16 on the tcgen05 kernels, fp32 on the CUDA-core ones): elementwise sites through `agb_dropout`, the
attention probabilities inside the tcgen05 attention forward and adjoint, masks regenerated from a counter hash in the
backward pass.  `module.agb_dropout = False` gives the deterministic p = 0 path that the parity tests compare with the
reference run in eval() mode (where its dropout is the identity); the random streams differ from torch's, so dropout
itself is checked against a torch re-statement that uses the exported masks (tests/test_gpu_dropout.py).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor

from . import engine, ops
from .engine import LayerWeights, _Policy, _f32, n_players_of

Grads = Dict[str, Tensor]


class _ZeroArena:
    """fp32 zeros for the small accumulate-into gradients of one backward pass (bias column sums, LayerNorm gamma / beta): slices
    of one zero-filled chunk, i.e. one memset instead of ~110 fill launches per ViT-Base step.  The slices are handed to autograd
    as they are (contiguous views; the chunk lives as long as any of them)."""
    CHUNK = 1 << 18

    def __init__(self):
        self.buf: Optional[Tensor] = None
        self.off = 0

    def take(self, shape, device) -> Tensor:
        n = 1
        for d in shape:
            n *= int(d)
        step = (n + 63) // 64 * 64          # 256-byte aligned slices
        if step > self.CHUNK:
            return torch.zeros(tuple(shape), dtype=torch.float32, device=device)
        if self.buf is None or self.buf.device != device or self.off + step > self.buf.numel():
            self.buf = torch.zeros((self.CHUNK,), dtype=torch.float32, device=device)
            self.off = 0
        out = self.buf[self.off:self.off + n].view(tuple(shape))
        self.off += step
        return out


_ARENA: Optional[_ZeroArena] = None     # set for the duration of a backward pass (autograd Function.backward below)


def _zeros(shape, device) -> Tensor:
    if _ARENA is not None:
        return _ARENA.take(shape, device)
    return torch.zeros(tuple(shape), dtype=torch.float32, device=device)


class _arena_scope:
    def __enter__(self):
        global _ARENA
        self.prev, _ARENA = _ARENA, _ZeroArena()

    def __exit__(self, *exc):
        global _ARENA
        _ARENA = self.prev


class _Drop:
    """Dropout state of one training forward (reference: nn.Dropout modules active in train() mode, hidden_dropout_prob on
    embeddings / attention-output / MLP-output, attention_probs_dropout_prob on the attention probabilities).  Masks come
    from a counter hash of (seed, site tag, element), so the tape stores two integers per site instead of a mask."""

    def __init__(self, cfg, enabled: bool):
        self.h = ops.dropout_thr(float(cfg.hidden_dropout_prob)) if enabled else 0
        self.a = ops.dropout_thr(float(cfg.attention_probs_dropout_prob)) if enabled else 0
        # CPU generator: follows torch.manual_seed and costs no device synchronisation; the data-parallel rank is mixed in
        # so that identically seeded replicas still draw different dropout masks
        self.seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if (self.h or self.a) else 0
        if self.seed and torch.distributed.is_available() and torch.distributed.is_initialized():
            self.seed = (self.seed ^ (0x51ED270B7F4A7C15 * (torch.distributed.get_rank() + 1))) & (2 ** 62 - 1)
        self.n = 0

    def tag(self) -> int:
        self.n += 1
        return self.n

    def site_seed(self, tag: int) -> int:
        return (self.seed + tag * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF


_NO_DROP = None


# bf16 training: residual + dropout(dense) inside the GEMM's epilogue (agb_gemm_bf16_dropout_residual) instead of a bf16
# GEMM output followed by agb_dropout; same mask stream, the dense output is no longer rounded to bf16 in between.
FUSE_DROPOUT = os.environ.get("AGB_FUSE_DROPOUT", "1") != "0"


def _linear_drop(pol: _Policy, drop, a: Tensor, w: Tensor, b: Tensor, residual: Tensor) -> Tuple[Tensor, int]:
    """residual + dropout(a @ w^T + b) -> (fp32, site tag | 0): with p = 0 the residual rides in the GEMM epilogue."""
    if drop is None or not drop.h:
        return pol.linear(a, w, b, residual=residual, out_f32=True), 0
    tag = drop.tag()
    if pol.bf16 and FUSE_DROPOUT and ops.gemm_dropout_residual_supported(a.shape[0], w.shape[0]):
        return ops.gemm_bf16_dropout_residual(a, w, b, residual, drop.h, drop.seed, tag), tag
    return ops.dropout(pol.linear(a, w, b), drop.h, drop.seed, tag, residual=residual, out_dtype=torch.float32), tag


def _act_drop(pol: _Policy, drop, dx: Tensor, tag: int) -> Tensor:
    """gradient w.r.t. the dense output behind a dropout site, in the activation dtype"""
    if not tag:
        return pol.act(dx)
    return ops.dropout(dx, drop.h, drop.seed, tag, out_dtype=pol.act_dtype)


def _attention_fwd(drop, qkv: Tensor, masks: Tensor, T: int, heads: int, mode: int) -> Tuple[Tensor, int]:
    if drop is None or not drop.a:
        return ops.masked_attention(qkv, masks, T, heads, mode), 0
    tag = drop.tag()
    return ops.masked_attention_dropout(qkv, masks, T, heads, mode, drop.a, drop.site_seed(tag)), tag


def _attention_bwd(drop, tag: int, qkv: Tensor, dctx: Tensor, masks: Tensor, T: int, heads: int, mode: int) -> Tensor:
    if not tag:
        return ops.masked_attention_bwd(qkv, dctx, masks, T, heads, mode)
    return ops.masked_attention_dropout_bwd(qkv, dctx, masks, T, heads, mode, drop.a, drop.site_seed(tag))


def _wgrad(pol: _Policy, dy: Tensor, x: Tensor) -> Tensor:
    """dW[N_out, K_in] = dy^T @ x   (dy [M, N_out], x [M, K_in]; both read as MN-major operands)"""
    if pol.bf16:
        return ops.gemm_bf16(dy, x, a_mn=True, w_mn=True, out_dtype=torch.float32)
    return ops.gemm_f32(dy, x, a_mn=True, w_mn=True)


def _dgrad(pol: _Policy, dy: Tensor, w: Tensor, *, out_f32: bool, residual: Optional[Tensor] = None) -> Tensor:
    """dx[M, K_in] = dy @ W   (W stored [N_out, K_in] = MN-major B operand)"""
    if pol.bf16:
        return ops.gemm_bf16(dy, w, w_mn=True, residual=residual, out_dtype=torch.float32 if out_f32 else torch.bfloat16)
    return ops.gemm_f32(dy, w, w_mn=True, residual=residual)


# bf16 training: the GELU forward rides in the epilogue of the GEMM in front of it (two outputs: pre-activation for the adjoint,
# activation for the next GEMM) and the GELU adjoint in the epilogue of the dgrad GEMM behind it, instead of separate
# elementwise kernels (agb_gemm_bf16_gelu_dual / agb_gemm_bf16_gelu_bwd).  Wide layers only (N >= 192: the tcgen05 pair kernel).
FUSE_GELU = os.environ.get("AGB_FUSE_GELU", "1") != "0"            # environment switches: A/B runs of bench.py


def _linear_gelu(pol: _Policy, a: Tensor, w: Tensor, b: Tensor) -> Tuple[Tensor, Tensor]:
    """-> (z = a @ w^T + b, GELU(z))"""
    if pol.bf16 and FUSE_GELU and w.shape[0] >= 192 and w.shape[0] % 8 == 0:
        return ops.gemm_bf16_gelu_dual(a, w, b)
    z = pol.linear(a, w, b)
    return z, ops.gelu_fwd(z)


def _dgrad_gelu(pol: _Policy, dy: Tensor, w: Tensor, z: Tensor) -> Tensor:
    """dz = (dy @ W) * GELU'(z)   (W [N_out, K_in]: the layer behind the GELU; z its input's pre-activation)"""
    if pol.bf16 and FUSE_GELU and w.shape[1] >= 192 and w.shape[1] % 8 == 0:
        return ops.gemm_bf16_gelu_bwd(dy, w, z)
    return ops.gelu_bwd(_dgrad(pol, dy, w, out_f32=False), z)


def _bias_grad(dy: Tensor) -> Tensor:
    g = _zeros((dy.shape[1],), dy.device)
    ops.colsum_into(dy, g)
    return g


def _linear_bwd(pol: _Policy, grads: Grads, name: str, dy: Tensor, x: Tensor) -> None:
    grads[name + ".weight"] = _wgrad(pol, dy, x)
    grads[name + ".bias"] = _bias_grad(dy)


def _qkv_bwd(pol: _Policy, grads: Grads, prefix: str, dqkv: Tensor, x: Tensor, H: int) -> None:
    dw = _wgrad(pol, dqkv, x)
    db = _bias_grad(dqkv)
    for i, nm in enumerate(("query", "key", "value")):
        grads[f"{prefix}.attention.self.{nm}.weight"] = dw[i * H:(i + 1) * H]
        grads[f"{prefix}.attention.self.{nm}.bias"] = db[i * H:(i + 1) * H]


def _ln_bwd(grads: Grads, name: str, x: Tensor, dy: Tensor, gamma: Tensor, eps: float, dres: Optional[Tensor]) -> Tensor:
    assert gamma.dtype == torch.float32
    dg = _zeros(gamma.shape, gamma.device)
    db = _zeros(gamma.shape, gamma.device)
    dx = ops.layernorm_bwd(x, dy, gamma, eps, dres, dg, db)
    grads[name + ".weight"], grads[name + ".bias"] = dg, db
    return dx


# ------------------------------------------------------------------------------------------------
# ViT block (pre-LN), reference models/vanilla_vit.py:364-377
# ------------------------------------------------------------------------------------------------
def vit_layer_fwd(pol, lw: LayerWeights, x: Tensor, masks: Tensor, T: int, heads: int, eps: float, drop=None):
    h1 = pol.ln(x, lw.ln1[0], lw.ln1[1], eps)[0] if lw.ln1 is not None else pol.act(x)
    qkv = pol.linear(h1, lw.wqkv, lw.bqkv)
    ctx, ta = _attention_fwd(drop, qkv, masks, T, heads, ops.MASK_MUL0)
    x_mid, t1 = _linear_drop(pol, drop, ctx, lw.wo, lw.bo, x)
    h2 = pol.ln(x_mid, lw.ln2[0], lw.ln2[1], eps)[0]
    z, f = _linear_gelu(pol, h2, lw.w1, lw.b1)
    x_out, t2 = _linear_drop(pol, drop, f, lw.w2, lw.b2, x_mid)
    return x_out, dict(x_in=x, h1=h1, qkv=qkv, ctx=ctx, x_mid=x_mid, h2=h2, z=z, f=f, drop=drop, tags=(ta, t1, t2))


def vit_layer_bwd(pol, lw: LayerWeights, prefix: str, t: dict, dx_out: Tensor, masks: Tensor, T: int, heads: int,
                  eps: float, grads: Grads, need_dx: bool = True) -> Optional[Tensor]:
    H = dx_out.shape[1]
    drop, (ta, t1, t2) = t["drop"], t["tags"]
    g = _act_drop(pol, drop, dx_out, t2)
    _linear_bwd(pol, grads, prefix + ".output.dense", g, t["f"])
    dz = _dgrad_gelu(pol, g, lw.w2, t["z"])
    _linear_bwd(pol, grads, prefix + ".intermediate.dense", dz, t["h2"])
    dh2 = _dgrad(pol, dz, lw.w1, out_f32=True)
    dx_mid = _ln_bwd(grads, prefix + ".layernorm_after", t["x_mid"], dh2, lw.ln2[0], eps, dx_out)
    g = _act_drop(pol, drop, dx_mid, t1)
    _linear_bwd(pol, grads, prefix + ".attention.output.dense", g, t["ctx"])
    dctx = _dgrad(pol, g, lw.wo, out_f32=False)
    dqkv = _attention_bwd(drop, ta, t["qkv"], dctx, masks, T, heads, ops.MASK_MUL0)
    _qkv_bwd(pol, grads, prefix, dqkv, t["h1"], H)
    if not need_dx and lw.ln1 is None:
        return None          # nothing trainable below this block (frozen backbone)
    if lw.ln1 is not None:
        dh1 = _dgrad(pol, dqkv, lw.wqkv, out_f32=True)
        return _ln_bwd(grads, prefix + ".layernorm_before", t["x_in"], dh1, lw.ln1[0], eps, dx_mid)
    return _dgrad(pol, dqkv, lw.wqkv, out_f32=True, residual=dx_mid)


# ------------------------------------------------------------------------------------------------
# BERT block (post-LN), reference models/vanilla_bert.py:396-427, 556-560, 600-604
# ------------------------------------------------------------------------------------------------
def bert_layer_fwd(pol, lw: LayerWeights, x: Tensor, xa: Tensor, masks: Tensor, T: int, heads: int, eps: float, drop=None):
    qkv = pol.linear(xa, lw.wqkv, lw.bqkv)
    ctx, ta = _attention_fwd(drop, qkv, masks, T, heads, ops.MASK_NEGINF)
    a_pre, t1 = _linear_drop(pol, drop, ctx, lw.wo, lw.bo, x)
    if lw.ln1 is not None:
        aa, a = pol.ln(a_pre, lw.ln1[0], lw.ln1[1], eps, want_f32=True)
    else:
        a, aa = a_pre, pol.act(a_pre)
    z, f = _linear_gelu(pol, aa, lw.w1, lw.b1)
    y_pre, t2 = _linear_drop(pol, drop, f, lw.w2, lw.b2, a)
    ya, y = pol.ln(y_pre, lw.ln2[0], lw.ln2[1], eps, want_f32=True)
    return y, ya, dict(xa=xa, qkv=qkv, ctx=ctx, a_pre=a_pre, aa=aa, z=z, f=f, y_pre=y_pre, drop=drop, tags=(ta, t1, t2))


def bert_layer_bwd(pol, lw: LayerWeights, prefix: str, t: dict, dy: Tensor, masks: Tensor, T: int, heads: int,
                   eps: float, grads: Grads, need_dx: bool = True) -> Optional[Tensor]:
    H = dy.shape[1]
    drop, (ta, t1, t2) = t["drop"], t["tags"]
    d_ypre = _ln_bwd(grads, prefix + ".output.LayerNorm", t["y_pre"], dy, lw.ln2[0], eps, None)
    g = _act_drop(pol, drop, d_ypre, t2)
    _linear_bwd(pol, grads, prefix + ".output.dense", g, t["f"])
    dz = _dgrad_gelu(pol, g, lw.w2, t["z"])
    _linear_bwd(pol, grads, prefix + ".intermediate.dense", dz, t["aa"])
    da = _dgrad(pol, dz, lw.w1, out_f32=True, residual=d_ypre)
    d_apre = _ln_bwd(grads, prefix + ".attention.output.LayerNorm", t["a_pre"], da, lw.ln1[0], eps, None) \
        if lw.ln1 is not None else da
    g = _act_drop(pol, drop, d_apre, t1)
    _linear_bwd(pol, grads, prefix + ".attention.output.dense", g, t["ctx"])
    dctx = _dgrad(pol, g, lw.wo, out_f32=False)
    dqkv = _attention_bwd(drop, ta, t["qkv"], dctx, masks, T, heads, ops.MASK_NEGINF)
    _qkv_bwd(pol, grads, prefix, dqkv, t["xa"], H)
    if not need_dx:
        return None          # nothing trainable below this block (frozen backbone)
    return _dgrad(pol, dqkv, lw.wqkv, out_f32=True, residual=d_apre)


def _embed_fwd(tp, bw, cfg, pol, xs: Tensor, drop=None):
    """Embeddings with S = 1 (one mask row per input); keeps what the adjoint needs on the tape.  The embedding dropout
    (reference models/vanilla_vit.py:253, models/vanilla_bert.py:325) is applied here when `drop` is active."""
    T, H, eps = tp.T, cfg.hidden_size, cfg.layer_norm_eps
    B = xs.shape[0]
    tp.drop, tp.embed_tag = drop, 0
    if bw.vit:
        tp.patches = ops.vit_im2col(xs.float(), cfg.img_patch_size, pol.act_dtype)
        pe = pol.linear(tp.patches, bw.w_patch, bw.b_patch, out_f32=True)
        x = ops.vit_assemble(pe, bw.cls_token, bw.pos_emb, B, 1, T, H).reshape(B * T, H)
    else:
        x = ops.bert_embed(xs, bw.word, bw.pos, bw.type0, bw.emb_ln[0], bw.emb_ln[1], eps, 1).reshape(B * T, H)
    if drop is not None and drop.h:
        tp.embed_tag = drop.tag()
        x = ops.dropout(x, drop.h, drop.seed, tp.embed_tag)
    return x, (None if bw.vit else pol.act(x))


def _embed_bwd(tp, dx: Tensor, grads: Grads) -> None:
    """Adjoint of the embeddings (reference models/vanilla_vit.py:242-253, models/vanilla_bert.py:307-325)."""
    pol, cfg, bw = tp.pol, tp.cfg, tp.bw
    vit = bw.vit
    T, B = tp.T, tp.B
    H, eps = cfg.hidden_size, cfg.layer_norm_eps
    if tp.embed_tag:
        dx = ops.dropout(dx, tp.drop.h, tp.drop.seed, tp.embed_tag)
    if vit:
        dpos = torch.zeros((T, H), dtype=torch.float32, device=dx.device)
        dcls = torch.zeros((H,), dtype=torch.float32, device=dx.device)
        dpatch = ops.vit_embed_bwd(dx, B, T, H, dpos, dcls, pol.act_dtype)
        e = "vit.embeddings."
        grads[e + "position_embeddings"] = dpos.reshape(1, T, H)
        grads[e + "cls_token"] = dcls.reshape(1, 1, H)
        dwp = _wgrad(pol, dpatch, tp.patches)
        P = cfg.img_patch_size
        grads[e + "patch_embeddings.projection.weight"] = dwp.reshape(H, cfg.img_channels, P, P)
        grads[e + "patch_embeddings.projection.bias"] = _bias_grad(dpatch)
    else:
        e = "bert.embeddings."
        pre = ops.bert_embed_sum(tp.xs, bw.word, bw.pos, bw.type0)
        dsum = _ln_bwd(grads, e + "LayerNorm", pre, dx, bw.emb_ln[0], eps, None)
        dword = torch.zeros_like(bw.word)
        dpos = torch.zeros_like(bw.pos)
        dtype0 = torch.zeros((H,), dtype=torch.float32, device=dx.device)
        ops.bert_embed_scatter(tp.xs, dsum, cfg.pad_token_id, dword, dpos, dtype0)
        dtt = torch.zeros((cfg.type_vocab_size, H), dtype=torch.float32, device=dx.device)
        dtt[0] = dtype0
        grads[e + "word_embeddings.weight"] = dword
        grads[e + "position_embeddings.weight"] = dpos
        grads[e + "token_type_embeddings.weight"] = dtt


# ------------------------------------------------------------------------------------------------
# whole explainer
# ------------------------------------------------------------------------------------------------
class _Tape:
    pass


def forward_train(sd: Dict[str, Tensor], cfg, precision: str, xs: Tensor, masks: Tensor, grand, null,
                  train_backbone: bool = True, dropout: bool = False) -> Tuple[Tensor, _Tape]:
    """train_backbone=False (Froyo, reference models/froyo_vit.py:88-97 / froyo_bert.py:92-101: every `vit.` / `bert.`
    parameter frozen): the encoder stack runs on the inference engine without a tape and the adjoint stops at the
    first explainer_attn block."""
    pol = _Policy(precision)
    tp = _Tape()
    tp.pol, tp.cfg, tp.masks, tp.xs = pol, cfg, masks, xs
    tp.train_backbone = train_backbone
    # dropout=True: hidden / attention-probability dropout as in the reference's train() mode (the frozen backbone of the
    # Froyo / LTT variants always runs deterministically on the inference engine)
    drop = _Drop(cfg, True) if dropout else None
    tp.drop, tp.embed_tag, tp.head_tag = drop, 0, 0
    bw = engine.BackboneWeights(sd, cfg, pol)
    vit = bw.vit
    T = n_players_of(cfg) + 1
    H, heads, eps = cfg.hidden_size, cfg.num_attention_heads, cfg.layer_norm_eps
    B = xs.shape[0]
    tp.bw, tp.T, tp.B = bw, T, B
    root = "vit" if vit else "bert"
    tp.layers = []
    if train_backbone:
        x, xa = _embed_fwd(tp, bw, cfg, pol, xs, drop)
        for i, lw in enumerate(bw.layers):
            prefix = f"{root}.encoder.layers.{i}"
            if vit:
                x, t = vit_layer_fwd(pol, lw, x, masks, T, heads, eps, drop)
            else:
                x, xa, t = bert_layer_fwd(pol, lw, x, xa, masks, T, heads, eps, drop)
            tp.layers.append((prefix, lw, t))
    else:
        x, xa = engine.run_backbone(bw, cfg, pol, xs, masks, 1)
        if not vit and xa is None:
            xa = pol.act(x)
    if vit:
        tp.x_pre_final = x if train_backbone else None
        _, x = ops.layernorm(x, bw.final_ln[0], bw.final_ln[1], eps, want_bf16=False, want_f32=True)
    # CLS row of the backbone output (ViT: after vit.layernorm; BERT: last block's output): the Duo variants hang a
    # classification head on it (reference models/duo_vanilla_vit.py:104-109, duo_vanilla_bert.py:120-125)
    tp.x_cls = x.reshape(B, T, H)[:, 0, :].clone()
    for i in range(cfg.explainer_attn_num_layers):
        prefix = f"explainer_attn.{i}"
        lw = LayerWeights(sd, prefix, pol, vit)
        if vit:
            x, t = vit_layer_fwd(pol, lw, x, masks, T, heads, eps, drop)
        else:
            x, xa, t = bert_layer_fwd(pol, lw, x, xa, masks, T, heads, eps, drop)
        tp.layers.append((prefix, lw, t))
    if vit:
        tp.mlp_ln = (_f32(sd["explainer_mlp.0.weight"]), _f32(sd["explainer_mlp.0.bias"]))
        tp.names = ("explainer_mlp.1", "explainer_mlp.3", "explainer_mlp.5")
        tp.x_last = x
        h0 = pol.ln(x, tp.mlp_ln[0], tp.mlp_ln[1], 1e-5)[0]
    else:
        tp.names = ("explainer_mlp.0", "explainer_mlp.2", "explainer_mlp.4")
        if drop is not None and drop.h:      # explainer_dropout (reference models/vanilla_bert.py:152)
            tp.head_tag = drop.tag()
            h0 = ops.dropout(x, drop.h, drop.seed, tp.head_tag, out_dtype=pol.act_dtype)
        else:
            h0 = xa
    na, nb, nc = tp.names
    tp.w_a, tp.w_b = pol.weight(sd[na + ".weight"]), pol.weight(sd[nb + ".weight"])
    tp.w_c, b_c = _f32(sd[nc + ".weight"]), _f32(sd[nc + ".bias"])
    tp.h0 = h0
    tp.za, tp.ha = _linear_gelu(pol, h0, tp.w_a, _f32(sd[na + ".bias"]))
    tp.zb, tp.hb = _linear_gelu(pol, tp.ha, tp.w_b, _f32(sd[nb + ".bias"]))
    phi = ops.explainer_head_fwd(tp.hb, B, T, tp.w_c, b_c, grand, null, bool(cfg.explainer_normalize))
    return phi, tp


def backward_train(tp: _Tape, dphi: Tensor, dx_cls: Optional[Tensor] = None, sink=None) -> Grads:
    """dx_cls (B, H): gradient w.r.t. tape.x_cls coming from a head outside this node (Duo variants), or None.
    sink (dist.OverlappedGradReducer | None): gets the gradients of every stage as soon as they exist (head first, then the
    blocks from the last to the first), so that their all-reduce overlaps the adjoint of the remaining blocks."""
    pol, cfg, bw = tp.pol, tp.cfg, tp.bw
    vit = bw.vit
    T, B = tp.T, tp.B
    H, heads, eps = cfg.hidden_size, cfg.num_attention_heads, cfg.layer_norm_eps
    grads: Grads = {}
    na, nb, nc = tp.names
    dWc, dbc = torch.zeros_like(tp.w_c), torch.zeros((tp.w_c.shape[0],), dtype=torch.float32, device=dphi.device)
    dhb = ops.explainer_head_bwd(dphi, tp.hb, B, T, tp.w_c, bool(cfg.explainer_normalize), dWc, dbc)
    grads[nc + ".weight"], grads[nc + ".bias"] = dWc, dbc
    dzb = ops.gelu_bwd(dhb, tp.zb)
    _linear_bwd(pol, grads, nb, dzb, tp.ha)
    dza = _dgrad_gelu(pol, dzb, tp.w_b, tp.za)
    _linear_bwd(pol, grads, na, dza, tp.h0)
    dx = _dgrad(pol, dza, tp.w_a, out_f32=True)
    if vit:
        dx = _ln_bwd(grads, "explainer_mlp.0", tp.x_last, dx, tp.mlp_ln[0], 1e-5, None)
    elif tp.head_tag:
        dx = ops.dropout(dx, tp.drop.h, tp.drop.seed, tp.head_tag)
    handed = 0

    def hand_over():
        nonlocal handed
        if sink is not None and len(grads) > handed:
            names = list(grads.keys())[handed:]
            handed = len(grads)
            sink.push(grads, names)

    hand_over()
    n_backbone = len(bw.layers) if tp.train_backbone else 0
    for idx in range(len(tp.layers) - 1, -1, -1):
        prefix, lw, t = tp.layers[idx]
        if tp.train_backbone and idx == n_backbone - 1 and dx_cls is not None:
            dx = dx.clone()
            dx.view(B, T, H)[:, 0, :] += dx_cls          # the classification head's share, at the CLS rows
        if vit and tp.train_backbone and idx == n_backbone - 1:
            # crossing from explainer_attn back into the backbone: adjoint of vit.layernorm
            dx = _ln_bwd(grads, "vit.layernorm", tp.x_pre_final, dx, bw.final_ln[0], eps, None)
        need_dx = tp.train_backbone or idx > 0
        if vit:
            dx = vit_layer_bwd(pol, lw, prefix, t, dx, tp.masks, T, heads, eps, grads, need_dx)
        else:
            dx = bert_layer_bwd(pol, lw, prefix, t, dx, tp.masks, T, heads, eps, grads, need_dx)
        hand_over()
    if tp.train_backbone:
        _embed_bwd(tp, dx, grads)
    hand_over()
    return grads


class _ExplainerTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xs, masks, grand, null, cfg, precision, dropout, names, sink, *params):
        sd = {n: p.detach() for n, p in zip(names, params)}
        train_backbone = any(ctx.needs_input_grad[9 + i] for i, n in enumerate(names) if n.startswith(("vit.", "bert.")))
        with torch.no_grad():
            phi, tape = forward_train(sd, cfg, precision, xs, masks, grand, null, train_backbone, dropout)
        ctx.tape, ctx.names, ctx.sink = tape, names, sink
        ctx.shapes = [p.shape for p in params]
        ctx.set_materialize_grads(False)
        return phi, tape.x_cls

    @staticmethod
    def backward(ctx, dphi, dx_cls):
        with torch.no_grad(), _arena_scope():
            if dphi is None:      # only the class output was used downstream
                dphi = torch.zeros((ctx.tape.B, ctx.tape.cfg.num_labels, ctx.tape.T - 1), dtype=torch.float32,
                                   device=ctx.tape.hb.device)
            grads = backward_train(ctx.tape, dphi.contiguous().float(),
                                   dx_cls.contiguous().float() if dx_cls is not None else None, sink=ctx.sink)
            if ctx.sink is not None:
                grads = ctx.sink.finish(grads)          # averaged over the data-parallel ranks, views of the flat buckets
        ctx.tape = None
        out = []
        for n, shp in zip(ctx.names, ctx.shapes):
            g = grads.get(n)
            out.append(g.reshape(shp) if g is not None else None)
        return (None, None, None, None, None, None, None, None, None, *out)


def _wants_dropout(model) -> bool:
    """train() mode turns the reference's nn.Dropout modules on; `module.agb_dropout = False` keeps this path
    deterministic (p = 0), e.g. for gradient parity against the reference run in eval() mode."""
    cfg = model.config
    return bool(model.training and getattr(model, "agb_dropout", True)
                and (cfg.hidden_dropout_prob > 0 or cfg.attention_probs_dropout_prob > 0))


def explainer_forward_train(model, xs: Tensor, words: Tensor, grand: Optional[Tensor], null: Optional[Tensor]) -> Tensor:
    """Differentiable explainer forward for `VanillaViTExplainer` / `VanillaBertExplainer` (w.r.t. parameters)."""
    named = list(model.named_parameters())
    names = [n for n, _ in named]
    params = [p for _, p in named]
    if params[0].device.type != "cuda":
        raise RuntimeError("autognothi_b200 models run on CUDA only (no CPU fallback)")
    g = grand.detach() if grand is not None else None
    nl = null.detach() if null is not None else None
    phi, _x_cls = _ExplainerTrainFn.apply(xs, words, g, nl, model.config, model.agb_precision, _wants_dropout(model), names,
                                          getattr(model, "agb_grad_reducer", None), *params)
    return phi


def duo_explainer_forward_train(model, xs: Tensor, words: Tensor, grand: Optional[Tensor], null: Optional[Tensor]
                                ) -> Tuple[Tensor, Tensor]:
    """Differentiable forward of the Duo explainers (reference models/duo_vanilla_vit.py:96-123, duo_vanilla_bert.py:110-148;
    trained with cross_entropy(class output) + Shapley loss, scripts/train_duo_explainer.py:180-196).
    -> (phi, class output): ViT softmax probabilities, BERT raw logits — as the reference returns them.  The backbone +
    explainer tail is one autograd node with two outputs (phi and the CLS row of the backbone output); the small
    classification head stays in torch autograd and its gradient re-enters the node at the CLS rows."""
    named_all = dict(model.named_parameters())
    head = ("classifier.", "bert_pooler.")
    named = [(n, p) for n, p in named_all.items() if not n.startswith(head)]
    names = [n for n, _ in named]
    if named[0][1].device.type != "cuda":
        raise RuntimeError("autognothi_b200 models run on CUDA only (no CPU fallback)")
    cfg = model.config
    g = grand.detach() if grand is not None else None
    nl = null.detach() if null is not None else None
    dropout = _wants_dropout(model)
    phi, x_cls = _ExplainerTrainFn.apply(xs, words, g, nl, cfg, model.agb_precision, dropout, names, None, *[p for _, p in named])
    if hasattr(cfg, "img_px_size"):
        logits = torch.nn.functional.linear(x_cls, named_all["classifier.weight"], named_all["classifier.bias"])
        return phi, torch.softmax(logits, dim=-1)
    h = torch.tanh(torch.nn.functional.linear(x_cls, named_all["bert_pooler.dense.weight"], named_all["bert_pooler.dense.bias"]))
    h = torch.nn.functional.dropout(h, float(cfg.hidden_dropout_prob), training=dropout)
    return phi, torch.nn.functional.linear(h, named_all["classifier.weight"], named_all["classifier.bias"])


# ------------------------------------------------------------------------------------------------
# surrogate training (reference scripts/train_surrogate.py:131-150): the masked backbone is ONE
# autograd node returning the CLS rows of the last hidden state; the tiny head (final LayerNorm / pooler, Linear,
# Softmax on B rows) and loss_logits_kl_divergence stay in torch autograd.
# ------------------------------------------------------------------------------------------------
def backbone_forward_train(sd: Dict[str, Tensor], cfg, precision: str, xs: Tensor, masks: Tensor, dropout: bool = False
                           ) -> Tuple[Tensor, _Tape]:
    """-> (x_cls (B, H) fp32: ViT = last block output BEFORE vit.layernorm, BERT = last block output; tape)"""
    pol = _Policy(precision)
    tp = _Tape()
    tp.pol, tp.cfg, tp.masks, tp.xs = pol, cfg, masks, xs
    bw = engine.BackboneWeights(sd, cfg, pol)
    T = n_players_of(cfg) + 1
    heads, eps = cfg.num_attention_heads, cfg.layer_norm_eps
    tp.bw, tp.T, tp.B = bw, T, xs.shape[0]
    assert masks.shape[0] == tp.B, "surrogate training takes one mask row per input"
    root = "vit" if bw.vit else "bert"
    drop = _Drop(cfg, True) if dropout else None
    x, xa = _embed_fwd(tp, bw, cfg, pol, xs, drop)
    tp.layers = []
    for i, lw in enumerate(bw.layers):
        if bw.vit:
            x, t = vit_layer_fwd(pol, lw, x, masks, T, heads, eps, drop)
        else:
            x, xa, t = bert_layer_fwd(pol, lw, x, xa, masks, T, heads, eps, drop)
        tp.layers.append((f"{root}.encoder.layers.{i}", lw, t))
    return x.reshape(tp.B, T, -1)[:, 0, :].contiguous(), tp


def backbone_backward_train(tp: _Tape, dx_cls: Tensor) -> Grads:
    pol, cfg, bw = tp.pol, tp.cfg, tp.bw
    T, B = tp.T, tp.B
    H, heads, eps = cfg.hidden_size, cfg.num_attention_heads, cfg.layer_norm_eps
    grads: Grads = {}
    dx = torch.zeros((B, T, H), dtype=torch.float32, device=dx_cls.device)
    dx[:, 0, :] = dx_cls
    dx = dx.reshape(B * T, H)
    for prefix, lw, t in reversed(tp.layers):
        if bw.vit:
            dx = vit_layer_bwd(pol, lw, prefix, t, dx, tp.masks, T, heads, eps, grads)
        else:
            dx = bert_layer_bwd(pol, lw, prefix, t, dx, tp.masks, T, heads, eps, grads)
    _embed_bwd(tp, dx, grads)
    return grads


class _BackboneTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xs, masks, cfg, precision, dropout, names, *params):
        sd = {n: p.detach() for n, p in zip(names, params)}
        with torch.no_grad():
            x_cls, tape = backbone_forward_train(sd, cfg, precision, xs, masks, dropout)
        ctx.tape, ctx.names = tape, names
        ctx.shapes = [p.shape for p in params]
        return x_cls

    @staticmethod
    def backward(ctx, dx_cls):
        with torch.no_grad(), _arena_scope():
            grads = backbone_backward_train(ctx.tape, dx_cls.contiguous().float())
        ctx.tape = None
        out = []
        for n, shp in zip(ctx.names, ctx.shapes):
            g = grads.get(n)
            out.append(g.reshape(shp) if g is not None else None)
        return (None, None, None, None, None, None, *out)


def surrogate_forward_train(model, xs: Tensor, words: Tensor) -> Tensor:
    """Differentiable (w.r.t. parameters) masked surrogate / classifier forward -> (B, num_labels) probabilities.
    reference models/vanilla_vit.py:51-56 and models/vanilla_bert.py:61-77 in train() mode, dropout p = 0."""
    named = dict(model.named_parameters())
    if next(iter(named.values())).device.type != "cuda":
        raise RuntimeError("autognothi_b200 models run on CUDA only (no CPU fallback)")
    cfg = model.config
    vit = hasattr(cfg, "img_px_size")
    root = "vit." if vit else "bert."
    names = [n for n in named if n.startswith(root)]   # vit.layernorm.* ride along unused (their grads come from the torch head)
    dropout = _wants_dropout(model)
    if any(named[n].requires_grad for n in names):
        x_cls = _BackboneTrainFn.apply(xs, words, cfg, model.agb_precision, dropout, names, *[named[n] for n in names])
    else:
        # frozen backbone (Froyo surrogate, reference models/froyo_vit.py:76-85): inference engine, CLS row only, no tape
        with torch.no_grad():
            pol = _Policy(model.agb_precision)
            bw = engine.BackboneWeights({n: named[n].detach() for n in names}, cfg, pol)
            x_cls, _ = engine.run_backbone(bw, cfg, pol, xs, words, 1, cls_only=True)
            x_cls = x_cls.reshape(xs.shape[0], -1, cfg.hidden_size)[:, 0, :].contiguous()
    if vit:
        h = torch.nn.functional.layer_norm(x_cls, (cfg.hidden_size,), named["vit.layernorm.weight"],
                                           named["vit.layernorm.bias"], cfg.layer_norm_eps)
    else:
        h = torch.tanh(torch.nn.functional.linear(x_cls, named["bert_pooler.dense.weight"], named["bert_pooler.dense.bias"]))
        h = torch.nn.functional.dropout(h, float(cfg.hidden_dropout_prob), training=dropout)   # reference vanilla_bert.py:74
    logits = torch.nn.functional.linear(h, named["classifier.weight"], named["classifier.bias"])
    return torch.softmax(logits, dim=-1)


# ------------------------------------------------------------------------------------------------
# LTT side-ladder training (reference models/ltt_vit.py:407-440, models/ltt_bert.py:467-499 trained by
# scripts/train_explainer.py / train_surrogate.py with the backbone frozen, models/ltt_vit.py:68-74).  The frozen
# backbone runs on the inference engine (no tape); each rung keeps the backbone activation it tapped (for the map's
# weight gradient), its pre-GELU map output and the tape of its narrow block.  No gradient flows into the backbone.
# ------------------------------------------------------------------------------------------------
def _ltt_side_forward(tp: _Tape, sd, cfg, pol, xs: Tensor, masks: Tensor, freeze_layer, drop=None
                      ) -> Tuple[Tensor, Optional[Tensor], Tensor]:
    """-> (side state (B*T, Hs) fp32, its activation copy (BERT) | None, backbone class probabilities (B, C))"""
    bw = engine.BackboneWeights(sd, cfg, pol)
    br = engine.SideBranch(sd, cfg, pol, bw.vit, 0)
    T = n_players_of(cfg) + 1
    heads, eps = cfg.num_attention_heads, cfg.layer_norm_eps
    B = xs.shape[0]
    L = len(bw.layers)
    stop = L if freeze_layer is None else max(1, min(L, int(freeze_layer)))
    tp.bw, tp.br, tp.T, tp.B, tp.rungs = bw, br, T, B, []
    state: Dict[str, Optional[Tensor]] = {"s": None, "sa": None}

    def hook(i: int, x: Tensor, x_act: Optional[Tensor], _masks=None) -> None:
        if i >= stop:
            return
        if x_act is None:
            x_act = pol.act(x) if pol.bf16 else x.clone()     # fp32 mode: x is updated in place by the next block
        w, b = br.maps[i]
        z = pol.linear(x_act, w, b, out_f32=True)
        s_in = ops.gelu_fwd(z)
        if state["s"] is not None:
            s_in.add_(state["s"])
        if bw.vit:
            s_out, t = vit_layer_fwd(pol, br.layers[i], s_in, masks, T, heads, eps, drop)
            sa = None
        else:
            s_out, sa, t = bert_layer_fwd(pol, br.layers[i], s_in, pol.act(s_in), masks, T, heads, eps, drop)
        tp.rungs.append((i, x_act, z, br.layers[i], t))
        state["s"], state["sa"] = s_out, sa

    x, _ = engine.run_backbone(bw, cfg, pol, xs, masks, 1, layer_hook=hook)
    x3 = x.reshape(B, -1, cfg.hidden_size)
    w_cls, b_cls = _f32(sd["classifier.weight"]), _f32(sd["classifier.bias"])
    if bw.vit:
        cls = ops.cls_head(x3, 0, w_cls, b_cls, ln=(bw.final_ln[0], bw.final_ln[1], eps))
    else:
        cls = ops.cls_head(x3, 1, w_cls, b_cls, pool=(_f32(sd["bert_pooler.dense.weight"]), _f32(sd["bert_pooler.dense.bias"])))
    return state["s"], state["sa"], cls


def _ltt_rungs_backward(tp: _Tape, dx: Tensor, grads: Grads) -> None:
    pol, cfg, bw = tp.pol, tp.cfg, tp.bw
    T, heads, eps = tp.T, cfg.num_attention_heads, cfg.layer_norm_eps
    root = "vit" if bw.vit else "bert"
    for i, x_act, z, lw, t in reversed(tp.rungs):
        prefix = f"{root}.encoder.s_attn_layers.0_{i}"
        if bw.vit:
            ds_in = vit_layer_bwd(pol, lw, prefix, t, dx, tp.masks, T, heads, eps, grads)
        else:
            ds_in = bert_layer_bwd(pol, lw, prefix, t, dx, tp.masks, T, heads, eps, grads)
        dz = ops.gelu_bwd(ds_in, z)
        _linear_bwd(pol, grads, f"{root}.encoder.s_attn_maps.0_{i}", pol.act(dz), x_act)
        dx = ds_in                                   # s_in = s_prev + GELU(map): the residual passes straight through


def ltt_forward_train(sd, cfg, precision: str, xs: Tensor, masks: Tensor, grand, null, kind: str, freeze_layer,
                      dropout: bool = False):
    """kind "explainer": -> (phi (B, C, n), backbone probabilities, tape);
       kind "surrogate": -> (CLS rows of the side state (B, Hs) [ViT: before vit.s_attn_layernorm.0], backbone
                             probabilities, tape) — the tiny side head stays in torch autograd."""
    pol = _Policy(precision)
    tp = _Tape()
    tp.pol, tp.cfg, tp.masks, tp.xs, tp.kind = pol, cfg, masks, xs, kind
    drop = _Drop(cfg, True) if dropout else None      # the side ladder's dropout sites; the frozen backbone has none here
    tp.drop, tp.head_tag = drop, 0
    s, sa, cls = _ltt_side_forward(tp, sd, cfg, pol, xs, masks, freeze_layer, drop)
    vit = tp.bw.vit
    T, B = tp.T, tp.B
    heads, eps = cfg.num_attention_heads, cfg.layer_norm_eps
    Hs = cfg.s_attn_hidden_size
    if kind == "surrogate":
        return s.reshape(B, T, Hs)[:, 0, :].contiguous(), cls, tp
    tp.layers = []
    if vit:
        tp.s_pre_ln = s
        _, s = ops.layernorm(s, tp.br.final_ln[0], tp.br.final_ln[1], eps, want_bf16=False, want_f32=True)
    attn = "s_explainer_attn" if vit else "s_attn_attention_layers"
    mlp = "s_explainer_mlp" if vit else "s_attn_explainer"
    for i in range(cfg.explainer_s_attn_num_layers):
        prefix = f"{attn}.{i}"
        lw = LayerWeights(sd, prefix, pol, vit)
        if vit:
            s, t = vit_layer_fwd(pol, lw, s, masks, T, heads, eps, drop)
        else:
            s, sa, t = bert_layer_fwd(pol, lw, s, sa, masks, T, heads, eps, drop)
        tp.layers.append((prefix, lw, t))
    if vit:
        tp.mlp_ln_name = mlp + ".0"
        tp.mlp_ln = (_f32(sd[mlp + ".0.weight"]), _f32(sd[mlp + ".0.bias"]))
        tp.names = (mlp + ".1", mlp + ".3", mlp + ".5")
        tp.x_last = s
        h0 = pol.ln(s, tp.mlp_ln[0], tp.mlp_ln[1], 1e-5)[0]
    else:
        tp.names = (mlp + ".0", mlp + ".2", mlp + ".4")
        if drop is not None and drop.h:      # s_attn_exp_dropout (reference models/ltt_bert.py:203)
            tp.head_tag = drop.tag()
            h0 = ops.dropout(s, drop.h, drop.seed, tp.head_tag, out_dtype=pol.act_dtype)
        else:
            h0 = sa
    na, nb, nc = tp.names
    tp.w_a, tp.w_b = pol.weight(sd[na + ".weight"]), pol.weight(sd[nb + ".weight"])
    tp.w_c, b_c = _f32(sd[nc + ".weight"]), _f32(sd[nc + ".bias"])
    tp.h0 = h0
    tp.za, tp.ha = _linear_gelu(pol, h0, tp.w_a, _f32(sd[na + ".bias"]))
    tp.zb, tp.hb = _linear_gelu(pol, tp.ha, tp.w_b, _f32(sd[nb + ".bias"]))
    phi = ops.explainer_head_fwd(tp.hb, B, T, tp.w_c, b_c, grand, null, bool(cfg.explainer_normalize))
    return phi, cls, tp


def ltt_backward_train(tp: _Tape, dout: Tensor) -> Grads:
    pol, cfg, bw = tp.pol, tp.cfg, tp.bw
    vit = bw.vit
    T, B = tp.T, tp.B
    heads, eps = cfg.num_attention_heads, cfg.layer_norm_eps
    Hs = cfg.s_attn_hidden_size
    grads: Grads = {}
    if tp.kind == "surrogate":
        dx = torch.zeros((B, T, Hs), dtype=torch.float32, device=dout.device)
        dx[:, 0, :] = dout
        _ltt_rungs_backward(tp, dx.reshape(B * T, Hs), grads)
        return grads
    na, nb, nc = tp.names
    dWc, dbc = torch.zeros_like(tp.w_c), torch.zeros((tp.w_c.shape[0],), dtype=torch.float32, device=dout.device)
    dhb = ops.explainer_head_bwd(dout, tp.hb, B, T, tp.w_c, bool(cfg.explainer_normalize), dWc, dbc)
    grads[nc + ".weight"], grads[nc + ".bias"] = dWc, dbc
    dzb = ops.gelu_bwd(dhb, tp.zb)
    _linear_bwd(pol, grads, nb, dzb, tp.ha)
    dza = _dgrad_gelu(pol, dzb, tp.w_b, tp.za)
    _linear_bwd(pol, grads, na, dza, tp.h0)
    dx = _dgrad(pol, dza, tp.w_a, out_f32=True)
    if vit:
        dx = _ln_bwd(grads, tp.mlp_ln_name, tp.x_last, dx, tp.mlp_ln[0], 1e-5, None)
    elif tp.head_tag:
        dx = ops.dropout(dx, tp.drop.h, tp.drop.seed, tp.head_tag)
    for prefix, lw, t in reversed(tp.layers):
        if vit:
            dx = vit_layer_bwd(pol, lw, prefix, t, dx, tp.masks, T, heads, eps, grads)
        else:
            dx = bert_layer_bwd(pol, lw, prefix, t, dx, tp.masks, T, heads, eps, grads)
    if vit:
        dx = _ln_bwd(grads, "vit.s_attn_layernorm.0", tp.s_pre_ln, dx, tp.br.final_ln[0], eps, None)
    _ltt_rungs_backward(tp, dx, grads)
    return grads


class _LttTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xs, masks, grand, null, cfg, precision, kind, freeze_layer, dropout, names, *params):
        sd = {n: p.detach() for n, p in zip(names, params)}
        with torch.no_grad():
            out, cls, tape = ltt_forward_train(sd, cfg, precision, xs, masks, grand, null, kind, freeze_layer, dropout)
        ctx.tape, ctx.names = tape, names
        ctx.shapes = [p.shape for p in params]
        ctx.mark_non_differentiable(cls)
        return out, cls

    @staticmethod
    def backward(ctx, dout, _dcls):
        with torch.no_grad(), _arena_scope():
            grads = ltt_backward_train(ctx.tape, dout.contiguous().float())
        ctx.tape = None
        out = []
        for i, (n, shp) in enumerate(zip(ctx.names, ctx.shapes)):
            g = grads.get(n) if ctx.needs_input_grad[10 + i] else None
            out.append(g.reshape(shp) if g is not None else None)
        return (None,) * 10 + tuple(out)


def _ltt_apply(model, xs, words, grand, null, kind):
    named = list(model.named_parameters())
    if named[0][1].device.type != "cuda":
        raise RuntimeError("autognothi_b200 models run on CUDA only (no CPU fallback)")
    names = [n for n, _ in named]
    g = grand.detach() if grand is not None else None
    nl = null.detach() if null is not None else None
    return _LttTrainFn.apply(xs, words, g, nl, model.config, model.agb_precision, kind,
                             getattr(model, "_ltt_freeze_layer", None), _wants_dropout(model), names, *[p for _, p in named])


def ltt_explainer_forward_train(model, xs: Tensor, words: Tensor, grand: Optional[Tensor], null: Optional[Tensor]
                                ) -> Tuple[Tensor, Tensor]:
    """Differentiable (w.r.t. the side ladder and the side explainer) forward of LttViTExplainer / LttBertExplainer
    -> (phi, backbone probabilities).  The backbone (`*.embeddings`, `*.encoder.layers`, ViT `vit.layernorm`, poolers,
    `classifier`) never receives gradients on this path: the reference freezes it in train() (models/ltt_vit.py:132-138)."""
    return _ltt_apply(model, xs, words, grand, null, "explainer")


def ltt_surrogate_forward_train(model, xs: Tensor, words: Tensor) -> Tuple[Tensor, Tensor]:
    """Differentiable forward of LttViTSurrogate / LttBertSurrogate -> (side-ladder probabilities, backbone probabilities).
    The ladder is one autograd node returning the CLS rows of the side state; the side head (ViT: vit.s_attn_layernorm.0
    + s_attn_classifier; BERT: bert_s_attn_pooler + s_attn_classifier) and the softmax stay in torch autograd."""
    s_cls, cls = _ltt_apply(model, xs, words, None, None, "surrogate")
    named = dict(model.named_parameters())
    cfg = model.config
    if hasattr(cfg, "img_px_size"):
        h = torch.nn.functional.layer_norm(s_cls, (cfg.s_attn_hidden_size,), named["vit.s_attn_layernorm.0.weight"],
                                           named["vit.s_attn_layernorm.0.bias"], cfg.layer_norm_eps)
    else:
        h = torch.tanh(torch.nn.functional.linear(s_cls, named["bert_s_attn_pooler.dense.weight"],
                                                  named["bert_s_attn_pooler.dense.bias"]))
        h = torch.nn.functional.dropout(h, float(cfg.hidden_dropout_prob), training=_wants_dropout(model))  # ltt_bert.py:115
    logits = torch.nn.functional.linear(h, named["s_attn_classifier.weight"], named["s_attn_classifier.bias"])
    return torch.softmax(logits, dim=-1), cls
