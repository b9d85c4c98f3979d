"""Explainer training path: forward with a tape + hand-written backward over the CUDA kernels.

The reference trains the explainer with torch autograd (scripts/train_explainer.py:182-198 over
models/vanilla_vit.py:102-130 / models/vanilla_bert.py:123-162).  Here the whole explainer is ONE
autograd node: `explainer_forward_train` returns phi attached to the graph, and `loss.backward()` runs
the adjoint below — every Linear's dgrad/wgrad on the tcgen05 GEMM (MN-major operand modes, no
transposed copies), attention / LayerNorm / GELU / head adjoints on their own kernels.  So the
reference's loop (`phi = fw_explainer(...); loss = loss_shapley_new(...); loss.backward();
optimizer.step()`) runs unchanged.

Dropout: the reference keeps hidden/attention dropout (p = 0.1) active while training.  This path
implements p = 0 (deterministic); bench.py states that setting.  Parity tests compare gradients with the
reference in eval() mode, where the reference's dropout is the identity.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor

from . import engine, ops
from .engine import LayerWeights, _Policy, _f32, n_players_of

Grads = Dict[str, Tensor]


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


def _bias_grad(dy: Tensor) -> Tensor:
    g = torch.zeros((dy.shape[1],), dtype=torch.float32, device=dy.device)
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
    dg = torch.zeros_like(gamma)
    db = torch.zeros_like(gamma)
    dx = ops.layernorm_bwd(x, dy, gamma, eps, dres, dg, db)
    grads[name + ".weight"], grads[name + ".bias"] = dg, db
    return dx


# ------------------------------------------------------------------------------------------------
# ViT block (pre-LN), reference models/vanilla_vit.py:364-377
# ------------------------------------------------------------------------------------------------
def vit_layer_fwd(pol, lw: LayerWeights, x: Tensor, masks: Tensor, T: int, heads: int, eps: float):
    h1 = pol.ln(x, lw.ln1[0], lw.ln1[1], eps)[0] if lw.ln1 is not None else pol.act(x)
    qkv = pol.linear(h1, lw.wqkv, lw.bqkv)
    ctx = ops.masked_attention(qkv, masks, T, heads, ops.MASK_MUL0)
    x_mid = pol.linear(ctx, lw.wo, lw.bo, residual=x, out_f32=True)
    h2 = pol.ln(x_mid, lw.ln2[0], lw.ln2[1], eps)[0]
    z = pol.linear(h2, lw.w1, lw.b1)
    f = ops.gelu_fwd(z)
    x_out = pol.linear(f, lw.w2, lw.b2, residual=x_mid, out_f32=True)
    return x_out, dict(x_in=x, h1=h1, qkv=qkv, ctx=ctx, x_mid=x_mid, h2=h2, z=z, f=f)


def vit_layer_bwd(pol, lw: LayerWeights, prefix: str, t: dict, dx_out: Tensor, masks: Tensor, T: int, heads: int,
                  eps: float, grads: Grads) -> Tensor:
    H = dx_out.shape[1]
    g = pol.act(dx_out)
    _linear_bwd(pol, grads, prefix + ".output.dense", g, t["f"])
    df = _dgrad(pol, g, lw.w2, out_f32=False)
    dz = ops.gelu_bwd(df, t["z"])
    _linear_bwd(pol, grads, prefix + ".intermediate.dense", dz, t["h2"])
    dh2 = _dgrad(pol, dz, lw.w1, out_f32=True)
    dx_mid = _ln_bwd(grads, prefix + ".layernorm_after", t["x_mid"], dh2, lw.ln2[0], eps, dx_out)
    g = pol.act(dx_mid)
    _linear_bwd(pol, grads, prefix + ".attention.output.dense", g, t["ctx"])
    dctx = _dgrad(pol, g, lw.wo, out_f32=False)
    dqkv = ops.masked_attention_bwd(t["qkv"], dctx, masks, T, heads, ops.MASK_MUL0)
    _qkv_bwd(pol, grads, prefix, dqkv, t["h1"], H)
    if lw.ln1 is not None:
        dh1 = _dgrad(pol, dqkv, lw.wqkv, out_f32=True)
        return _ln_bwd(grads, prefix + ".layernorm_before", t["x_in"], dh1, lw.ln1[0], eps, dx_mid)
    return _dgrad(pol, dqkv, lw.wqkv, out_f32=True, residual=dx_mid)


# ------------------------------------------------------------------------------------------------
# BERT block (post-LN), reference models/vanilla_bert.py:396-427, 556-560, 600-604
# ------------------------------------------------------------------------------------------------
def bert_layer_fwd(pol, lw: LayerWeights, x: Tensor, xa: Tensor, masks: Tensor, T: int, heads: int, eps: float):
    qkv = pol.linear(xa, lw.wqkv, lw.bqkv)
    ctx = ops.masked_attention(qkv, masks, T, heads, ops.MASK_NEGINF)
    a_pre = pol.linear(ctx, lw.wo, lw.bo, residual=x, out_f32=True)
    if lw.ln1 is not None:
        aa, a = pol.ln(a_pre, lw.ln1[0], lw.ln1[1], eps, want_f32=True)
    else:
        a, aa = a_pre, pol.act(a_pre)
    z = pol.linear(aa, lw.w1, lw.b1)
    f = ops.gelu_fwd(z)
    y_pre = pol.linear(f, lw.w2, lw.b2, residual=a, out_f32=True)
    ya, y = pol.ln(y_pre, lw.ln2[0], lw.ln2[1], eps, want_f32=True)
    return y, ya, dict(xa=xa, qkv=qkv, ctx=ctx, a_pre=a_pre, aa=aa, z=z, f=f, y_pre=y_pre)


def bert_layer_bwd(pol, lw: LayerWeights, prefix: str, t: dict, dy: Tensor, masks: Tensor, T: int, heads: int,
                   eps: float, grads: Grads) -> Tensor:
    H = dy.shape[1]
    d_ypre = _ln_bwd(grads, prefix + ".output.LayerNorm", t["y_pre"], dy, lw.ln2[0], eps, None)
    g = pol.act(d_ypre)
    _linear_bwd(pol, grads, prefix + ".output.dense", g, t["f"])
    df = _dgrad(pol, g, lw.w2, out_f32=False)
    dz = ops.gelu_bwd(df, t["z"])
    _linear_bwd(pol, grads, prefix + ".intermediate.dense", dz, t["aa"])
    da = _dgrad(pol, dz, lw.w1, out_f32=True, residual=d_ypre)
    d_apre = _ln_bwd(grads, prefix + ".attention.output.LayerNorm", t["a_pre"], da, lw.ln1[0], eps, None) \
        if lw.ln1 is not None else da
    g = pol.act(d_apre)
    _linear_bwd(pol, grads, prefix + ".attention.output.dense", g, t["ctx"])
    dctx = _dgrad(pol, g, lw.wo, out_f32=False)
    dqkv = ops.masked_attention_bwd(t["qkv"], dctx, masks, T, heads, ops.MASK_NEGINF)
    _qkv_bwd(pol, grads, prefix, dqkv, t["xa"], H)
    return _dgrad(pol, dqkv, lw.wqkv, out_f32=True, residual=d_apre)


def _embed_fwd(tp, bw, cfg, pol, xs: Tensor):
    """Embeddings with S = 1 (one mask row per input); keeps what the adjoint needs on the tape."""
    T, H, eps = tp.T, cfg.hidden_size, cfg.layer_norm_eps
    B = xs.shape[0]
    if bw.vit:
        tp.patches = ops.vit_im2col(xs.float(), cfg.img_patch_size, pol.act_dtype)
        pe = pol.linear(tp.patches, bw.w_patch, bw.b_patch, out_f32=True)
        return ops.vit_assemble(pe, bw.cls_token, bw.pos_emb, B, 1, T, H).reshape(B * T, H), None
    x = ops.bert_embed(xs, bw.word, bw.pos, bw.type0, bw.emb_ln[0], bw.emb_ln[1], eps, 1).reshape(B * T, H)
    return x, pol.act(x)


def _embed_bwd(tp, dx: Tensor, grads: Grads) -> None:
    """Adjoint of the embeddings (reference models/vanilla_vit.py:242-253, models/vanilla_bert.py:307-325)."""
    pol, cfg, bw = tp.pol, tp.cfg, tp.bw
    vit = bw.vit
    T, B = tp.T, tp.B
    H, eps = cfg.hidden_size, cfg.layer_norm_eps
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


def forward_train(sd: Dict[str, Tensor], cfg, precision: str, xs: Tensor, masks: Tensor, grand, null) -> Tuple[Tensor, _Tape]:
    pol = _Policy(precision)
    tp = _Tape()
    tp.pol, tp.cfg, tp.masks, tp.xs = pol, cfg, masks, xs
    bw = engine.BackboneWeights(sd, cfg, pol)
    vit = bw.vit
    T = n_players_of(cfg) + 1
    H, heads, eps = cfg.hidden_size, cfg.num_attention_heads, cfg.layer_norm_eps
    B = xs.shape[0]
    tp.bw, tp.T, tp.B = bw, T, B
    root = "vit" if vit else "bert"
    x, xa = _embed_fwd(tp, bw, cfg, pol, xs)
    tp.layers = []
    for i, lw in enumerate(bw.layers):
        prefix = f"{root}.encoder.layers.{i}"
        if vit:
            x, t = vit_layer_fwd(pol, lw, x, masks, T, heads, eps)
        else:
            x, xa, t = bert_layer_fwd(pol, lw, x, xa, masks, T, heads, eps)
        tp.layers.append((prefix, lw, t))
    if vit:
        tp.x_pre_final = x
        _, x = ops.layernorm(x, bw.final_ln[0], bw.final_ln[1], eps, want_bf16=False, want_f32=True)
    for i in range(cfg.explainer_attn_num_layers):
        prefix = f"explainer_attn.{i}"
        lw = LayerWeights(sd, prefix, pol, vit)
        if vit:
            x, t = vit_layer_fwd(pol, lw, x, masks, T, heads, eps)
        else:
            x, xa, t = bert_layer_fwd(pol, lw, x, xa, masks, T, heads, eps)
        tp.layers.append((prefix, lw, t))
    if vit:
        tp.mlp_ln = (_f32(sd["explainer_mlp.0.weight"]), _f32(sd["explainer_mlp.0.bias"]))
        tp.names = ("explainer_mlp.1", "explainer_mlp.3", "explainer_mlp.5")
        tp.x_last = x
        h0 = pol.ln(x, tp.mlp_ln[0], tp.mlp_ln[1], 1e-5)[0]
    else:
        tp.names = ("explainer_mlp.0", "explainer_mlp.2", "explainer_mlp.4")
        h0 = xa
    na, nb, nc = tp.names
    tp.w_a, tp.w_b = pol.weight(sd[na + ".weight"]), pol.weight(sd[nb + ".weight"])
    tp.w_c, b_c = _f32(sd[nc + ".weight"]), _f32(sd[nc + ".bias"])
    tp.h0 = h0
    tp.za = pol.linear(h0, tp.w_a, _f32(sd[na + ".bias"]))
    tp.ha = ops.gelu_fwd(tp.za)
    tp.zb = pol.linear(tp.ha, tp.w_b, _f32(sd[nb + ".bias"]))
    tp.hb = ops.gelu_fwd(tp.zb)
    phi = ops.explainer_head_fwd(tp.hb, B, T, tp.w_c, b_c, grand, null, bool(cfg.explainer_normalize))
    return phi, tp


def backward_train(tp: _Tape, dphi: Tensor) -> Grads:
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
    dha = _dgrad(pol, dzb, tp.w_b, out_f32=False)
    dza = ops.gelu_bwd(dha, tp.za)
    _linear_bwd(pol, grads, na, dza, tp.h0)
    dx = _dgrad(pol, dza, tp.w_a, out_f32=True)
    if vit:
        dx = _ln_bwd(grads, "explainer_mlp.0", tp.x_last, dx, tp.mlp_ln[0], 1e-5, None)
    n_backbone = len(bw.layers)
    for idx in range(len(tp.layers) - 1, -1, -1):
        prefix, lw, t = tp.layers[idx]
        if vit and idx == n_backbone - 1:
            # crossing from explainer_attn back into the backbone: adjoint of vit.layernorm
            dx = _ln_bwd(grads, "vit.layernorm", tp.x_pre_final, dx, bw.final_ln[0], eps, None)
        if vit:
            dx = vit_layer_bwd(pol, lw, prefix, t, dx, tp.masks, T, heads, eps, grads)
        else:
            dx = bert_layer_bwd(pol, lw, prefix, t, dx, tp.masks, T, heads, eps, grads)
    _embed_bwd(tp, dx, grads)
    return grads


class _ExplainerTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xs, masks, grand, null, cfg, precision, names, *params):
        sd = {n: p.detach() for n, p in zip(names, params)}
        with torch.no_grad():
            phi, tape = forward_train(sd, cfg, precision, xs, masks, grand, null)
        ctx.tape, ctx.names = tape, names
        ctx.shapes = [p.shape for p in params]
        return phi

    @staticmethod
    def backward(ctx, dphi):
        with torch.no_grad():
            grads = backward_train(ctx.tape, dphi.contiguous().float())
        ctx.tape = None
        out = []
        for n, shp in zip(ctx.names, ctx.shapes):
            g = grads.get(n)
            out.append(g.reshape(shp) if g is not None else None)
        return (None, None, None, None, None, None, None, *out)


def explainer_forward_train(model, xs: Tensor, words: Tensor, grand: Optional[Tensor], null: Optional[Tensor]) -> Tensor:
    """Differentiable explainer forward for `VanillaViTExplainer` / `VanillaBertExplainer` (w.r.t. parameters)."""
    named = list(model.named_parameters())
    names = [n for n, _ in named]
    params = [p for _, p in named]
    if params[0].device.type != "cuda":
        raise RuntimeError("autognothi_b200 models run on CUDA only (no CPU fallback)")
    g = grand.detach() if grand is not None else None
    nl = null.detach() if null is not None else None
    return _ExplainerTrainFn.apply(xs, words, g, nl, model.config, model.agb_precision, names, *params)


# ------------------------------------------------------------------------------------------------
# surrogate training (reference scripts/train_surrogate.py:131-150): the masked backbone is ONE
# autograd node returning the CLS rows of the last hidden state; the tiny head (final LayerNorm / pooler, Linear,
# Softmax on B rows) and loss_logits_kl_divergence stay in torch autograd.
# ------------------------------------------------------------------------------------------------
def backbone_forward_train(sd: Dict[str, Tensor], cfg, precision: str, xs: Tensor, masks: Tensor) -> Tuple[Tensor, _Tape]:
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
    x, xa = _embed_fwd(tp, bw, cfg, pol, xs)
    tp.layers = []
    for i, lw in enumerate(bw.layers):
        if bw.vit:
            x, t = vit_layer_fwd(pol, lw, x, masks, T, heads, eps)
        else:
            x, xa, t = bert_layer_fwd(pol, lw, x, xa, masks, T, heads, eps)
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
    def forward(ctx, xs, masks, cfg, precision, names, *params):
        sd = {n: p.detach() for n, p in zip(names, params)}
        with torch.no_grad():
            x_cls, tape = backbone_forward_train(sd, cfg, precision, xs, masks)
        ctx.tape, ctx.names = tape, names
        ctx.shapes = [p.shape for p in params]
        return x_cls

    @staticmethod
    def backward(ctx, dx_cls):
        with torch.no_grad():
            grads = backbone_backward_train(ctx.tape, dx_cls.contiguous().float())
        ctx.tape = None
        out = []
        for n, shp in zip(ctx.names, ctx.shapes):
            g = grads.get(n)
            out.append(g.reshape(shp) if g is not None else None)
        return (None, None, None, None, None, *out)


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
    x_cls = _BackboneTrainFn.apply(xs, words, cfg, model.agb_precision, names, *[named[n] for n in names])
    if vit:
        h = torch.nn.functional.layer_norm(x_cls, (cfg.hidden_size,), named["vit.layernorm.weight"],
                                           named["vit.layernorm.bias"], cfg.layer_norm_eps)
    else:
        h = torch.tanh(torch.nn.functional.linear(x_cls, named["bert_pooler.dense.weight"], named["bert_pooler.dense.bias"]))
    logits = torch.nn.functional.linear(h, named["classifier.weight"], named["classifier.bias"])
    return torch.softmax(logits, dim=-1)
