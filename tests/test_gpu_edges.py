"""Edge cases of the drop-in boundary on the GPU: empty batches, ragged (input, coalition) shapes, row chunking, the
reference-shaped call vs the additive fast path.  The reference handles these through plain torch broadcasting; here they
cross hand-sized grids, so each is pinned explicitly."""
import numpy as np
import pytest
import torch

from oracle import configs as ocfg
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _np(t):
    return t.detach().float().cpu().numpy()


def _build(name, precision):
    from autognothi_b200.recipes.vanilla_bert import vanilla_bert_recipe
    from autognothi_b200.recipes.vanilla_vit import vanilla_vit_recipe
    cfgd = ocfg.get_config(name)
    rec = vanilla_vit_recipe() if ocfg.is_vit(cfgd) else vanilla_bert_recipe()
    cfg = rec.t_config(**cfgd)
    srg, exp = rec.t_surrogate(cfg), rec.t_explainer(cfg)
    srg.load_state_dict({k: torch.from_numpy(v) for k, v in synth.surrogate_state(cfgd, seed=0).items()}, strict=True)
    exp.load_state_dict({k: torch.from_numpy(v) for k, v in synth.explainer_state(cfgd, seed=1).items()}, strict=True)
    srg, exp = srg.to(DEV).eval(), exp.to(DEV).eval()
    srg.agb_precision = exp.agb_precision = precision
    return rec, cfgd, cfg, srg, exp


@pytest.mark.parametrize("name,precision", [("vit_mini", "bf16"), ("vit_mini", "fp32"), ("bert_mini", "bf16"), ("bert_mini", "fp32")])
def test_empty_batch(agb, name, precision):
    from autognothi_b200.models import shapley as ash
    rec, cfgd, cfg, srg, exp = _build(name, precision)
    n, C = rec.n_players(cfg), cfgd["num_labels"]
    xs = torch.from_numpy(synth.inputs(cfgd, 2, seed=0)).to(DEV)[:0]
    m0 = torch.zeros((0, n), dtype=torch.int64, device=DEV)
    with torch.no_grad():
        ys, _ = rec.fw_surrogate(srg, xs, m0)
        phi, _ = rec.fw_explainer(exp, xs, m0, torch.zeros((0, C), device=DEV), torch.zeros((1, C), device=DEV))
    assert ys.shape == (0, C) and phi.shape == (0, C, n)
    if name.startswith("vit"):       # the LTT / Froyo bundles take the same early exit
        from autognothi_b200.recipes.froyo_vit import froyo_vit_recipe
        from autognothi_b200.recipes.ltt_vit import ltt_vit_recipe
        for r, cname in ((froyo_vit_recipe(), "vit_mini"), (ltt_vit_recipe(), "ltt_vit_mini")):
            fin = r.t_final(r.t_config(**ocfg.get_config(cname))).to(DEV).eval()
            with torch.no_grad():
                lg, at = r.fw_final(fin, xs)
            assert lg.shape == (0, C) and at.shape == (0, C, n)
    assert ash.mask_shapley_new(0, n, device=DEV).shape == (0, n)
    pm = ash.mask_shapley_new(0, n, device=DEV, rng="philox", seed=1, packed=True)
    assert pm.rows == 0


@pytest.mark.parametrize("name", ["vit_mini", "bert_mini"])
@pytest.mark.parametrize("B,S", [(1, 1), (2, 3), (3, 5), (5, 2)])
def test_ragged_coalition_shapes_match_the_replicated_call(agb, name, B, S):
    """(B, S, n) fast path == the reference-shaped call on inputs replicated S times (row order b*S+s), for S odd / B odd,
    where Shapley pairs straddle inputs (the reference only asserts that B*S is even, models/shapley.py:62)."""
    rec, cfgd, cfg, srg, exp = _build(name, "fp32")
    n = rec.n_players(cfg)
    xs = torch.from_numpy(synth.inputs(cfgd, B, seed=3)).to(DEV)
    g = torch.Generator(device="cpu").manual_seed(B * 10 + S)
    masks = (torch.rand((B * S, n), generator=g) > 0.5).to(torch.int64).to(DEV)
    with torch.no_grad():
        a, _ = rec.fw_surrogate(srg, xs.repeat_interleave(S, dim=0), masks)
        b, _ = rec.fw_surrogate(srg, xs, masks.reshape(B, S, n))
    assert torch.equal(a, b)
    srg.agb_precision = "bf16"
    with torch.no_grad():
        c, _ = rec.fw_surrogate(srg, xs, masks.reshape(B, S, n))
    np.testing.assert_allclose(_np(c), _np(a), atol=2e-2)
    np.testing.assert_allclose(_np(c).sum(1), 1.0, atol=1e-5)


@pytest.mark.parametrize("name", ["vit_mini", "bert_mini"])
def test_row_chunking_is_invisible(agb, name):
    """More (input, coalition) rows than the engine's chunk (1024): results equal the per-chunk calls."""
    rec, cfgd, cfg, srg, exp = _build(name, "bf16")
    n = rec.n_players(cfg)
    B, S = 70, 16                      # 1120 rows -> two chunks of 64 and 6 inputs
    xs = torch.from_numpy(synth.inputs(cfgd, B, seed=5)).to(DEV)
    from autognothi_b200.models import shapley as ash
    pm = ash.mask_shapley_new(B * S, n, device=DEV, rng="philox", seed=9, packed=True)
    with torch.no_grad():
        full, _ = rec.fw_surrogate(srg, xs, pm)
        first, _ = rec.fw_surrogate(srg, xs[:64], ash.PackedMasks(pm.words[:64 * S].contiguous(), n))
        rest, _ = rec.fw_surrogate(srg, xs[64:], ash.PackedMasks(pm.words[64 * S:].contiguous(), n))
    assert full.shape == (B * S, cfgd["num_labels"])
    assert torch.equal(full, torch.cat([first, rest], 0))


@pytest.mark.parametrize("name", ["vit_mini", "vit_tiny", "bert_mini"])
def test_cuda_graph_replay_of_small_calls_is_invisible(agb, name):
    """Calls of <= GRAPH_MAX_ROWS rows are captured once per shape and replayed: bit-identical to the eager launches, for
    fresh inputs on every replay, across alternating shapes, and after the weights change (the engine is rebuilt)."""
    from autognothi_b200 import engine
    rec, cfgd, cfg, srg, exp = _build(name, "bf16")
    n = rec.n_players(cfg)
    old_rows, old_drop = engine.GRAPH_MAX_ROWS, engine.DROP_MASKED_TOKENS
    engine.DROP_MASKED_TOKENS = False           # the packed BERT path is never captured (data-dependent buffer sizes)
    try:
        for it in range(6):
            B, S = ((2, 4), (1, 8), (3, 2))[it % 3]
            xs = torch.from_numpy(synth.inputs(cfgd, B, seed=10 + it)).to(DEV)
            g = torch.Generator(device="cpu").manual_seed(100 + it)
            masks = (torch.rand((B, S, n), generator=g) > 0.5).to(torch.int64).to(DEV)
            with torch.no_grad():
                engine.GRAPH_MAX_ROWS = 128
                a, _ = rec.fw_surrogate(srg, xs, masks)
                engine.GRAPH_MAX_ROWS = 0
                b, _ = rec.fw_surrogate(srg, xs, masks)
            assert torch.equal(a, b), (it, float((a - b).abs().max()))
            if it == 3:
                with torch.no_grad():
                    srg.classifier.weight.mul_(1.5)        # bumps the parameter version -> packed weights and graphs rebuilt
    finally:
        engine.GRAPH_MAX_ROWS, engine.DROP_MASKED_TOKENS = old_rows, old_drop


@pytest.mark.parametrize("name", ["vit_mini", "bert_mini"])
def test_cuda_graph_replay_of_small_explainer_calls_is_invisible(agb, name):
    from autognothi_b200 import engine
    rec, cfgd, cfg, srg, exp = _build(name, "bf16")
    n, C = rec.n_players(cfg), cfgd["num_labels"]
    old_rows = engine.GRAPH_MAX_ROWS
    try:
        for it in range(4):
            B = (2, 3)[it % 2]
            xs = torch.from_numpy(synth.inputs(cfgd, B, seed=20 + it)).to(DEV)
            g = torch.Generator(device="cpu").manual_seed(200 + it)
            masks = (torch.rand((B, n), generator=g) > 0.3).to(torch.int64).to(DEV)
            grand, null = torch.rand((B, C), generator=g).to(DEV), torch.rand((1, C), generator=g).to(DEV)
            with torch.no_grad():
                engine.GRAPH_MAX_ROWS = 128
                a, _ = rec.fw_explainer(exp, xs, masks, grand, null)
                engine.GRAPH_MAX_ROWS = 0
                b, _ = rec.fw_explainer(exp, xs, masks, grand, null)
            assert torch.equal(a, b), (it, float((a - b).abs().max()))
    finally:
        engine.GRAPH_MAX_ROWS = old_rows


@pytest.mark.parametrize("name", ["vit_mini", "vit_tiny", "vit_base"])
def test_kept_first_token_order_is_exact_work_skipping(agb, golden_dir, name):
    """ViT surrogate: permuting every row's tokens so that the kept ones come first and folding the masked keys (all with
    the logit 0) into one virtual key changes nothing but rounding — incl. the empty and the full coalition."""
    import os
    from autognothi_b200 import engine
    g = np.load(os.path.join(golden_dir, f"model_{name}.npz"))
    B, S, n = (int(v) for v in g["meta"])
    rec, cfgd, cfg, srg, exp = _build(name, "bf16")
    xs = torch.from_numpy(synth.inputs(cfgd, B, seed=0)).to(DEV)
    masks = torch.from_numpy(g["masks"].astype(np.int64)).to(DEV).reshape(B, S, n)
    edge = masks.clone()
    edge[0, 0, :] = 0
    edge[0, 1, :] = 1
    edge[0, 1, 5] = 0
    old = engine.KEPT_FIRST_ORDER
    try:
        with torch.no_grad():
            engine.KEPT_FIRST_ORDER = True
            a, _ = rec.fw_surrogate(srg, xs, masks)
            ae, _ = rec.fw_surrogate(srg, xs, edge)
            engine.KEPT_FIRST_ORDER = False
            b, _ = rec.fw_surrogate(srg, xs, masks)
            be, _ = rec.fw_surrogate(srg, xs, edge)
    finally:
        engine.KEPT_FIRST_ORDER = old
    np.testing.assert_allclose(_np(a), _np(b), atol=3e-3)
    np.testing.assert_allclose(_np(ae), _np(be), atol=3e-3)
    np.testing.assert_allclose(_np(a), g["v_s"], atol=2e-2)
    np.testing.assert_allclose(_np(a).sum(1), 1.0, atol=1e-5)


@pytest.mark.parametrize("T,rows", [(197, 33), (17, 9), (128, 4), (208, 3), (33, 70), (2, 3), (256, 2)])
def test_kept_first_order_kernel_vs_numpy(agb, T, rows):
    """agb_kept_first_order: stable kept/masked partition (bit-exact index work), incl. empty and full coalitions."""
    rng = np.random.default_rng(T * 31 + rows)
    dense = (rng.random((rows, max(T - 1, 0))) > rng.random((rows, 1))).astype(np.int64)
    if T > 1:
        dense[0, :] = 0
        dense[1, :] = 1
    packed = agb.pack_masks(torch.from_numpy(dense).to(DEV), prepend_cls=True)
    order, pos, nkeep, prefix = agb.kept_first_order(packed, T)
    keep = np.concatenate([np.ones((rows, 1), np.int64), dense], 1)
    ref_order = np.argsort(1 - keep, axis=1, kind="stable")
    assert np.array_equal(order.cpu().numpy().astype(np.int64), ref_order)
    ref_pos = np.empty_like(ref_order)
    np.put_along_axis(ref_pos, ref_order, np.arange(T)[None, :].repeat(rows, 0), axis=1)
    assert np.array_equal(pos.cpu().numpy().astype(np.int64), ref_pos)
    assert np.array_equal(nkeep.cpu().numpy(), keep.sum(1).astype(np.int32))
    ref_prefix = agb.pack_masks(torch.from_numpy((np.arange(T - 1)[None, :] < (keep.sum(1)[:, None] - 1)).astype(np.int64)).to(DEV),
                                prepend_cls=True)
    assert torch.equal(prefix, ref_prefix)


@pytest.mark.parametrize("B,S,T,H,dtype", [(3, 4, 197, 768, torch.float32), (2, 5, 17, 24, torch.bfloat16), (1, 1, 33, 8, torch.float32)])
def test_gather_token_rows_equals_index_select(agb, B, S, T, H, dtype):
    torch.manual_seed(B * 7 + T)
    x = torch.randn((B, T, H), device=DEV).to(dtype)
    order = torch.stack([torch.randperm(T, device=DEV) for _ in range(B * S)]).to(torch.uint8)
    got = agb.gather_token_rows(x, order, S)
    r = torch.arange(B * S, device=DEV)
    ref = x[(r // S)[:, None], order.long()]
    assert torch.equal(got, ref)


@pytest.mark.parametrize("T,heads,rows,share", [(197, 12, 8, 4), (197, 3, 3, 1), (128, 2, 6, 3), (17, 1, 4, 2), (208, 2, 2, 1), (65, 1, 5, 5)])
def test_split_softmax_attention_kernels(agb, T, heads, rows, share):
    """Third-generation attention kernel (agb_attention_split.cu), all three entry points against the fp32 CUDA-core kernel:
    ViT masks in token order, the same with scattered output rows, and kept-first order with the virtual key."""
    torch.manual_seed(T + heads + rows)
    H = heads * 64
    B = rows // share
    qkv_b = torch.randn(B * T, 3 * H, device=DEV).to(torch.bfloat16)
    qkv = qkv_b.reshape(B, 1, T, 3 * H).expand(B, share, T, 3 * H).reshape(rows * T, 3 * H).contiguous()
    g = torch.Generator(device="cpu").manual_seed(rows)
    dense = (torch.rand((rows, T - 1), generator=g) > torch.rand((rows, 1), generator=g)).to(torch.int64).to(DEV)
    dense[0, :] = 0
    if rows > 1:
        dense[1, :] = 1
    packed = agb.pack_masks(dense, prepend_cls=True)
    ref = agb.masked_attention(qkv.float(), packed, T, heads, agb.MASK_MUL0).float()           # fp32 exact kernel
    scale = float(ref.abs().max())
    # (1) token order, shared projections
    got = agb.masked_attention(qkv_b, packed, T, heads, agb.MASK_MUL0, share=share).float()
    assert torch.isfinite(got).all()
    assert float((got - ref).abs().max()) <= 3e-2 * scale and float((got - ref).norm() / ref.norm()) < 1e-2
    # (2) the same, output rows scattered into kept-first order
    order, pos, nkeep, prefix = agb.kept_first_order(packed, T)
    sc = agb.masked_attention_scatter(qkv_b, packed, T, heads, share, pos).float().reshape(rows, T, H)
    r = torch.arange(rows, device=DEV)
    assert torch.equal(sc[r[:, None], pos.long()], got.reshape(rows, T, H))
    # (3) kept-first order (split kernel = variant 3): permuted projections + nkeep -> the permuted rows of the reference
    from autognothi_b200 import _native as nat
    qkv_p = qkv.reshape(rows, T, 3 * H)[r[:, None], order.long()].reshape(rows * T, 3 * H).contiguous()
    prev = nat.lib.agb_attention_set_variant(3)
    try:
        pf = agb.attention_prefix(qkv_p, nkeep, T, heads).float().reshape(rows, T, H)
    finally:
        nat.lib.agb_attention_set_variant(prev)
    ref_p = ref.reshape(rows, T, H)[r[:, None], order.long()]
    assert torch.isfinite(pf).all()
    assert float((pf - ref_p).abs().max()) <= 3e-2 * scale and float((pf - ref_p).norm() / ref_p.norm()) < 1e-2
    # and the pipelined second-generation kernel (variant 2) agrees on (1) and (3)
    prev = nat.lib.agb_attention_set_variant(2)
    try:
        got2 = agb.masked_attention(qkv_b, packed, T, heads, agb.MASK_MUL0, share=share).float()
        pf2 = agb.attention_prefix(qkv_p, nkeep, T, heads).float().reshape(rows, T, H)
    finally:
        nat.lib.agb_attention_set_variant(prev)
    assert float((got2 - got).abs().max()) <= 2e-2 * scale
    assert float((pf2 - pf).abs().max()) <= 2e-2 * scale


@pytest.mark.parametrize("hot_keys", [(40,), (150,), (40, 150), (100, 196), (33, 70, 120, 190)])
def test_split_softmax_lazy_maximum_rescale(agb, hot_keys):
    """The split-softmax kernel reads S from TMEM once: the row's reference maximum comes from the first chunk of each key
    half and is raised later only when a chunk exceeds it by 2^24, rescaling the P columns already written.  Keys with
    huge logits late in the row force that path (and the reconciliation between the two halves); result vs the fp32 kernel."""
    torch.manual_seed(sum(hot_keys))
    T, heads, rows = 197, 2, 6
    H = heads * 64
    qkv = torch.randn(rows, T, 3 * H, device=DEV)
    for i, kx in enumerate(hot_keys):
        qkv[:, kx, H:2 * H] *= 60.0 * (i + 1)          # logits of this key: std ~ 140 * (i + 1), far above the first chunk's
    qkv_b = qkv.reshape(rows * T, 3 * H).to(torch.bfloat16).contiguous()
    g = torch.Generator(device="cpu").manual_seed(3)
    dense = (torch.rand((rows, T - 1), generator=g) > 0.3).to(torch.int64).to(DEV)
    dense[1, :] = 1
    for kx in hot_keys:
        dense[:, kx - 1] = 1                           # the hot keys stay live (a masked key's logit is 0)
    packed = agb.pack_masks(dense, prepend_cls=True)
    ref = agb.masked_attention(qkv_b.float(), packed, T, heads, agb.MASK_MUL0).float()
    got = agb.masked_attention(qkv_b, packed, T, heads, agb.MASK_MUL0).float()
    assert torch.isfinite(got).all()
    scale = float(ref.abs().max())
    assert float((got - ref).abs().max()) <= 3e-2 * scale and float((got - ref).norm() / ref.norm()) < 1e-2
    order, pos, nkeep, prefix = agb.kept_first_order(packed, T)
    r = torch.arange(rows, device=DEV)
    qkv_p = qkv_b.reshape(rows, T, 3 * H)[r[:, None], order.long()].reshape(rows * T, 3 * H).contiguous()
    from autognothi_b200 import _native as nat
    prev = nat.lib.agb_attention_set_variant(3)
    try:
        pf = agb.attention_prefix(qkv_p, nkeep, T, heads).float().reshape(rows, T, H)
    finally:
        nat.lib.agb_attention_set_variant(prev)
    ref_p = ref.reshape(rows, T, H)[r[:, None], order.long()]
    assert torch.isfinite(pf).all()
    assert float((pf - ref_p).abs().max()) <= 3e-2 * scale and float((pf - ref_p).norm() / ref_p.norm()) < 1e-2


@pytest.mark.parametrize("shape,S,dtype", [((3, 197, 768), 32, torch.float32), ((5, 17, 24), 7, torch.bfloat16),
                                           ((1, 128, 768), 2, torch.float32), ((4, 3, 5), 3, torch.float32),
                                           ((0, 8, 16), 4, torch.float32)])
def test_repeat_rows_equals_repeat_interleave(agb, shape, S, dtype):
    """agb_repeat_rows (the residual stream of an input fanned out to its S coalition rows) is bit-identical to
    torch.repeat_interleave; rows that are not whole 16-byte vectors take the torch path."""
    torch.manual_seed(1)
    x = torch.randn(shape, device=DEV).to(dtype)
    got = agb.repeat_rows(x, S)
    assert got.shape == (shape[0] * S,) + shape[1:] and got.dtype == dtype
    assert torch.equal(got, x.repeat_interleave(S, dim=0))
    assert got.data_ptr() != x.data_ptr() or x.numel() == 0


@pytest.mark.parametrize("rows,T,H,C", [(1003, 2, 768, 2), (520, 1, 256, 10)])
def test_bert_head_rows_per_cta_kernel_is_bit_identical(agb, rows, T, H, C):
    """>= 512 rows take cls_head_pool_rows_kernel (8 rows per CTA, pooler weight streamed once per 8 rows); the same rows in
    chunks of < 512 take the one-row kernel.  Same summation order per dot product -> identical bits; also checked
    against torch (reference models/vanilla_bert.py:73-77, 615-619)."""
    torch.manual_seed(5)
    x = torch.randn(rows, T, H, device=DEV)
    wp, bp = torch.randn(H, H, device=DEV) / H ** 0.5, torch.randn(H, device=DEV) * 0.1
    wc, bc = torch.randn(C, H, device=DEV) / H ** 0.5, torch.randn(C, device=DEV) * 0.1
    probs, logits = agb.cls_head(x, 1, wc, bc, pool=(wp, bp), want_logits=True)
    parts = [agb.cls_head(x[r0:r0 + 500].contiguous(), 1, wc, bc, pool=(wp, bp), want_logits=True) for r0 in range(0, rows, 500)]
    assert torch.equal(probs, torch.cat([p for p, _ in parts], 0))
    assert torch.equal(logits, torch.cat([l for _, l in parts], 0))
    ref = torch.tanh(x[:, 0].double() @ wp.double().t() + bp.double()) @ wc.double().t() + bc.double()
    torch.testing.assert_close(logits.double(), ref, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(probs.double(), torch.softmax(ref, -1), rtol=1e-4, atol=1e-6)


def test_normalize_and_loss_follow_the_reference_on_gradient_and_dtype_edges(agb):
    """a `null` that alone requires grad gets its gradient, in the shape it was passed ((1, C) or (C,)), and
    loss_shapley_new accepts a non-fp32 phi (computed in fp32) — compared with the formulas of reference
    models/shapley.py:82-93 and 40-50 evaluated by torch autograd."""
    from autognothi_b200.models import shapley as ash
    torch.manual_seed(4)
    B, T, C, S = 3, 9, 4, 6
    n = T - 1
    pred = torch.randn(B, T, C, device=DEV)
    grand = torch.rand(B, C, device=DEV)
    for shape in ((1, C), (C,)):
        null = torch.rand(shape, device=DEV, requires_grad=True)
        out = ash.normalize_shapley_explanation(pred, grand, null)
        out.square().sum().backward()
        null2 = null.detach().clone().requires_grad_(True)
        ref = pred + ((grand - null2.reshape(1, C)) - pred.sum(dim=1)).unsqueeze(1) / T
        ref.square().sum().backward()
        assert null.grad is not None and null.grad.shape == null.shape
        torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(null.grad, null2.grad, rtol=1e-4, atol=1e-5)
    masks = (torch.rand(B, S, n, device=DEV) > 0.5).to(torch.int64)
    phi = torch.randn(B, C, n, device=DEV)
    v0, vs = torch.rand(1, C, device=DEV), torch.rand(B * S, C, device=DEV)
    l32 = ash.loss_shapley_new(B, S, n, masks, v0, vs, None, phi)
    l16 = ash.loss_shapley_new(B, S, n, masks, v0, vs, None, phi.to(torch.bfloat16))
    ref = n * ((v0 + torch.bmm(masks.float(), phi.permute(0, 2, 1)).reshape(B * S, C) - vs) ** 2).mean()
    torch.testing.assert_close(l32, ref, rtol=1e-5, atol=1e-6)
    assert abs(float(l16) - float(ref)) <= 3e-2 * abs(float(ref))
