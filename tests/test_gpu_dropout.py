"""Training-mode dropout (reference: nn.Dropout modules of models/vanilla_vit.py:253,457,501-503,512-516 and
models/vanilla_bert.py:325,530,559,603 active in train()).  The masks come from a counter hash, not torch's generator, so
parity is checked against a torch re-statement that uses the EXPORTED masks: same forward, and the same gradients from
torch autograd, within bf16 tolerance.  Plus the statistics of the masks themselves."""
import numpy as np
import pytest
import torch

from oracle import configs as ocfg
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _np(t):
    return t.detach().float().cpu().numpy()


@pytest.mark.parametrize("p", [0.1, 0.5])
def test_elementwise_dropout_statistics_residual_and_adjoint(agb, p):
    thr = agb.dropout_thr(p)
    n = 1 << 20
    y = torch.ones(n, device=DEV)
    res = torch.full((n,), 3.0, device=DEV)
    out = agb.dropout(y, thr, seed=1234, tag=7, residual=res)
    kept = (out != 3.0)
    frac = float(kept.float().mean())
    assert abs(frac - (1 - p)) < 4 * np.sqrt(p * (1 - p) / n) + 1e-4, frac
    scale = 65536.0 / (65536.0 - thr)
    np.testing.assert_allclose(_np(out[kept]), 3.0 + scale, rtol=1e-6)
    # deterministic in (seed, tag); different streams differ; bf16 in / out uses the same mask
    assert torch.equal(out, agb.dropout(y, thr, seed=1234, tag=7, residual=res))
    assert not torch.equal(out, agb.dropout(y, thr, seed=1234, tag=8, residual=res))
    assert not torch.equal(out, agb.dropout(y, thr, seed=1235, tag=7, residual=res))
    out16 = agb.dropout(y.to(torch.bfloat16), thr, seed=1234, tag=7, out_dtype=torch.bfloat16)
    assert torch.equal(out16 != 0, kept)
    # the adjoint is the same map on the gradient: <dropout(y), g> == <y, dropout(g)>
    g = torch.randn(n, device=DEV)
    yy = torch.randn(n, device=DEV)
    lhs = float((agb.dropout(yy, thr, 99, 3) * g).double().sum())
    rhs = float((yy * agb.dropout(g, thr, 99, 3)).double().sum())
    # the two sides round y * scale resp. g * scale in fp32: ~6e-8 relative per term, a random walk over n = 2^20 terms
    assert abs(lhs - rhs) <= 1e-3
    # no visible structure: neighbouring elements are uncorrelated
    k = kept.float() - (1 - p)
    assert abs(float((k[:-1] * k[1:]).mean())) < 5e-3


def _torch_attention(qkv, dense_tok, keep, T, heads, mode, scale_keep):
    """fp32 re-statement: reference models/vanilla_vit.py:444-459 (mode 0) / vanilla_bert.py:517-532 (mode 1) with the
    dropout mask `keep` (rows, heads, T, T) applied to the probabilities."""
    rows = dense_tok.shape[0]
    H = qkv.shape[1] // 3
    d = H // heads
    q, k, v = (qkv[:, i * H:(i + 1) * H].reshape(rows, T, heads, d).permute(0, 2, 1, 3) for i in range(3))
    s = q @ k.transpose(-1, -2) / np.sqrt(d)
    m = dense_tok.reshape(rows, 1, 1, T).float()
    s = s * m if mode == 0 else s + (1.0 - m) * torch.finfo(torch.float32).min
    pr = torch.softmax(s, dim=-1) * keep.float() * scale_keep
    return (pr @ v).permute(0, 2, 1, 3).reshape(rows * T, H)


@pytest.mark.parametrize("M,N,K", [(300, 768, 768), (12608, 768, 3072), (25216 + 40, 768, 768)])
def test_gemm_with_fused_dropout_residual(agb, M, N, K):
    """agb_gemm_bf16_dropout_residual == agb_gemm_bf16 followed by agb_dropout with the same (seed, tag): identical keep mask
    (the adjoint regenerates it through agb_dropout), values equal up to the bf16 rounding the unfused path applies to the
    dense output in between."""
    torch.manual_seed(11)
    a = torch.randn(M, K, device=DEV).bfloat16()
    w = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
    b = torch.randn(N, device=DEV) * 0.1
    res = torch.randn(M, N, device=DEV)
    thr, seed, tag = agb.dropout_thr(0.1), 0x1234567890ABCDEF, 5
    fused = agb.gemm_bf16_dropout_residual(a, w, b, res, thr, seed, tag)
    y = agb.gemm_bf16(a, w, b)
    unfused = agb.dropout(y, thr, seed, tag, residual=res, out_dtype=torch.float32)
    keep = agb.dropout(torch.ones(M, N, device=DEV), thr, seed, tag) != 0
    assert abs(float(keep.float().mean()) - 0.9) < 5e-3
    assert torch.equal(fused[~keep], res[~keep])                       # dropped elements: the residual passes through untouched
    scale = 65536.0 / (65536.0 - thr)
    ref = res + keep * (a.float() @ w.float().t() + b) * scale
    torch.testing.assert_close(fused, ref, rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(fused, unfused, rtol=1e-2, atol=1e-2)   # y rounded to bf16 on the unfused side


@pytest.mark.parametrize("M", [300, 12608, 25216 + 8])
def test_gemm_with_fused_gelu_forward_and_adjoint(agb, M):
    """agb_gemm_bf16_gelu_dual == GEMM then GELU kernel (bit-identical z; GELU evaluated at the rounded z on both sides), and
    agb_gemm_bf16_gelu_bwd == dgrad GEMM then GELU-adjoint kernel up to the bf16 rounding of the intermediate gradient; both
    against torch fp32."""
    torch.manual_seed(13)
    K, N = 768, 3072
    a = torch.randn(M, K, device=DEV).bfloat16()
    w1 = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
    b1 = torch.randn(N, device=DEV) * 0.1
    z, f = agb.gemm_bf16_gelu_dual(a, w1, b1)
    z_ref = agb.gemm_bf16(a, w1, b1)
    assert torch.equal(z, z_ref)
    assert torch.equal(f, agb.gelu_fwd(z_ref))
    torch.testing.assert_close(f.float(), torch.nn.functional.gelu(a.float() @ w1.float().t() + b1), rtol=2e-2, atol=2e-2)
    # adjoint: dz = (dy W2) * GELU'(z), W2 (K, N) = forward weight of the layer behind the GELU
    dy = torch.randn(M, K, device=DEV).bfloat16()
    w2 = (torch.randn(K, N, device=DEV) / N ** 0.5).bfloat16()
    dz = agb.gemm_bf16_gelu_bwd(dy, w2, z)
    dz_unfused = agb.gelu_bwd(agb.gemm_bf16(dy, w2, w_mn=True), z)
    zr = z.float().requires_grad_(True)
    torch.nn.functional.gelu(zr).backward(dy.float() @ w2.float())
    torch.testing.assert_close(dz.float(), zr.grad, rtol=2e-2, atol=5e-3)
    assert float((dz.float() - zr.grad).norm() / zr.grad.norm()) < 5e-3
    assert float((dz.float() - dz_unfused.float()).norm() / dz_unfused.float().norm()) < 5e-3


def test_gelu_bf16_fast_forms_vs_torch(agb):
    """bf16 GELU forward / adjoint kernels (one MUFU each: the tanh-form refit of erf-GELU and ITS derivative) against torch's
    exact erf GELU and autograd; fp32 I/O keeps the exact erf / exp forms."""
    torch.manual_seed(12)
    z = (torch.randn(4096, 768, device=DEV) * 2.5)
    dy = torch.randn(4096, 768, device=DEV)
    zr = z.clone().requires_grad_(True)
    f_ref = torch.nn.functional.gelu(zr)
    f_ref.backward(dy)
    f32 = agb.gelu_fwd(z)
    d32 = agb.gelu_bwd(dy, z)
    torch.testing.assert_close(f32, f_ref.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(d32, zr.grad, rtol=1e-4, atol=1e-5)
    z16, dy16 = z.bfloat16(), dy.bfloat16()
    z16r = z16.float().requires_grad_(True)
    f16_ref = torch.nn.functional.gelu(z16r)
    f16_ref.backward(dy16.float())
    f16 = agb.gelu_fwd(z16)
    d16 = agb.gelu_bwd(dy16, z16)
    assert f16.dtype == torch.bfloat16 and d16.dtype == torch.bfloat16
    # approximant within 3e-4 (forward) / 9e-4 (derivative) of the exact forms, then one bf16 rounding
    torch.testing.assert_close(f16.float(), f16_ref.detach(), rtol=8e-3, atol=1e-3)
    torch.testing.assert_close(d16.float(), z16r.grad, rtol=8e-3, atol=4e-3)


@pytest.mark.parametrize("T,heads,d,mode", [(197, 2, 64, 0), (128, 2, 64, 1), (33, 1, 64, 1), (256, 1, 64, 0), (197, 3, 16, 0),
                                            (128, 4, 8, 1), (40, 2, 32, 0)])
def test_attention_dropout_forward_and_adjoint_vs_torch(agb, T, heads, d, mode):
    torch.manual_seed(T + d + mode)
    rows, H, p = 3, heads * d, 0.1
    thr = agb.dropout_thr(p)
    seed = 0x1234567 + T
    qkv16 = (torch.randn(rows * T, 3 * H, device=DEV) * 1.2).to(torch.bfloat16)
    dctx16 = torch.randn(rows * T, H, device=DEV).to(torch.bfloat16)
    dense = (torch.rand(rows, T - 1, device=DEV) > 0.35).to(torch.int64)
    dense[0] = 1
    masks = agb.pack_masks(dense, prepend_cls=True)
    tok = torch.cat([torch.ones((rows, 1), dtype=torch.int64, device=DEV), dense], 1)
    keep = agb.attention_dropout_mask(rows, heads, T, thr, seed, DEV)
    frac = float(keep.float().mean())
    assert abs(frac - 0.9) < 4 * np.sqrt(0.09 / keep.numel()) + 1e-3, frac
    q32 = qkv16.float().requires_grad_(True)
    ref = _torch_attention(q32, tok, keep, T, heads, mode, 65536.0 / (65536.0 - thr))
    ref.backward(dctx16.float())
    got = agb.masked_attention_dropout(qkv16, masks, T, heads, mode, thr, seed)
    scale = float(ref.detach().abs().max())
    np.testing.assert_allclose(_np(got), _np(ref), rtol=2e-2, atol=2e-2 * scale)
    dq = agb.masked_attention_dropout_bwd(qkv16, dctx16, masks, T, heads, mode, thr, seed)
    gref = q32.grad
    a, b = _np(dq).reshape(-1).astype(np.float64), _np(gref).reshape(-1).astype(np.float64)
    rel = float(np.linalg.norm(a - b) / np.linalg.norm(b))
    assert rel < 2e-2, rel
    # and it is NOT the undropped attention
    plain = agb.masked_attention(qkv16, masks, T, heads, mode)
    assert float((plain.float() - got.float()).abs().max()) > 0.05 * scale


@pytest.mark.parametrize("T,heads,d,mode", [(197, 2, 64, 0), (64, 2, 64, 1), (33, 2, 16, 1)])
def test_attention_dropout_fp32_mode_vs_torch(agb, T, heads, d, mode):
    """The exact (fp32, CUDA-core) mode applies the same counter-hash mask: forward and adjoint against torch autograd."""
    torch.manual_seed(T + d)
    rows, H, p = 2, heads * d, 0.1
    thr = agb.dropout_thr(p)
    seed = 0xABCDEF + T
    qkv = torch.randn(rows * T, 3 * H, device=DEV) * 1.2
    dctx = torch.randn(rows * T, H, device=DEV)
    dense = (torch.rand(rows, T - 1, device=DEV) > 0.35).to(torch.int64)
    masks = agb.pack_masks(dense, prepend_cls=True)
    tok = torch.cat([torch.ones((rows, 1), dtype=torch.int64, device=DEV), dense], 1)
    keep = agb.attention_dropout_mask(rows, heads, T, thr, seed, DEV)
    q32 = qkv.clone().requires_grad_(True)
    ref = _torch_attention(q32, tok, keep, T, heads, mode, 65536.0 / (65536.0 - thr))
    ref.backward(dctx)
    got = agb.masked_attention_dropout(qkv, masks, T, heads, mode, thr, seed)
    np.testing.assert_allclose(_np(got), _np(ref), rtol=1e-4, atol=1e-5)
    dq = agb.masked_attention_dropout_bwd(qkv, dctx, masks, T, heads, mode, thr, seed)
    np.testing.assert_allclose(_np(dq), _np(q32.grad), rtol=2e-3, atol=2e-4 * float(q32.grad.abs().max()))


@pytest.mark.parametrize("name", ["vit_mini", "bert_mini"])
def test_explainer_training_with_dropout_is_seeded_and_learns(agb, golden_dir, name):
    """train() mode on the drop-in explainer: dropout active (gradients differ from the p = 0 path), reproducible under
    torch.manual_seed, and the reference's loop body still reduces the loss."""
    import os
    from autognothi_b200.models import shapley as ash
    from autognothi_b200.recipes.vanilla_bert import vanilla_bert_recipe
    from autognothi_b200.recipes.vanilla_vit import vanilla_vit_recipe
    g = np.load(os.path.join(golden_dir, f"model_{name}.npz"))
    B, S, n = (int(v) for v in g["meta"])
    cfgd = ocfg.get_config(name)
    rec = vanilla_vit_recipe() if ocfg.is_vit(cfgd) else vanilla_bert_recipe()
    exp = rec.t_explainer(rec.t_config(**cfgd))
    exp.load_state_dict({k: torch.from_numpy(v) for k, v in synth.explainer_state(cfgd, seed=1).items()}, strict=True)
    exp = exp.to(DEV).train()
    exp.agb_precision = "bf16"
    xs = torch.from_numpy(synth.inputs(cfgd, B, seed=0)).to(DEV)
    masks = torch.from_numpy(g["masks"].astype(np.int64)).to(DEV).reshape(B, S, n)
    v_s, grand, null = (torch.from_numpy(g[k]).to(DEV) for k in ("v_s", "grand", "null"))
    ones = torch.ones((B, n), dtype=torch.int64, device=DEV)

    def grads(seed, dropout=True):
        exp.agb_dropout = dropout
        exp.zero_grad(set_to_none=True)
        torch.manual_seed(seed)
        phi, _ = rec.fw_explainer(exp, xs, ones, grand, null)
        loss = ash.loss_shapley_new(B, S, n, masks, null, v_s, grand, phi)
        loss.backward()
        key = "explainer_mlp.1.weight" if ocfg.is_vit(cfgd) else "explainer_mlp.0.weight"
        return float(loss.detach()), dict(exp.named_parameters())[key].grad.clone()

    l1, g1 = grads(5)
    l2, g2 = grads(5)
    l3, g3 = grads(6)
    l0, g0 = grads(5, dropout=False)
    assert l1 == l2 and torch.equal(g1, g2)
    assert not torch.equal(g1, g3)
    assert not torch.equal(g1, g0) and np.isfinite(l1) and np.isfinite(l3)
    cos = float((g1 * g0).sum() / (g1.norm() * g0.norm()))
    assert cos > 0.2, cos          # positively aligned with the p = 0 gradient, but a different (noisy) estimate
    exp.agb_dropout = True
    opt = torch.optim.AdamW(exp.parameters(), lr=1e-4)
    losses = []
    torch.manual_seed(0)
    for _ in range(10):
        opt.zero_grad()
        phi, _ = rec.fw_explainer(exp, xs, ones, grand, null)
        loss = ash.loss_shapley_new(B, S, n, masks, null, v_s, grand, phi)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert np.isfinite(losses).all() and min(losses[-3:]) < 0.8 * losses[0], losses
    # eval() switches it off again: two calls agree bit for bit
    exp.eval()
    with torch.no_grad():
        a, _ = rec.fw_explainer(exp, xs, ones, grand, null)
        b, _ = rec.fw_explainer(exp, xs, ones, grand, null)
    assert torch.equal(a, b)
