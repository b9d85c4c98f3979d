#!/usr/bin/env python
"""Throughput of the non-headline configurations of BASELINE.json (parity-test cases, measured once for the record):
KernelSHAP batched Gram + Cholesky solve, BERT-base T=128 masked evaluation, surrogate training.
Test infrastructure; run under gpurun:  python tools/side_benches.py"""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from autognothi_b200 import ops  # noqa: E402
from autognothi_b200.models import shapley as ash  # noqa: E402


def timed(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def kernelshap():
    dev = torch.device("cuda:0")
    for d, S, C, B in ((128, 2048, 2, 256), (128, 2048, 2, 1024), (197, 2048, 10, 128), (197, 2048, 10, 512), (512, 2048, 2, 32),
                       (512, 2048, 2, 148)):
        n = d
        dense = (torch.rand(B * S, n - 1, device=dev) > 0.5).to(torch.int64)
        Zp = ops.pack_masks(dense, prepend_cls=True).reshape(B, S, -1)
        w = torch.rand(B, S, device=dev) + 0.1
        probs = torch.rand(B, S, C, device=dev) * 0.9 + 0.05
        fx = torch.rand(B, C, device=dev) * 0.9 + 0.05
        fnull = torch.rand(C, device=dev) * 0.9 + 0.05
        ms = timed(lambda: ops.kernelshap_solve(Zp, w, probs, fx, fnull, d))
        bytes_s = S * Zp.shape[2] * 4 + S * C * 8 + S * 8 + (d - 1) * (d - 1) * 8
        flops = 2.0 * S * (d - 1) ** 2 + (d - 1) ** 3 / 3.0          # full Gram + Cholesky (what the definition counts)
        tiles = (d - 1 + 63) // 64
        executed = 2.0 * S * 64 * 64 * tiles * (tiles + 1) / 2 + (d - 1) ** 3 / 3.0   # lower-triangle 64x64 tiles only
        print(f"kernelshap d={d} S={S} C={C} B={B}: {ms:8.3f} ms  {B / ms * 1e3:10.0f} solves/s  "
              f"{B * bytes_s / ms * 1e-6:7.1f} GB/s algorithmic  {B * flops / ms * 1e-9:7.2f} TFLOP/s fp64 dense-equivalent, "
              f"{B * executed / ms * 1e-9:7.2f} TFLOP/s executed (fp64 FMA pipe; peak measured by tools/fp64_peak.cu)")


def bert_eval():
    from autognothi_b200.recipes.vanilla_bert import vanilla_bert_recipe
    dev = torch.device("cuda:0")
    rec = vanilla_bert_recipe()
    # reference experiments/bert_base_tayp_vanilla/.hparams.json:14-30 with max_position_embeddings = 128
    cfgd = dict(attention_probs_dropout_prob=0.1, explainer_attn_num_layers=1, explainer_head_hidden_size=3072,
                explainer_normalize=True, hidden_dropout_prob=0.1, hidden_size=768, intermediate_size=3072, layer_norm_eps=1e-12,
                max_position_embeddings=128, num_attention_heads=12, num_hidden_layers=12, num_labels=2, pad_token_id=0,
                type_vocab_size=2, vocab_size=30522)
    cfg = rec.t_config(**cfgd)
    n = rec.n_players(cfg)
    torch.manual_seed(3407)
    srg = rec.t_surrogate(cfg).to(dev).eval()
    srg.agb_precision = "bf16"
    B, S = 32, 32
    ids = torch.randint(1000, 30000, (B, n + 1), device=dev)
    ids[:, 0] = 101

    def step():
        pm = ash.mask_shapley_new(B * S, n, device=dev, rng="philox", seed=1, offset=0, packed=True)
        with torch.no_grad():
            return rec.fw_surrogate(srg, ids, pm)[0]

    ms = timed(step, iters=8, warm=3)
    T, H, I, L = n + 1, cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers
    fl = L * (2 * T * H * 3 * H + 2 * T * H * H + 4 * T * H * I + 4 * T * T * H) + 2 * H * H
    print(f"bert_base T={T}: {B * S} masked evals in {ms:7.2f} ms  {B * S / ms * 1e3:9.0f} evals/s  "
          f"{B * S * fl / ms * 1e-9:7.1f} TFLOP/s dense-equivalent ({fl * 1e-9:.2f} GFLOP/eval; masked tokens are dropped, "
          f"so about half of it is actually executed)")


def surrogate_training():
    import bench
    from autognothi_b200.recipes.vanilla_vit import vanilla_vit_recipe
    dev = torch.device("cuda:0")
    rec = vanilla_vit_recipe()
    cfgd = dict(bench.VIT_BASE)
    cfg = rec.t_config(**cfgd)
    n = rec.n_players(cfg)
    torch.manual_seed(3407)
    cls = rec.t_classifier(cfg).to(dev).eval()
    srg = rec.t_surrogate(cfg).to(dev).train()
    cls.agb_precision = srg.agb_precision = "bf16"
    opt = torch.optim.AdamW(srg.parameters(), lr=1e-5, fused=True)
    B = 128
    xs = torch.randn(B, 3, 224, 224, device=dev)
    ones = ash.PackedMasks.ones(B, n, dev)
    state = {"i": 0}

    def step():
        masks = ash.mask_purely_uniform(B, n, device=dev, rng="philox", seed=3, offset=state["i"] * B, packed=True)
        state["i"] += 1
        with torch.no_grad():
            _, orig = rec.fw_classifier(cls, xs, ones)
        opt.zero_grad(set_to_none=True)
        adapt, _ = rec.fw_surrogate(srg, xs, masks)
        loss = ash.loss_logits_kl_divergence(orig, adapt)
        loss.backward()
        opt.step()
        return loss

    ms = timed(step, iters=5, warm=2)
    if "profile" in sys.argv:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                step()
            torch.cuda.synchronize()
        rows = [(k.key, k.device_time_total / 3e3, k.count // 3) for k in prof.key_averages() if k.device_time_total > 0]
        rows.sort(key=lambda r: -r[1])
        for k, t, c in rows[:28]:
            print(f"  {t:8.3f} ms  x{c:<4d} {k[:120]}")
    fl = 4 * bench.flops_per_eval(cfgd)     # teacher forward + student forward + backward (2x)
    print(f"surrogate training ViT-B/16: {B} samples in {ms:7.2f} ms  {B / ms * 1e3:8.0f} samples/s  "
          f"{B * fl / ms * 1e-9:7.1f} TFLOP/s (4 x 35.1 GFLOP per sample)")


def ltt():
    """LTT (ladder side tuning): masked surrogate evals/s through the side ladder and side-ladder explainer training."""
    import bench
    from autognothi_b200.recipes.ltt_bert import ltt_bert_recipe
    from autognothi_b200.recipes.ltt_vit import ltt_vit_recipe
    dev = torch.device("cuda:0")
    vit_cfg = {k: v for k, v in bench.VIT_BASE.items() if k not in ("explainer_attn_num_layers", "explainer_head_hidden_size")}
    # ladder ratios of the reference's only LTT configuration (hidden / 8, MLP 4 x): the same as bench.py's `ltt` leg
    vit_cfg.update(explainer_s_attn_num_layers=1, explainer_s_head_hidden_size=3072, s_attn_hidden_size=96, s_attn_intermediate_size=384)
    # reference experiments/bert_base_tayp_ltt/.hparams.json:14-32 with max_position_embeddings = 128
    bert_cfg = dict(attention_probs_dropout_prob=0.1, explainer_s_attn_num_layers=1, explainer_s_head_hidden_size=3072,
                    explainer_normalize=True, hidden_dropout_prob=0.1, hidden_size=768, intermediate_size=3072, layer_norm_eps=1e-12,
                    max_position_embeddings=128, num_attention_heads=12, num_hidden_layers=12, num_labels=2, pad_token_id=0,
                    s_attn_hidden_size=96, s_attn_intermediate_size=384, type_vocab_size=2, vocab_size=30522)
    for tag, rec, cfgd in (("ViT-B/16 + ladder 96", ltt_vit_recipe(), vit_cfg), ("BERT-base T=128 + ladder 96", ltt_bert_recipe(), bert_cfg)):
        cfg = rec.t_config(**cfgd)
        n = rec.n_players(cfg)
        torch.manual_seed(3407)
        srg = rec.t_surrogate(cfg).to(dev).eval()
        srg.agb_precision = "bf16"
        B, S = 32, 32
        if "ViT" in tag:
            xs = torch.randn(B, 3, 224, 224, device=dev)
        else:
            xs = torch.randint(1000, 30000, (B, n + 1), device=dev)
            xs[:, 0] = 101
        ones = ash.PackedMasks.ones(B, n, dev)

        def step_eval():
            pm = ash.mask_shapley_new(B * S, n, device=dev, rng="philox", seed=1, offset=0, packed=True)
            with torch.no_grad():
                return rec.fw_surrogate(srg, xs, pm)[0]

        ms = timed(step_eval, iters=6, warm=3)
        print(f"ltt {tag}: {B * S} masked evals (side + backbone heads) in {ms:7.2f} ms  {B * S / ms * 1e3:9.0f} evals/s")
        exp = rec.conv_surrogate_explainer(cfg, None, srg).train()
        exp.agb_precision = "bf16"
        opt = torch.optim.AdamW([p for p in exp.parameters() if p.requires_grad], lr=1e-5, fused=True)
        with torch.no_grad():
            null = rec.fw_surrogate(srg, rec.gen_null(cfg, type("M", (), {"tokenizer": None})(), dev), ash.PackedMasks.ones(1, n, dev))[0]
        state = {"i": 0}

        def step_train():
            pm = ash.mask_shapley_new(B * S, n, device=dev, rng="philox", seed=2, offset=state["i"] * B * S, packed=True)
            state["i"] += 1
            with torch.no_grad():
                v_s = rec.fw_surrogate(srg, xs, pm)[0]
                grand = rec.fw_surrogate(srg, xs, ones)[0]
            opt.zero_grad(set_to_none=True)
            phi, _ = rec.fw_explainer(exp, xs, ones, grand, null)
            loss = ash.loss_shapley_new(B, S, n, pm, null, v_s, grand, phi)
            loss.backward()
            opt.step()
            return loss

        ms = timed(step_train, iters=4, warm=2)
        print(f"ltt {tag}: explainer (side ladder) training step, {B} samples x {S} coalitions in {ms:7.2f} ms  {B / ms * 1e3:8.0f} samples/s")
        if "profile" in sys.argv:
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                step_eval()
                torch.cuda.synchronize()
            rows = [(k.key, k.device_time_total / 1e3, k.count) for k in prof.key_averages() if k.device_time_total > 0]
            rows.sort(key=lambda r: -r[1])
            for k, t, c in rows[:14]:
                print(f"  {t:8.3f} ms  x{c:<4d} {k[:110]}")


if __name__ == "__main__":
    which = sys.argv[1:] or ["kernelshap", "bert", "surrogate", "ltt"]
    if "kernelshap" in which:
        kernelshap()
    if "bert" in which:
        bert_eval()
    if "surrogate" in which:
        surrogate_training()
    if "ltt" in which:
        ltt()
