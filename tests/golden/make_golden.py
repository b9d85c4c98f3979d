"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference.

Run in the development container only (needs /root/reference):
    python tests/golden/make_golden.py
The reference is imported as a package from its parent directory (as reference main.py:7-9 does).
Weights and inputs are oracle.synth tensors (a pure function of names/seeds, so nothing large is
stored); outputs of the reference's own recipes/*.py and models/shapley.py are saved as small .npz
files.  Each fixture records the seeds needed to rebuild its inputs.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_PARENT = os.environ.get("AGB_REFERENCE_PARENT", "/root")
REF_NAME = os.environ.get("AGB_REFERENCE_NAME", "reference")
sys.path.insert(0, REF_PARENT)

from oracle import configs as ocfg  # noqa: E402
from oracle import synth  # noqa: E402

ref_shapley = __import__(f"{REF_NAME}.models.shapley", fromlist=["x"])
ref_vit = __import__(f"{REF_NAME}.models.vanilla_vit", fromlist=["x"])
ref_bert = __import__(f"{REF_NAME}.models.vanilla_bert", fromlist=["x"])
ref_rvit = __import__(f"{REF_NAME}.recipes.vanilla_vit", fromlist=["x"])
ref_rbert = __import__(f"{REF_NAME}.recipes.vanilla_bert", fromlist=["x"])

torch.set_grad_enabled(False)


def to_torch_state(sd):
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}


def sample_masks_with_uniforms(seed: int, n_rows: int, n_players: int):
    """Run the reference sampler under `seed` and replay the same generator to capture the uniforms
    it consumed (reference models/shapley.py:69 draws (P,n) first, then l.133 draws (P,1))."""
    torch.manual_seed(seed)
    masks = ref_shapley.mask_shapley_new(n_rows, n_players)
    torch.manual_seed(seed)
    u_players = torch.rand(n_rows // 2, n_players)
    u_size = torch.rand((n_rows // 2, 1)).reshape(-1)
    # prefix table exactly as the reference builds it (l.65-67, 132)
    probs = torch.arange(1, n_players) * (n_players - torch.arange(1, n_players))
    probs = 1 / probs
    probs = probs / probs.sum()
    prefix = torch.cumsum(probs, dim=0) - probs
    return masks.numpy(), u_players.numpy(), u_size.numpy(), prefix.numpy()


def gen_sampler():
    out = {}
    for n_players, rows, seed in [(196, 64, 3407), (127, 32, 11), (511, 16, 5), (16, 64, 7), (2, 8, 1)]:
        masks, u_p, u_s, prefix = sample_masks_with_uniforms(seed, rows, n_players)
        key = f"n{n_players}"
        out[key + "_masks"] = masks.astype(np.int8)
        out[key + "_u_players"] = u_p
        out[key + "_u_size"] = u_s
        out[key + "_prefix"] = prefix
        out[key + "_seed"] = np.array([seed, rows], dtype=np.int64)
    # mask_purely_uniform (reference models/shapley.py:109-115)
    torch.manual_seed(99)
    pu = ref_shapley.mask_purely_uniform(16, 196)
    torch.manual_seed(99)
    a = torch.rand((16, 196))
    b = torch.rand((16, 1))
    out["pu_masks"] = pu.numpy().astype(np.int8)
    out["pu_u_players"] = a.numpy()
    out["pu_u_row"] = b.numpy().reshape(-1)
    np.savez_compressed(os.path.join(HERE, "sampler.npz"), **out)
    print("sampler.npz", {k: v.shape for k, v in out.items() if k.endswith("_masks")})


def gen_shapley_math():
    """normalize_shapley_explanation + loss_shapley_new (+ autograd dphi) on random tensors."""
    out = {}
    g = torch.Generator().manual_seed(1234)
    for tag, (B, S, n, C) in {"vit": (3, 8, 196, 10), "bert": (2, 4, 127, 2), "tiny": (1, 2, 5, 3)}.items():
        T = n + 1
        pred = torch.randn(B, T, C, generator=g)
        grand = torch.rand(B, C, generator=g)
        null = torch.rand(1, C, generator=g)
        norm = ref_shapley.normalize_shapley_explanation(pred, grand, null)
        phi = norm[:, 1:, :].permute(0, 2, 1).contiguous()
        torch.manual_seed(77)
        mask = ref_shapley.mask_shapley_new(B * S, n).reshape(B, S, n)
        v_s = torch.rand(B * S, C, generator=g)
        with torch.enable_grad():
            phi_g = phi.clone().requires_grad_(True)
            loss = ref_shapley.loss_shapley_new(B, S, n, mask, null, v_s, grand, phi_g)
            loss.backward()
        out.update({
            f"{tag}_pred": pred.numpy(), f"{tag}_grand": grand.numpy(), f"{tag}_null": null.numpy(),
            f"{tag}_norm": norm.numpy(), f"{tag}_phi": phi.numpy(), f"{tag}_mask": mask.numpy().astype(np.int8),
            f"{tag}_v_s": v_s.numpy(), f"{tag}_loss": loss.detach().numpy(), f"{tag}_dphi": phi_g.grad.numpy(),
        })
    np.savez_compressed(os.path.join(HERE, "shapley_math.npz"), **out)
    print("shapley_math.npz ok")


MODEL_CASES = [
    # name, rows B (inputs), coalitions per input S, seeds
    ("vit_mini", 2, 4),
    ("vit_mini_px64", 3, 4),
    ("vit_tiny", 2, 4),
    ("vit_base", 1, 2),
    ("vit_base_b4s32", 4, 32),     # the bench's kernel variants (>= 128 rows): see oracle/configs.py
    ("vit_large", 1, 2),
    ("bert_mini", 3, 4),
    ("bert_base_128", 1, 2),
    ("bert_mini_512", 2, 2),
]


def gen_models(only=None):
    keys = {}
    keys_path = os.path.join(HERE, "state_dict_keys.json")
    if only and os.path.exists(keys_path):
        with open(keys_path) as f:
            keys = json.load(f)
    for name, B, S in MODEL_CASES:
        if only and name not in only:
            continue
        cfg = ocfg.get_config(name)
        vit = ocfg.is_vit(cfg)
        n = ocfg.n_players(cfg)
        if vit:
            rcfg = ref_vit.VanillaViTConfig(**cfg)
            srg, exp, rec = ref_vit.VanillaViTSurrogate(rcfg), ref_vit.VanillaViTExplainer(rcfg), ref_rvit
        else:
            rcfg = ref_bert.VanillaBertConfig(**cfg)
            srg, exp, rec = ref_bert.VanillaBertSurrogate(rcfg), ref_bert.VanillaBertExplainer(rcfg), ref_rbert
        keys[name] = {
            "surrogate": {k: list(v.shape) for k, v in srg.state_dict().items()},
            "explainer": {k: list(v.shape) for k, v in exp.state_dict().items()},
        }
        srg.load_state_dict(to_torch_state(synth.surrogate_state(cfg, seed=0)), strict=True)
        exp.load_state_dict(to_torch_state(synth.explainer_state(cfg, seed=1)), strict=True)
        srg.eval()
        exp.eval()
        xs = torch.from_numpy(synth.inputs(cfg, B, seed=0))
        torch.manual_seed(3407)
        masks = ref_shapley.mask_shapley_new(B * S, n)                      # (B*S, n) int64
        xs_ext = xs.repeat_interleave(S, dim=0)                              # row b*S+s (train_explainer.py:159-163)
        ones = torch.ones((B, n), dtype=torch.long)
        null_x = rec._gen_null(cfg["img_px_size"], cfg["img_patch_size"], torch.device("cpu")) if vit else \
            torch.from_numpy(__import__("oracle.transformer", fromlist=["x"]).null_input(cfg))
        v_s, _ = rec._fw_surrogate(srg, xs_ext, masks)
        grand, _ = rec._fw_surrogate(srg, xs, ones)
        null, _ = rec._fw_surrogate(srg, null_x, torch.ones((1, n), dtype=torch.long))
        phi, _ = rec._fw_explainer(exp, xs, ones, grand, null)
        # a coalition-masked explainer call too (exercises the mask inside the explainer layers)
        phi_masked, _ = rec._fw_explainer(exp, xs, masks.reshape(B, S, n)[:, 0, :], grand, null)
        loss = ref_shapley.loss_shapley_new(B, S, n, masks.reshape(B, S, n), null, v_s, grand, phi)
        np.savez_compressed(
            os.path.join(HERE, f"model_{name}.npz"),
            masks=masks.numpy().astype(np.int8), v_s=v_s.numpy(), grand=grand.numpy(), null=null.numpy(),
            phi=phi.numpy(), phi_masked=phi_masked.numpy(), loss=loss.numpy(),
            meta=np.array([B, S, n], dtype=np.int64),
        )
        print(f"model_{name}.npz  v_s{tuple(v_s.shape)} phi{tuple(phi.shape)} loss={float(loss):.6f}")
    with open(os.path.join(HERE, "state_dict_keys.json"), "w") as f:
        json.dump(keys, f, indent=0, sort_keys=True)


def gen_train_grads(only=None):
    """One explainer training step's loss and parameter gradients from the reference's autograd
    (scripts/train_explainer.py:182-197 semantics), in eval() mode so that dropout is the identity."""
    for name, B, S in [("vit_mini", 2, 4), ("vit_mini_px64", 3, 4), ("bert_mini", 3, 4), ("bert_mini_512", 2, 2),
                       ("vit_base_b4s32", 4, 32)]:
        if only and name not in only:
            continue
        if not only and name == "vit_base_b4s32":
            continue        # minutes of CPU time: generated on request (`make_golden.py train vit_base_b4s32`)
        cfg = ocfg.get_config(name)
        vit = ocfg.is_vit(cfg)
        n = ocfg.n_players(cfg)
        if vit:
            rcfg = ref_vit.VanillaViTConfig(**cfg)
            srg, exp, rec = ref_vit.VanillaViTSurrogate(rcfg), ref_vit.VanillaViTExplainer(rcfg), ref_rvit
        else:
            rcfg = ref_bert.VanillaBertConfig(**cfg)
            srg, exp, rec = ref_bert.VanillaBertSurrogate(rcfg), ref_bert.VanillaBertExplainer(rcfg), ref_rbert
        srg.load_state_dict(to_torch_state(synth.surrogate_state(cfg, seed=0)), strict=True)
        exp.load_state_dict(to_torch_state(synth.explainer_state(cfg, seed=1)), strict=True)
        srg.eval(); exp.eval()
        g = np.load(os.path.join(HERE, f"model_{name}.npz"))
        masks = torch.from_numpy(g["masks"].astype(np.int64))
        xs = torch.from_numpy(synth.inputs(cfg, B, seed=0))
        v_s, grand, null = (torch.from_numpy(g[k]) for k in ("v_s", "grand", "null"))
        ones = torch.ones((B, n), dtype=torch.long)
        with torch.enable_grad():
            for p_ in exp.parameters():
                p_.requires_grad_(True)
            phi, _ = rec._fw_explainer(exp, xs, ones, grand, null)
            loss = ref_shapley.loss_shapley_new(B, S, n, masks.reshape(B, S, n), null, v_s, grand, phi)
            loss.backward()
        out = {"loss": loss.detach().numpy()}
        norms = {}
        for k, p_ in exp.named_parameters():
            gr = p_.grad
            norms[k] = float(gr.norm()) if gr is not None else 0.0
            if gr is not None and (gr.numel() <= 1024 or ((("layers.0.attention.self.query.weight" in k
                                   or "explainer_attn.0.output.dense.weight" in k)) and gr.numel() <= 65536)):
                out["grad::" + k] = gr.numpy()
            elif gr is not None and name == "vit_base_b4s32":
                # full-size model: every large tensor is pinned by a strided sample of <= 2048 elements (relative L2 on the
                # sample estimates the tensor's), so that the fixture stays small
                flat = gr.reshape(-1)
                step = max(1, flat.numel() // 2048)
                out["sample::" + k] = flat[::step][:2048].numpy().copy()
        out["norm_names"] = np.array(list(norms.keys()))
        out["norm_values"] = np.array(list(norms.values()), dtype=np.float64)
        np.savez_compressed(os.path.join(HERE, f"train_{name}.npz"), **out)
        print(f"train_{name}.npz loss={float(loss):.6f} params={len(norms)} stored={len(out) - 3}")


def gen_surrogate_train_grads():
    """One surrogate training step (scripts/train_surrogate.py:131-150): classifier teacher on the full input, masked
    surrogate student, loss_logits_kl_divergence, parameter gradients from the reference's autograd; eval() mode so
    that dropout is the identity."""
    for name, B, S in [("vit_mini", 2, 4), ("bert_mini", 3, 4)]:
        cfg = ocfg.get_config(name)
        vit = ocfg.is_vit(cfg)
        n = ocfg.n_players(cfg)
        if vit:
            rcfg = ref_vit.VanillaViTConfig(**cfg)
            cls, srg, rec = ref_vit.VanillaViTClassifier(rcfg), ref_vit.VanillaViTSurrogate(rcfg), ref_rvit
        else:
            rcfg = ref_bert.VanillaBertConfig(**cfg)
            cls, srg, rec = ref_bert.VanillaBertClassifier(rcfg), ref_bert.VanillaBertSurrogate(rcfg), ref_rbert
        cls.load_state_dict(to_torch_state(synth.surrogate_state(cfg, seed=5)), strict=True)
        srg.load_state_dict(to_torch_state(synth.surrogate_state(cfg, seed=0)), strict=True)
        cls.eval(); srg.eval()
        g = np.load(os.path.join(HERE, f"model_{name}.npz"))
        masks = torch.from_numpy(g["masks"].astype(np.int64)).reshape(B, S, n)[:, 1, :].contiguous()   # one row per input
        xs = torch.from_numpy(synth.inputs(cfg, B, seed=0))
        ones = torch.ones((B, n), dtype=torch.long)
        with torch.no_grad():
            _, orig = rec._fw_classifier(cls, xs, ones)
        with torch.enable_grad():
            for p_ in srg.parameters():
                p_.requires_grad_(True)
            adapt, _ = rec._fw_surrogate(srg, xs, masks)
            loss = ref_shapley.loss_logits_kl_divergence(orig, adapt)
            loss.backward()
        out = {"loss": loss.detach().numpy(), "orig": orig.numpy(), "adapt": adapt.detach().numpy(),
               "masks": masks.numpy().astype(np.int8)}
        norms = {}
        for k, p_ in srg.named_parameters():
            gr = p_.grad
            norms[k] = float(gr.norm()) if gr is not None else 0.0
            if gr is not None and (gr.numel() <= 1024 or "layers.0.attention.self.key.weight" in k
                                   or "layers.1.output.dense.weight" in k):
                out["grad::" + k] = gr.numpy()
        out["norm_names"] = np.array(list(norms.keys()))
        out["norm_values"] = np.array(list(norms.values()), dtype=np.float64)
        np.savez_compressed(os.path.join(HERE, f"train_surrogate_{name}.npz"), **out)
        print(f"train_surrogate_{name}.npz loss={float(loss):.6e} params={len(norms)} stored={len(out) - 6}")


def gen_froyo():
    """Froyo bundles (reference models/froyo_vit.py:100-171, froyo_bert.py:105-204): one backbone pass, classifier + surrogate
    heads, explainer tail.  Also the key/shape tables of every Froyo class."""
    ref_fvit = __import__(f"{REF_NAME}.models.froyo_vit", fromlist=["x"])
    ref_fbert = __import__(f"{REF_NAME}.models.froyo_bert", fromlist=["x"])
    keys = {}
    for name, B in [("vit_mini", 3), ("bert_mini", 3), ("vit_tiny", 2)]:
        cfg = ocfg.get_config(name)
        vit = ocfg.is_vit(cfg)
        n = ocfg.n_players(cfg)
        mod = ref_fvit if vit else ref_fbert
        fcfg = (mod.FroyoViTConfig if vit else mod.FroyoBertConfig)(**cfg)
        final = (mod.FroyoViTFinal if vit else mod.FroyoBertFinal)(fcfg)
        classes = ("FroyoViT" if vit else "FroyoBert")
        keys[name] = {"final": {k: list(v.shape) for k, v in final.state_dict().items()}}
        for role in ("Classifier", "Surrogate", "Explainer"):
            m = getattr(mod, classes + role)(fcfg)
            m.train()
            keys[name][role.lower()] = {k: list(v.shape) for k, v in m.state_dict().items()}
            keys[name][role.lower() + "_trainable"] = sorted(k for k, p in m.named_parameters() if p.requires_grad)
        final.load_state_dict(to_torch_state(synth.froyo_final_state(cfg, seed=1)), strict=True)
        final.eval()
        xs = torch.from_numpy(synth.inputs(cfg, B, seed=0))
        torch.manual_seed(3407)
        masks = ref_shapley.mask_shapley_new(2 * B, n)[:B]
        out = {}
        for tag, m in (("ones", torch.ones((B, n), dtype=torch.long)), ("masked", masks)):
            tok = torch.cat([torch.ones((B, 1), dtype=torch.long), m], dim=1)     # recipes' _fw_xs_preprocess: CLS column
            if vit:
                cls, phi = final(xs, tok, None, None)
            else:
                cls, phi = final(xs, tok, torch.zeros_like(xs))
            out[f"{tag}_cls"], out[f"{tag}_phi"] = cls.numpy(), phi.numpy()
        out["masks"] = masks.numpy().astype(np.int8)
        np.savez_compressed(os.path.join(HERE, f"froyo_{name}.npz"), **out)
        print(f"froyo_{name}.npz cls{tuple(out['ones_cls'].shape)} phi{tuple(out['ones_phi'].shape)}")
    with open(os.path.join(HERE, "froyo_keys.json"), "w") as f:
        json.dump(keys, f, indent=0, sort_keys=True)


def gen_ltt():
    """LTT variants (reference models/ltt_{vit,bert}.py): surrogate (side + backbone probabilities) on coalition masks,
    explainer, bundle; side-ladder training gradients from the reference's autograd (eval() mode: dropout = identity);
    key/shape tables and the trainable sets of every class."""
    ref_lvit = __import__(f"{REF_NAME}.models.ltt_vit", fromlist=["x"])
    ref_lbert = __import__(f"{REF_NAME}.models.ltt_bert", fromlist=["x"])
    keys = {}
    for name, B, S in [("ltt_vit_mini", 2, 4), ("ltt_bert_mini", 3, 4), ("ltt_vit_tiny", 1, 2), ("ltt_bert_base_128", 1, 2)]:
        cfg = ocfg.get_config(name)
        vit = ocfg.is_vit(cfg)
        n = ocfg.n_players(cfg)
        mod = ref_lvit if vit else ref_lbert
        pre = "LttViT" if vit else "LttBert"
        lcfg = getattr(mod, pre + "Config")(**cfg)
        models = {r: getattr(mod, pre + r)(lcfg) for r in ("Surrogate", "Explainer", "Final")}
        keys[name] = {}
        for r, m in models.items():
            m.train()
            keys[name][r.lower()] = {k: list(v.shape) for k, v in m.state_dict().items()}
            keys[name][r.lower() + "_trainable"] = sorted(k for k, p in m.named_parameters() if p.requires_grad)
        for i, (r, m) in enumerate(models.items()):
            m.load_state_dict(to_torch_state(synth.state_like({k: v.shape for k, v in m.state_dict().items()}, seed=30 + i)), strict=True)
            m.eval()
        srg, exp, fin = models["Surrogate"], models["Explainer"], models["Final"]
        xs = torch.from_numpy(synth.inputs(cfg, B, seed=0))
        torch.manual_seed(3407)
        masks = ref_shapley.mask_shapley_new(B * S, n)
        cls_col = lambda m: torch.cat([torch.ones((m.shape[0], 1), dtype=torch.long), m], dim=1)   # noqa: E731
        xs_ext = xs.repeat_interleave(S, dim=0)
        ones = torch.ones((B, n), dtype=torch.long)
        tt = (lambda x: ()) if vit else (lambda x: (torch.zeros_like(x),))
        v_side, v_main = srg(xs_ext, cls_col(masks), *tt(xs_ext))
        grand, _ = srg(xs, cls_col(ones), *tt(xs))
        null = torch.from_numpy(synth.state_like({"surrogate_null": (1, cfg["num_labels"])}, seed=7)["surrogate_null"])
        phi, e_main = exp(xs, cls_col(ones), *tt(xs), grand, null)
        phi_m, _ = exp(xs, cls_col(masks.reshape(B, S, n)[:, 0, :]), *tt(xs), grand, null)
        f_cls, f_phi = fin(xs, cls_col(ones), *tt(xs))
        out = dict(masks=masks.numpy().astype(np.int8), v_side=v_side.numpy(), v_main=v_main.numpy(), grand=grand.numpy(),
                   null=null.numpy(), phi=phi.numpy(), e_main=e_main.numpy(), phi_masked=phi_m.numpy(), f_cls=f_cls.numpy(),
                   f_phi=f_phi.numpy(), meta=np.array([B, S, n], dtype=np.int64))
        if "mini" in name:
            # explainer training step: only the side ladder + side explainer receive gradients (train() freezes the rest)
            exp.train()
            exp.eval()      # keep the freezes, drop the dropout
            for k, p_ in exp.named_parameters():
                p_.requires_grad_(k in keys[name]["explainer_trainable"])
            with torch.enable_grad():
                phi_g, _ = exp(xs, cls_col(ones), *tt(xs), grand, null)
                loss = ref_shapley.loss_shapley_new(B, S, n, masks.reshape(B, S, n), null, v_side, grand, phi_g)
                loss.backward()
            out["train_loss"] = loss.detach().numpy()
            names, vals = [], []
            for k, p_ in exp.named_parameters():
                if p_.grad is None:
                    continue
                names.append(k); vals.append(float(p_.grad.norm()))
                if p_.grad.numel() <= 4096:
                    out["grad::" + k] = p_.grad.numpy()
            out["norm_names"], out["norm_values"] = np.array(names), np.array(vals, dtype=np.float64)
            # surrogate training step (KL to the frozen backbone's own prediction, scripts/train_surrogate.py:131-150)
            srg.train(); srg.eval()
            for k, p_ in srg.named_parameters():
                p_.requires_grad_(k in keys[name]["surrogate_trainable"])
            m1 = masks.reshape(B, S, n)[:, 1, :].contiguous()
            # teacher distribution: a fixed synthetic one, far from the student, so that the KL and its gradients are
            # well conditioned (the two ladders' outputs on nearby masks differ by ~1e-3, which makes the KL ~1e-6)
            target = torch.softmax(3.0 * torch.from_numpy(synth.uniform("ltt.target", (B, cfg["num_labels"]), 5)), dim=-1)
            with torch.enable_grad():
                side, main = srg(xs, cls_col(m1), *tt(xs))
                sloss = ref_shapley.loss_logits_kl_divergence(target, side)
                sloss.backward()
            out["srg_loss"], out["srg_side"], out["srg_target"] = sloss.detach().numpy(), side.detach().numpy(), target.numpy()
            names, vals = [], []
            for k, p_ in srg.named_parameters():
                if p_.grad is None:
                    continue
                names.append(k); vals.append(float(p_.grad.norm()))
            out["srg_norm_names"], out["srg_norm_values"] = np.array(names), np.array(vals, dtype=np.float64)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(f"{name}.npz v_side{tuple(v_side.shape)} phi{tuple(phi.shape)}")
    with open(os.path.join(HERE, "ltt_keys.json"), "w") as f:
        json.dump(keys, f, indent=0, sort_keys=True)


def gen_duo():
    """Duo variants (reference models/duo_vanilla_{vit,bert}.py): explainer outputs (attributions + class output), the bundle,
    and the dual-objective training step of scripts/train_duo_explainer.py:180-196 (cross_entropy(class output, labels) +
    Shapley loss; eval() mode: dropout = identity) with gradients from the reference's autograd."""
    ref_dvit = __import__(f"{REF_NAME}.models.duo_vanilla_vit", fromlist=["x"])
    ref_dbert = __import__(f"{REF_NAME}.models.duo_vanilla_bert", fromlist=["x"])
    keys = {}
    for name, B, S in [("vit_mini", 2, 4), ("bert_mini", 3, 4)]:
        cfg = ocfg.get_config(name)
        vit = ocfg.is_vit(cfg)
        n = ocfg.n_players(cfg)
        mod = ref_dvit if vit else ref_dbert
        pre = "DuoVanillaViT" if vit else "DuoVanillaBert"
        dcfg = getattr(mod, pre + "Config")(**cfg)
        exp, fin = getattr(mod, pre + "Explainer")(dcfg), getattr(mod, pre + "Final")(dcfg)
        keys[name] = {"explainer": {k: list(v.shape) for k, v in exp.state_dict().items()},
                      "final": {k: list(v.shape) for k, v in fin.state_dict().items()}}
        for i, m in enumerate((exp, fin)):
            m.load_state_dict(to_torch_state(synth.state_like({k: v.shape for k, v in m.state_dict().items()}, seed=50 + i)), strict=True)
            m.eval()
        g = np.load(os.path.join(HERE, f"model_{name}.npz"))
        masks = torch.from_numpy(g["masks"].astype(np.int64))
        v_s, grand, null = (torch.from_numpy(g[k]) for k in ("v_s", "grand", "null"))
        xs = torch.from_numpy(synth.inputs(cfg, B, seed=0))
        cls_col = lambda m: torch.cat([torch.ones((m.shape[0], 1), dtype=torch.long), m], dim=1)   # noqa: E731
        ones = torch.ones((B, n), dtype=torch.long)
        tt = (lambda x: ()) if vit else (lambda x: (torch.zeros_like(x),))
        labels = torch.from_numpy(synth.randint("duo.labels", (B,), 0, cfg["num_labels"], 3))

        def run(m_):
            o = exp(xs, cls_col(m_), *tt(xs), grand, null)
            return (o[0], o[1]) if vit else (o[1], o[0])          # -> (phi, class output)

        phi, cls = run(ones)
        phi_m, cls_m = run(masks.reshape(B, S, n)[:, 0, :])
        f_cls, f_phi = fin(xs, cls_col(ones), *tt(xs))
        with torch.enable_grad():
            for p_ in exp.parameters():
                p_.requires_grad_(True)
            phi_g, cls_g = run(ones)
            loss_cls = torch.nn.functional.cross_entropy(cls_g, labels)
            loss_shap = ref_shapley.loss_shapley_new(B, S, n, masks.reshape(B, S, n), null, v_s, grand, phi_g)
            (loss_cls + loss_shap).backward()
        out = dict(phi=phi.numpy(), cls=cls.numpy(), phi_masked=phi_m.numpy(), cls_masked=cls_m.numpy(), f_cls=f_cls.numpy(),
                   f_phi=f_phi.numpy(), labels=labels.numpy(), loss_cls=loss_cls.detach().numpy(),
                   loss_shap=loss_shap.detach().numpy(), meta=np.array([B, S, n], dtype=np.int64))
        names, vals = [], []
        for k, p_ in exp.named_parameters():
            names.append(k); vals.append(float(p_.grad.norm()) if p_.grad is not None else 0.0)
            if p_.grad is not None and (p_.grad.numel() <= 1024 or k.startswith("classifier.") or "layers.0.attention.self.query.weight" in k):
                out["grad::" + k] = p_.grad.numpy()
        out["norm_names"], out["norm_values"] = np.array(names), np.array(vals, dtype=np.float64)
        np.savez_compressed(os.path.join(HERE, f"duo_{name}.npz"), **out)
        print(f"duo_{name}.npz phi{tuple(phi.shape)} cls{tuple(cls.shape)} loss_cls={float(loss_cls):.5f} loss_shap={float(loss_shap):.5f}")
    with open(os.path.join(HERE, "duo_keys.json"), "w") as f:
        json.dump(keys, f, indent=0, sort_keys=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "duo":
        gen_duo()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "ltt":
        gen_ltt()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "froyo":
        gen_froyo()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "train":
        gen_train_grads(only=sys.argv[2:] or None)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "model":
        gen_models(only=sys.argv[2:] or None)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "surrogate":
        gen_surrogate_train_grads()
        sys.exit(0)
    gen_sampler()
    gen_shapley_math()
    gen_models()
    gen_train_grads()
    gen_surrogate_train_grads()
    gen_froyo()
    gen_ltt()
    gen_duo()
