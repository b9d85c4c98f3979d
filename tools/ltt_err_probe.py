import sys, os, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import test_gpu_ltt as t
from oracle import synth
DEV = t.DEV
for name in ("ltt_bert_base_128", "ltt_vit_tiny"):
    g = np.load(f"/root/repo/tests/golden/{name}.npz")
    B, S, n = (int(v) for v in g["meta"])
    rec, cfgd, cfg, (srg, exp, fin) = t._models(name, "bf16")
    xs = torch.from_numpy(synth.inputs(cfgd, B, seed=0)).to(DEV)
    masks = torch.from_numpy(g["masks"].astype(np.int64)).to(DEV)
    ones = torch.ones((B, n), dtype=torch.int64, device=DEV)
    grand, null = torch.from_numpy(g["grand"]).to(DEV), torch.from_numpy(g["null"]).to(DEV)
    with torch.no_grad():
        phi, _ = rec.fw_explainer(exp, xs, ones, grand, null)
        phi_m, _ = rec.fw_explainer(exp, xs, masks.reshape(B, S, n)[:, 0, :].contiguous(), grand, null)
        f_cls, f_phi = rec.fw_final(fin, xs)
    for tag, got, ref in (("phi", phi, g["phi"]), ("phi_masked", phi_m, g["phi_masked"]), ("f_phi", f_phi, g["f_phi"])):
        a, b = t._np(got).reshape(-1).astype(np.float64), ref.reshape(-1).astype(np.float64)
        print(name, tag, "pearson", float(np.corrcoef(a, b)[0, 1]), "rel-L2", float(np.linalg.norm(a - b) / np.linalg.norm(b)),
              "|ref| rms", float(np.sqrt((b * b).mean())), "mean shift", float((a - b).mean()))
