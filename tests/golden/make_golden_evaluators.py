"""Golden vectors for the evaluator mask generators, produced by the UNMODIFIED reference:
scripts/measure_faithfulness.py::_get_perturbed_samples and models/shapley.py::mask_uniform_selective.

Run in the development container only (needs /root/reference):
    python tests/golden/make_golden_evaluators.py
`shap` (third-party, absent here) is stubbed so that the reference's scripts package imports; neither function
touches it.
"""
from __future__ import annotations

import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PARENT = os.environ.get("AGB_REFERENCE_PARENT", "/root")
REF_NAME = os.environ.get("AGB_REFERENCE_NAME", "reference")
sys.path.insert(0, REF_PARENT)
sys.modules.setdefault("shap", types.ModuleType("shap"))

mf = __import__(f"{REF_NAME}.scripts.measure_faithfulness", fromlist=["x"])
ref_shapley = __import__(f"{REF_NAME}.models.shapley", fromlist=["x"])


def main():
    out = {}
    rng = np.random.RandomState(7)
    cases = [(196, 50, 1), (196, 300, 0), (127, 33, 1), (16, 7, 0), (5, 4, 1), (511, 64, 0)]
    out["perturb_cases"] = np.array(cases, dtype=np.int64)
    for idx, (n, steps, base) in enumerate(cases):
        attr = rng.randn(n).astype(np.float32)
        if n <= 16:                      # insertion-sort regime of numpy's argsort: ties are stable there
            attr[1] = attr[n - 1]
            attr[2] = attr[3]
        stops, masks = mf._get_perturbed_samples(torch.from_numpy(attr), n, steps, base)
        out[f"perturb_{idx}_attr"] = attr
        out[f"perturb_{idx}_stops"] = stops.numpy()
        out[f"perturb_{idx}_masks"] = masks.numpy().astype(np.int8)
    sel = [(6, 196, 50), (3, 127, 0), (2, 16, 16), (4, 511, 500)]
    out["selective_cases"] = np.array(sel, dtype=np.int64)
    random.seed(1234)
    for idx, (b, n, k) in enumerate(sel):
        out[f"selective_{idx}"] = ref_shapley.mask_uniform_selective(b, n, k).numpy().astype(np.int8)
    np.savez_compressed(os.path.join(HERE, "evaluators.npz"), **out)
    print("wrote evaluators.npz:", {k: v.shape for k, v in out.items() if "cases" in k})


if __name__ == "__main__":
    main()
