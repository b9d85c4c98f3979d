"""CPU oracle for the coalition-masked evaluation hot path — TEST INFRASTRUCTURE ONLY.

This package is a numpy restatement of the reference's algorithms (gszfwsb/AutoGnothi, mounted at
/root/reference while developing).  Every function cites the reference file:line it follows.  It is
the checker for the CUDA path, never the product: only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  `autognothi_b200/` never does.

Pinning status
  * ViT / BERT surrogate + explainer forward, mask sampler, efficiency normalisation, Shapley loss:
    PINNED — checked against outputs of the unmodified reference run in the development container
    (fixtures in tests/golden/*.npz, produced by tests/golden/make_golden.py).
  * KernelSHAP weighted least squares (oracle/kernelshap.py): PARITY UNPINNED — the arithmetic lives
    in the third-party package `shap ~= 0.44.1` (reference requirements.txt:10), which is not
    installed here and not vendored in the reference; the restatement follows the published
    KernelSHAP algorithm and is anchored on the reference call site models/kernel_shap_bert.py:170-185.
"""
