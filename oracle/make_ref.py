"""Recipe for oracle/_ref/: the UNMODIFIED reference, importable on the GPU box.  Test / benchmark infrastructure.

The reference is pure Python (no native code, no build step), so "building" it means placing its own
`*.py` files, byte for byte, where they can be imported as the package `reference` (it uses relative imports:
reference main.py:7-9).  Output goes ONLY to oracle/_ref/reference/, which is git-ignored (never part of the history)
but not gpurun-ignored, so it travels to the GPU box with the repo snapshot, where /root/reference does not exist.
`bench.py --impl reference` and the `cpu_baseline` leg import it from there and run the reference's own
recipes/vanilla_vit.py + models/shapley.py (+ scripts/measure_train_resources.py::_explainer_batch_train) on the host
cores.  Nothing under autognothi_b200/ ever imports it.

    python oracle/make_ref.py            # /root/reference -> oracle/_ref/reference (no-op when the source is absent)
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_SRC = os.environ.get("AGB_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref", "reference")
# the directories the hot path's modules import from (relative imports inside the package); `experiments/` only for the
# checked-in hparams the bench cites, `playground/` / `assets/` are not needed
PACKAGES = ("models", "recipes", "utils", "scripts", "datasets", "params")


def make_ref(src: str = DEFAULT_SRC, dst: str = DST) -> bool:
    """-> True when oracle/_ref/reference is in place (freshly copied or already there)."""
    if not os.path.isdir(src):
        return os.path.isdir(dst)
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(dst)
    manifest = {}
    for entry in sorted(os.listdir(src)):
        p = os.path.join(src, entry)
        if os.path.isfile(p) and entry.endswith(".py"):
            shutil.copy2(p, os.path.join(dst, entry))
        elif os.path.isdir(p) and entry in PACKAGES:
            shutil.copytree(p, os.path.join(dst, entry), ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.ipynb"))
    if not os.path.exists(os.path.join(dst, "__init__.py")):
        open(os.path.join(dst, "__init__.py"), "w").close()       # namespace marker only (the checkout has none at its root)
    for root, _, files in os.walk(dst):
        for f in sorted(files):
            if f.endswith(".py"):
                full = os.path.join(root, f)
                with open(full, "rb") as fh:
                    manifest[os.path.relpath(full, dst)] = hashlib.sha256(fh.read()).hexdigest()[:16]
    with open(os.path.join(os.path.dirname(dst), "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "files": manifest}, fh, indent=0, sort_keys=True)
    return True


def import_reference():
    """Import the reference package (oracle/_ref first, the container's /root/reference second).
    -> (package name, module getter) or None when neither exists.  `shap` is absent from the image; the
    scripts.* modules import it at module scope without using it on this path, so an empty stub module stands in."""
    import importlib
    import types
    for parent, name in ((os.path.join(HERE, "_ref"), "reference"),
                         (os.path.dirname(DEFAULT_SRC), os.path.basename(DEFAULT_SRC))):
        if os.path.isdir(os.path.join(parent, name, "models")):
            if parent not in sys.path:
                sys.path.insert(0, parent)
            if "shap" not in sys.modules:
                try:
                    import shap  # noqa: F401
                except Exception:
                    sys.modules["shap"] = types.ModuleType("shap")
            return name, (lambda sub, _n=name: importlib.import_module(f"{_n}.{sub}")), parent
    return None


if __name__ == "__main__":
    ok = make_ref()
    print("oracle/_ref/reference:", "ready" if ok else "source absent, nothing copied")
