"""autognothi_b200 — B200 (sm_100a) native implementation of AutoGnothi's coalition-masked
surrogate / explainer evaluation hot path.  Importing the package loads the CUDA C-ABI library and
raises if it is missing: there is no CPU fallback."""
from . import _native  # noqa: F401  (fails loudly when the extension is not built)

__version__ = "0.1.0"
