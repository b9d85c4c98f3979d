"""CPU: the C-ABI library loads without a GPU and exports every symbol include/autognothi_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "autognothi_b200", "lib", "libautognothi_b200.so")
HEADER = os.path.join(ROOT, "include", "autognothi_b200.h")


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(LIB):
        subprocess.run(["make", "-j8", LIB.replace(ROOT + "/", "")], cwd=ROOT, check=True)
    return LIB


def test_header_prototypes_parse():
    from autognothi_b200 import _native
    src = open(HEADER).read()
    declared = set(re.findall(r"\b(agb_\w+)\s*\(", re.sub(r"/\*.*?\*/", " ", src, flags=re.S)))
    assert declared == set(_native.PROTOTYPES), declared ^ set(_native.PROTOTYPES)
    assert len(declared) >= 20


def test_library_exports_every_declared_symbol(built):
    from autognothi_b200 import _native
    lib = ctypes.CDLL(built)
    for name in _native.PROTOTYPES:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert lib.agb_version() >= 100


def test_torch_extension_has_one_op_per_launching_entry_point():
    """lib/libagb_torch.so (TORCH_LIBRARY "agb", generated from the header): every C-ABI function that takes a stream is
    callable as torch.ops.agb.<name>; the remaining (configuration / diagnostics) entry points stay on ctypes.  Loads and
    resolves without a GPU."""
    import torch
    from autognothi_b200 import _native
    assert _native.BINDING == "torch" and _native.TORCH_OPS
    launching = {n for n, (r, a) in _native.PROTOTYPES.items() if r == "int" and any(pn == "stream" and "*" in pt for pt, pn in a)}
    assert set(_native.TORCH_OPS) == launching and len(launching) >= 40
    for name in launching:
        assert hasattr(torch.ops.agb, name), name
    # CPU tensors are refused by the op itself (no CPU fallback); no launch is attempted here: the ops ask at::cuda for the
    # current stream, which needs a device
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        torch.ops.agb.agb_pack_masks_i64(torch.zeros(4, 196, dtype=torch.int64), 4, 196, 1, None, 7)


def test_every_prototype_cites_the_reference():
    """each entry point's comment names the reference file:line it replaces (drop-in boundary rule)"""
    src = open(HEADER).read()
    assert src.count("reference ") >= 15
    assert "models/shapley.py" in src and "models/vanilla_vit.py" in src and "models/vanilla_bert.py" in src


def test_invalid_arguments_are_rejected_without_a_gpu(built):
    lib = ctypes.CDLL(built)
    lib.agb_last_error.restype = ctypes.c_char_p
    # null pointers / bad shapes must come back as AGB_ERR_INVALID before any CUDA call
    rc = lib.agb_pack_masks_i64(None, 4, 196, 1, None, 7, None)
    assert rc == 1 and b"null" in lib.agb_last_error()
    rc = lib.agb_pack_masks_i64(None, 4, 196, 1, None, 3, None)
    assert rc == 1
    rc = lib.agb_masked_attention_bf16(None, None, 19, 1, 600, 768, 12, 0, None, None)
    assert rc == 3  # AGB_ERR_UNSUPPORTED: T > 256 belongs to the CUDA-core kernel
    # row fan-out: rows must be whole 16-byte vectors, S positive, pointers non-null; an empty batch is a no-op
    lib.agb_repeat_rows.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    assert lib.agb_repeat_rows(None, 2, 24, 4, None, None) == 1 and b"16-byte" in lib.agb_last_error()
    assert lib.agb_repeat_rows(None, 2, 32, 0, None, None) == 1
    assert lib.agb_repeat_rows(None, 2, 32, 4, None, None) == 1 and b"null" in lib.agb_last_error()
    assert lib.agb_repeat_rows(None, 0, 32, 4, None, None) == 0
    # narrow-head attention (side ladders) validates its shape before touching the device
    assert lib.agb_masked_attention_simt(None, 1, None, 1, 2, 40, 96, 12, 0, None, None) == 1    # 1 mask word < 40 tokens


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under autognothi_b200/ may import it."""
    pkg = os.path.join(ROOT, "autognothi_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
