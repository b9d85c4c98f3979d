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


def test_round2_entry_points_validate_before_touching_the_device(built):
    """hi/lo residual GEMM, fused training epilogues, hi/lo gather / split: shapes and pointers are checked first"""
    lib = ctypes.CDLL(built)
    lib.agb_last_error.restype = ctypes.c_char_p
    vp, ci, cl = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong
    lib.agb_gemm_bf16_hilo.argtypes = [vp, ci, vp, ci, ci, ci, ci, vp, vp, vp, ci, vp, vp]
    assert lib.agb_gemm_bf16_hilo(None, 768, None, 768, 256, 768, 768, None, None, None, 768, None, None) == 1
    assert b"operands" in lib.agb_last_error()
    one = ctypes.c_void_p(16)          # any non-null 16-byte-aligned value: the call must fail on the shape, before any use
    assert lib.agb_gemm_bf16_hilo(one, 768, one, 768, 256, 700, 768, None, one, one, 768, one, None) == 1      # N % 256
    assert b"alignment" in lib.agb_last_error()
    lib.agb_gemm_bf16_dropout_residual.argtypes = [vp, ci, vp, ci, ci, ci, ci, vp, vp, ci, vp, ci, ctypes.c_uint64, ci, vp]
    assert lib.agb_gemm_bf16_dropout_residual(one, 768, one, 768, 256, 768, 768, None, one, 768, one, 0, 1, 1, None) == 1
    assert b"threshold" in lib.agb_last_error()
    assert lib.agb_gemm_bf16_dropout_residual(one, 768, one, 768, 256, 768, 768, None, None, 768, one, 6554, 1, 1, None) == 1
    lib.agb_gemm_bf16_gelu_dual.argtypes = [vp, ci, vp, ci, ci, ci, ci, vp, vp, vp, ci, vp]
    assert lib.agb_gemm_bf16_gelu_dual(one, 768, one, 768, 256, 3072, 768, None, None, one, 3072, None) == 1
    lib.agb_gemm_bf16_gelu_bwd.argtypes = [vp, ci, vp, ci, ci, ci, ci, vp, ci, vp, ci, vp]
    assert lib.agb_gemm_bf16_gelu_bwd(one, 768, one, 3072, 256, 3072, 768, None, 3072, one, 3072, None) == 1
    lib.agb_split_hilo.argtypes = [vp, cl, vp, vp, vp]
    assert lib.agb_split_hilo(None, 12, None, None, None) == 1 and b"multiple of 8" in lib.agb_last_error()
    assert lib.agb_split_hilo(None, 0, None, None, None) == 0
    lib.agb_gather_token_rows_hilo.argtypes = [vp, vp, ci, ci, ci, ci, vp, vp, vp]
    assert lib.agb_gather_token_rows_hilo(None, None, 6, 197, 4, 768, None, None, None) == 1      # rows % S
    assert lib.agb_gather_token_rows_hilo(None, None, 0, 197, 4, 768, None, None, None) == 0


def test_zero_arena_hands_out_disjoint_zero_slices():
    """training._ZeroArena: the small accumulate-into gradients of one backward pass are slices of one zero chunk"""
    import torch

    from autognothi_b200 import training
    arena = training._ZeroArena()
    cpu = torch.device("cpu")
    a = arena.take((768,), cpu)
    b = arena.take((3, 5), cpu)
    c = arena.take((training._ZeroArena.CHUNK + 1,), cpu)          # larger than a chunk: its own allocation
    assert a.shape == (768,) and b.shape == (3, 5) and c.numel() == training._ZeroArena.CHUNK + 1
    assert a.is_contiguous() and b.is_contiguous() and float(a.abs().sum() + b.abs().sum() + c.abs().sum()) == 0.0
    a += 1.0
    assert float(b.abs().sum()) == 0.0                             # disjoint
    assert a.data_ptr() % 256 == b.data_ptr() % 256                # 256-byte steps between slices
    first = arena.buf
    for _ in range(training._ZeroArena.CHUNK // 1024 + 2):         # exhausting the chunk opens a fresh one, zero again
        x = arena.take((1024,), cpu)
    assert arena.buf is not first and float(x.abs().sum()) == 0.0
    with training._arena_scope():
        assert training._ARENA is not None
        z = training._zeros((4,), cpu)
    assert training._ARENA is None and z.shape == (4,)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under autognothi_b200/ may import it."""
    pkg = os.path.join(ROOT, "autognothi_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
