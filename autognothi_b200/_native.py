"""Binding of libautognothi_b200.so (the C-ABI declared in include/autognothi_b200.h).

Kernel launches go through a thin torch C++ extension, lib/libagb_torch.so: one TORCH_LIBRARY op per C-ABI entry point
(generated from the header by tools/gen_torch_binding.py) that takes `Tensor?` arguments, checks that they live on the
current CUDA device, launches on at::cuda's current stream and forwards to the `extern "C"` function of the same name.
Configuration / diagnostics entry points (no stream argument) and `AGB_BINDING=ctypes` use ctypes on the same library;
its prototypes are parsed from the header itself, so Python can never drift from the C declarations.
There is NO fallback: if a shared library is missing or a symbol is absent, importing this module raises — the
product path must fail loudly without its CUDA extension.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "lib", "libautognothi_b200.so")
HEADER_PATH = os.path.join(ROOT, "include", "autognothi_b200.h")

_CTYPES = {
    "int": ctypes.c_int,
    "float": ctypes.c_float,
    "long long": ctypes.c_longlong,
    "uint64_t": ctypes.c_uint64,
    "double": ctypes.c_double,
}


def parse_header(path: str = HEADER_PATH) -> Dict[str, Tuple[str, List[Tuple[str, str]]]]:
    """-> {name: (return_type, [(ctype_string, arg_name), ...])} for every prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", " ", src, flags=re.M)
    src = src.replace('extern "C" {', " ")
    protos: Dict[str, Tuple[str, List[Tuple[str, str]]]] = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(agb_\w+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        parsed: List[Tuple[str, str]] = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"(.+?)\s*(\w+)$", a)
                assert mm, f"cannot parse argument {a!r} of {name}"
                parsed.append((mm.group(1).strip(), mm.group(2)))
        protos[name] = (ret, parsed)
    return protos


def _to_ctype(t: str):
    if "*" in t:
        return ctypes.c_char_p if t.replace(" ", "") == "constchar*" else ctypes.c_void_p
    t = t.replace("const ", "").strip()
    return _CTYPES[t]


PROTOTYPES = parse_header()

if not os.path.exists(LIB_PATH):
    raise RuntimeError(
        f"autognothi_b200: native library not found at {LIB_PATH}. Build it with "
        "`python -c 'import __graft_entry__ as g; g.build()'` (or `make`). There is no CPU fallback."
    )

lib = ctypes.CDLL(LIB_PATH)
for _name, (_ret, _args) in PROTOTYPES.items():
    _fn = getattr(lib, _name)  # AttributeError here == header/library mismatch: fail loudly
    _fn.restype = _to_ctype(_ret)
    _fn.argtypes = [_to_ctype(t) for t, _ in _args]


TORCH_LIB_PATH = os.path.join(_HERE, "lib", "libagb_torch.so")
BINDING = os.environ.get("AGB_BINDING", "torch")          # "torch" (default) | "ctypes"
assert BINDING in ("torch", "ctypes"), "AGB_BINDING must be 'torch' or 'ctypes'"
TORCH_OPS = None
if BINDING == "torch":
    if not os.path.exists(TORCH_LIB_PATH):
        raise RuntimeError(
            f"autognothi_b200: torch extension not found at {TORCH_LIB_PATH}. Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or `make`). There is no CPU fallback.")
    torch.ops.load_library(TORCH_LIB_PATH)
    TORCH_OPS = {n: getattr(torch.ops.agb, n) for n, (_r, a) in PROTOTYPES.items()
                 if _r == "int" and any(pn == "stream" and "*" in pt for pt, pn in a)}


class NativeError(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib.agb_last_error()
        raise NativeError(f"{what} failed (status {rc}): {msg.decode() if msg else '?'}")


class _Stream:
    """placeholder for the C-ABI's stream argument: the torch extension supplies at::cuda's current stream itself"""


_STREAM = _Stream()


def ptr(t):
    """device-pointer argument of a C-ABI call (None -> NULL): the tensor itself for the torch extension (which checks the
    device and takes data_ptr() in C++), a ctypes pointer for the ctypes binding."""
    if t is None:
        return None
    assert t.is_cuda, "autognothi_b200 kernels take CUDA tensors only"
    if TORCH_OPS is not None:
        return t
    assert t.device.index == torch.cuda.current_device(), \
        f"tensor on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}"
    return ctypes.c_void_p(t.data_ptr())


def stream():
    if TORCH_OPS is not None:
        return _STREAM
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _invoke(name: str, args) -> int:
    op = TORCH_OPS.get(name) if TORCH_OPS is not None else None
    if op is not None:
        conv = []
        for a in args:
            if a is _STREAM:
                continue
            if isinstance(a, int) and a >= (1 << 63):
                a -= 1 << 64                      # uint64_t arguments (hash seeds) travel as two's-complement int64
            conv.append(a)
        return op(*conv)
    # ctypes path: configuration entry points, or AGB_BINDING=ctypes
    conv = []
    for a in args:
        if a is _STREAM:
            a = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        elif isinstance(a, torch.Tensor):
            a = ctypes.c_void_p(a.data_ptr())
        conv.append(a)
    return getattr(lib, name)(*conv)


LAUNCHES = 0       # number of kernel-launching C-ABI calls made through `call` (bench.py's gpu_launches)
PROFILE = None     # when a list: every call appends (name, meta, start_event, end_event) — bench.py's roofline leg
NEXT_META = None   # set by ops wrappers right before `call` (e.g. algorithmic FLOPs of a GEMM)
NEXT_INFO = None   # optional (shape tag, algorithmic bytes) of the same call, for bench.py's per-shape roofline


def call(name: str, *args) -> None:
    global LAUNCHES, NEXT_META, NEXT_INFO
    LAUNCHES += 1
    if PROFILE is None:
        check(_invoke(name, args), name)
        return
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    check(_invoke(name, args), name)
    e1.record()
    PROFILE.append((name, NEXT_META, e0, e1, NEXT_INFO))
    NEXT_META = None
    NEXT_INFO = None
