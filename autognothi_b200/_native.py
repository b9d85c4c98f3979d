"""ctypes binding of libautognothi_b200.so (the C-ABI declared in include/autognothi_b200.h).

The prototypes are parsed from the header itself, so Python can never drift from the C declarations.
There is NO fallback: if the shared library is missing or a symbol is absent, importing this module
raises — the product path must fail loudly without its CUDA extension.
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


class NativeError(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib.agb_last_error()
        raise NativeError(f"{what} failed (status {rc}): {msg.decode() if msg else '?'}")


def ptr(t):
    """device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda, "autognothi_b200 kernels take CUDA tensors only"
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


LAUNCHES = 0       # number of kernel-launching C-ABI calls made through `call` (bench.py's gpu_launches)
PROFILE = None     # when a list: every call appends (name, meta, start_event, end_event) — bench.py's roofline leg
NEXT_META = None   # set by ops wrappers right before `call` (e.g. algorithmic FLOPs of a GEMM)
NEXT_INFO = None   # optional (shape tag, algorithmic bytes) of the same call, for bench.py's per-shape roofline


def call(name: str, *args) -> None:
    global LAUNCHES, NEXT_META, NEXT_INFO
    LAUNCHES += 1
    if PROFILE is None:
        check(getattr(lib, name)(*args), name)
        return
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    check(getattr(lib, name)(*args), name)
    e1.record()
    PROFILE.append((name, NEXT_META, e0, e1, NEXT_INFO))
    NEXT_META = None
    NEXT_INFO = None
