"""ctypes binding of libwx_b200.so (the C ABI declared in include/wx_b200.h).

The prototypes are parsed from the header itself, so the binding can never drift from the ABI and
`tests/test_cabi.py` can check that the library exports every declared symbol.

There is NO fallback: if the CUDA library is missing or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), "include", "wx_b200.h")
LIBPATH = os.environ.get("WX_B200_LIB", os.path.join(_HERE, "libwx_b200.so"))

WX_OK, WX_EINVAL, WX_ECUDA, WX_EUNSUPPORTED, WX_ENOMEM = 0, 1, 2, 3, 4


class WxError(RuntimeError):
    """non-argument failure reported by libwx_b200 (CUDA error, unsupported shape, out of memory)"""


_SCALARS = {"int": C.c_int, "long": C.c_long, "double": C.c_double, "size_t": C.c_size_t,
            "unsigned long long": C.c_ulonglong}


def _ctype(decl: str):
    decl = decl.strip()
    if decl == "void":
        return None
    if "*" in decl:
        return C.c_void_p
    base = re.sub(r"\bconst\b", "", decl).strip()
    base = re.sub(r"\s+\w+$", "", base).strip() if base not in _SCALARS else base
    return _SCALARS[base]


def parse_header(path: str = HEADER):
    """-> {name: (restype, [argtypes])} for every prototype in the header"""
    txt = open(path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"//[^\n]*", "", txt)
    txt = re.sub(r"^\s*#.*$", "", txt, flags=re.M)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(wx_\w+)\s*\(([^;{}]*)\)\s*;", txt):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if ret.endswith("*"):
            restype = C.c_char_p if "char" in ret else C.c_void_p
        else:
            restype = _SCALARS[ret]
        argtypes = [] if args in ("", "void") else [_ctype(a) for a in args.split(",")]
        protos[name] = (restype, argtypes)
    return protos


_lib = None
_protos = None


def lib() -> C.CDLL:
    """load libwx_b200.so (raises ImportError if it has not been built: run __graft_entry__.build())"""
    global _lib, _protos
    if _lib is None:
        if not os.path.exists(LIBPATH):
            raise ImportError(
                f"{LIBPATH} not found: the CUDA extension must be built (python -c 'import __graft_entry__ as g; "
                f"g.build()' or make -C waveletsext.jl_b200/csrc). There is no CPU fallback.")
        l = C.CDLL(LIBPATH)
        _protos = parse_header()
        for name, (restype, argtypes) in _protos.items():
            f = getattr(l, name)      # AttributeError if the library lacks a declared symbol
            f.restype = restype
            f.argtypes = argtypes
        _lib = l
    return _lib


def last_error() -> str:
    return lib().wx_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc == WX_OK:
        return
    msg = last_error()
    if rc == WX_EINVAL:
        if msg.startswith("ArgumentError"):
            raise ValueError(msg)       # the reference's ArgumentError
        raise AssertionError(msg)       # the reference signals these with @assert -> AssertionError
    if rc == WX_ENOMEM:
        raise MemoryError(msg)
    raise WxError(f"libwx_b200 error {rc}: {msg}")


def call(name: str, *args) -> None:
    check(getattr(lib(), name)(*args))


def launch_count() -> int:
    return int(lib().wx_launch_count())
