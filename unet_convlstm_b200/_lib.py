"""ctypes binding of libb200convlstm.so.

The prototypes are read from include/b200_convlstm.h, the single source of truth for the C ABI.
There is deliberately NO fallback: if the library is missing or a call fails, a RuntimeError is
raised (the product path never routes through PyTorch ops or the CPU oracle).
"""
from __future__ import annotations

import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), "include", "b200_convlstm.h")
LIB_PATH = os.path.join(HERE, "libb200convlstm.so")

_CTYPES = {
    "int": ctypes.c_int,
    "long long": ctypes.c_longlong,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
}


def parse_header(path: str = HEADER):
    """Returns {name: (restype, [argtypes])} for every prototype of the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
    protos = {}
    for m in re.finditer(r"([\w\s\*]+?)\b(b200_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        restype = ctypes.c_char_p if "char" in ret else ctypes.c_int
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    base = re.sub(r"\bconst\b", "", a).strip()
                    base = " ".join(base.split()[:-1])  # drop the parameter name
                    argtypes.append(_CTYPES[base])
        protos[name] = (restype, argtypes)
    return protos


_lib = None
_protos = None


def lib():
    global _lib, _protos
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m unet_convlstm_b200.build` "
                "(there is no fallback path)")
        l = ctypes.CDLL(LIB_PATH)
        _protos = parse_header()
        for name, (restype, argtypes) in _protos.items():
            fn = getattr(l, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = l
    return _lib


def last_error() -> str:
    return lib().b200_last_error().decode()


def call(name: str, *args):
    """Calls an int-returning entry point and raises on a non-zero status."""
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed: rc={rc}: {last_error()}")


def supported(name: str, *args) -> bool:
    return bool(getattr(lib(), name)(*args))
