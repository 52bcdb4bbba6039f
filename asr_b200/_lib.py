"""ctypes binding of libasr_b200.so (the C ABI declared in include/asr_b200.h).

The prototypes are parsed from the header itself, so the header is the single source of truth for
argument types.  There is NO fallback: if the shared library is missing or a symbol is absent,
importing this module raises; if a kernel returns a non-zero status, `call` raises RuntimeError
(CUDA status) or ValueError (bad argument / unsupported shape) -- SURVEY.md section 8b error contract.
"""
from __future__ import annotations

import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), "include", "asr_b200.h")
LIBRARY = os.path.join(_HERE, "libasr_b200.so")

_CTYPES = {
    "int": ctypes.c_int,
    "unsigned": ctypes.c_uint,
    "uint32_t": ctypes.c_uint32,
    "int32_t": ctypes.c_int32,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
    "size_t": ctypes.c_size_t,
    "long long": ctypes.c_longlong,
    "asrb_stream_t": ctypes.c_void_p,
}


def parse_header(path: str = HEADER):
    """Return {name: (restype, [(argtype, argname), ...])} for every asrb_* prototype in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"(const char\*|size_t|int)\s+(asrb_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = {"int": ctypes.c_int, "size_t": ctypes.c_size_t, "const char*": ctypes.c_char_p}[ret]
        argl = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    argl.append((ctypes.c_void_p, a.split("*")[-1].strip()))
                else:
                    ty, nm = a.rsplit(" ", 1)
                    argl.append((_CTYPES[ty.replace("const ", "").strip()], nm))
        protos[name] = (restype, argl)
    return protos


PROTOTYPES = parse_header()

if not os.path.exists(LIBRARY):
    raise ImportError(
        f"{LIBRARY} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(asr_b200 has no CPU or PyTorch fallback)")

_dll = ctypes.CDLL(LIBRARY)
_fns = {}
for _name, (_res, _args) in PROTOTYPES.items():
    _f = getattr(_dll, _name)  # AttributeError here = header/library mismatch
    _f.restype = _res
    _f.argtypes = [a for a, _ in _args]
    _fns[_name] = _f


def strerror(code: int) -> str:
    return _fns["asrb_strerror"](code).decode()


def call(name: str, *args):
    """Call an int-returning entry point; raise on a non-zero status."""
    rc = _fns[name](*args)
    if rc != 0:
        msg = f"{name}: {strerror(rc)} (status {rc})"
        raise (ValueError if rc < 0 else RuntimeError)(msg)


def query(name: str, *args):
    """Call a size_t / int returning query function and hand back its value."""
    return _fns[name](*args)
