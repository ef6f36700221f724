"""ctypes loader for libxdb200.so (the C ABI declared in include/xdb200.h).

The argument types of every entry point are derived from the header itself, so the Python binding
cannot drift from the declared ABI.  There is NO fallback: if the library is missing or fails to
load, importing any compute path of this package raises.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libxdb200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "xdb200.h")

_CTYPES = {
    "int": ctypes.c_int,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
    "long long": ctypes.c_longlong,
    "unsigned long long": ctypes.c_ulonglong,
}


class XdError(RuntimeError):
    pass


def parse_header(path=HEADER_PATH):
    """Returns {name: (restype, [argtypes])} for every function declared in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    out = {}
    for m in re.finditer(r"(const char\*|int)\s+(xd_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), " ".join(m.group(3).split())
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    ty = a.rsplit(" ", 1)[0].strip()
                    argtypes.append(_CTYPES[ty])
        out[name] = (ctypes.c_char_p if ret.startswith("const char") else ctypes.c_int, argtypes)
    return out


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise XdError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(xdiffusion_b200 has no CPU or PyTorch fallback)")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in parse_header().items():
            fn = getattr(l, name)          # AttributeError if the .so does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().xd_last_error().decode(errors="replace")
        raise XdError(f"{what} failed (code {rc}): {msg}")
