"""ctypes binding of libmednet_b200.so, generated from include/mednet_b200.h.

The struct layouts and prototypes are parsed from the public header at import time so the Python side
can never drift from the C ABI.  There is no fallback: if the shared object is missing the first call
raises (build it with ``python torch-mednet_b200/csrc/build.py`` or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os
import re

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MEDNET_B200_LIB", os.path.join(_PKG, "libmednet_b200.so"))   # override: A/B builds
HEADER_PATH = os.environ.get(
    "MEDNET_B200_HEADER",
    os.path.join(os.path.dirname(os.path.dirname(_PKG)), "include", "mednet_b200.h"))

_SCALARS = {"int32_t": C.c_int32, "int": C.c_int32, "int64_t": C.c_int64, "float": C.c_float, "size_t": C.c_size_t,
            "uint8_t": C.c_uint8}
_TYPE_WORDS = set(_SCALARS) | {"const", "void", "char", "unsigned"}


def _strip_comments(text):
    return re.sub(r"/\*.*?\*/", " ", text, flags=re.S)


def parse_header(path=HEADER_PATH):
    """Returns (defines, structs, functions).

    structs:   name -> [(field, ctype)]
    functions: name -> (restype, [argtypes])
    """
    raw = open(path).read()
    defines = {}
    for m in re.finditer(r"#define\s+(MEDNET_\w+)\s+\(?(-?\d+)\)?", raw):
        defines[m.group(1)] = int(m.group(2))
    text = _strip_comments(raw)
    structs = {}
    for m in re.finditer(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        fields = []
        for stmt in m.group(1).split(";"):
            stmt = stmt.strip()
            if not stmt:
                continue
            is_ptr = "*" in stmt
            words = stmt.replace("*", " ").replace(",", " ").split()
            base = [w for w in words if w in _TYPE_WORDS]
            names = [w for w in words if w not in _TYPE_WORDS]
            if is_ptr:
                ctype = C.c_void_p
            else:
                key = [b for b in base if b in _SCALARS]
                ctype = _SCALARS[key[-1]]
            fields.extend((n, ctype) for n in names)
        structs[m.group(2)] = fields
    text_nostruct = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", " ", text, flags=re.S)
    functions = {}
    for m in re.finditer(r"(const\s+char\s*\*|size_t|int)\s+(mednet_\w+)\s*\(([^)]*)\)\s*;", text_nostruct):
        ret = m.group(1)
        restype = C.c_char_p if "char" in ret else (C.c_size_t if ret == "size_t" else C.c_int)
        args = []
        params = m.group(3).strip()
        if params and params != "void":
            for prm in params.split(","):
                prm = prm.strip()
                if "*" in prm or "mednet_stream_t" in prm:
                    args.append(C.c_void_p)
                else:
                    key = [w for w in prm.split() if w in _SCALARS]
                    args.append(_SCALARS[key[0]])
        functions[m.group(2)] = (restype, args)
    return defines, structs, functions


DEFINES, _STRUCT_FIELDS, FUNCTIONS = parse_header()
globals().update(DEFINES)


def _make_struct(name, fields):
    return type(name, (C.Structure,), {"_fields_": fields})


STRUCTS = {name: _make_struct(name, fields) for name, fields in _STRUCT_FIELDS.items()}

_lib = None


def lib():
    """The loaded shared object with restype/argtypes set.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"mednet_b200: CUDA extension {LIB_PATH} is missing -- build it with "
                "`python torch-mednet_b200/csrc/build.py`; there is no CPU or PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (restype, args) in FUNCTIONS.items():
            fn = getattr(handle, name)          # AttributeError here = header/library drift
            fn.restype = restype
            fn.argtypes = args
        if handle.mednet_abi_version() != DEFINES["MEDNET_ABI_VERSION"]:
            raise RuntimeError("mednet_b200: ABI version mismatch between header and library")
        _lib = handle
    return _lib


class MednetError(RuntimeError):
    pass


def check(code, what):
    if code != 0:
        msg = lib().mednet_error_string(code)
        raise MednetError(f"{what} failed: {msg.decode() if msg else code} (code {code})")


def make(struct_name, **kw):
    """Instantiate a params struct; pointer fields accept ints or None."""
    s = STRUCTS[struct_name]()
    valid = {f for f, _ in _STRUCT_FIELDS[struct_name]}
    for k, v in kw.items():
        if k not in valid:
            raise KeyError(f"{struct_name} has no field {k}")
        setattr(s, k, v)
    return s
