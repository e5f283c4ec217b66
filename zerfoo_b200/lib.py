"""ctypes loader for libkernels.so -- the role internal/cuda/kernels/purego.go:129-275
plays in the reference: dlopen the library (``ZERFOO_KERNEL_LIB_PATH`` override,
internal/cuda/purego.go:210-245), resolve every launcher by name, and fail
loudly if the library or a symbol is missing.  No fallback path exists."""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
from typing import Dict, List, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
DEFAULT_LIB = os.path.join(_HERE, "lib", "libkernels.so")
INCLUDE = os.path.join(ROOT, "include")


class KernelLibraryError(RuntimeError):
    pass


def lib_path() -> str:
    return os.environ.get("ZERFOO_KERNEL_LIB_PATH") or DEFAULT_LIB


def build(force: bool = False, jobs: int = 8) -> str:
    """Compile libkernels.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    csrc = os.path.join(_HERE, "csrc")
    if force:
        subprocess.check_call(["make", "-C", csrc, "clean", "-s"])
    subprocess.check_call(["make", "-C", csrc, f"-j{jobs}", "-s"])
    return DEFAULT_LIB


_C_TYPES = {
    "int": C.c_int, "unsigned int": C.c_uint, "uint32_t": C.c_uint32, "uint64_t": C.c_uint64, "int64_t": C.c_int64,
    "int32_t": C.c_int32, "long long": C.c_longlong, "unsigned long long": C.c_ulonglong, "float": C.c_float,
    "cudaStream_t": C.c_void_p, "zb_stream_t": C.c_void_p, "cudaError_t": C.c_int, "void": None,
}


def _ctype(decl: str):
    decl = decl.strip()
    if decl.endswith("**"):
        return C.POINTER(C.c_void_p)
    if "*" in decl:
        base = decl.replace("const", "").replace("*", "").strip()
        if base == "char":
            return C.c_char_p
        return C.c_void_p
    decl = decl.replace("const ", "").strip()
    return _C_TYPES[decl]


_PROTO = re.compile(r"^\s*(?:const\s+)?([A-Za-z_][\w \*]*?)\s*\b(\w+)\s*\(([^;{]*)\)\s*;", re.M)


def declared_symbols(header: str) -> List[Tuple[str, str, List[str]]]:
    """[(return type, name, [argument type strings])] for every prototype in an include/*.h file."""
    text = open(os.path.join(INCLUDE, header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"typedef[^;]*\{[^}]*\}[^;]*;", "", text, flags=re.S)
    text = re.sub(r"typedef[^;]*;", "", text)
    text = re.sub(r"enum\s*\{[^}]*\}\s*;", "", text, flags=re.S)
    out = []
    for m in _PROTO.finditer(text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if name in ("defined",):
            continue
        argt = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                # drop the parameter name
                mm = re.match(r"^(.*?)(\b\w+)?$", a)
                t = a
                if "*" in a:
                    t = a[: a.rfind("*") + 1]
                else:
                    parts = a.split()
                    t = " ".join(parts[:-1]) if len(parts) > 1 else a
                argt.append(t.strip())
        out.append((ret, name, argt))
    return out


_lib = None


def load() -> C.CDLL:
    """dlopen libkernels.so and bind every declared symbol; raises KernelLibraryError otherwise."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise KernelLibraryError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  zerfoo_b200 has no CPU or PyTorch fallback.")
    try:
        L = C.CDLL(path, mode=C.RTLD_GLOBAL)
    except OSError as e:
        raise KernelLibraryError(f"dlopen({path}) failed: {e}") from e
    missing = []
    for header in ("zerfoo_kernels.h", "zb200.h"):
        for ret, name, argt in declared_symbols(header):
            try:
                fn = getattr(L, name)
            except AttributeError:
                missing.append(name)
                continue
            fn.argtypes = [_ctype(a) for a in argt]
            if ret.endswith("*") and "char" in ret:
                fn.restype = C.c_char_p
            elif "*" in ret or ret in ("zb_stream_t", "cudaStream_t"):
                fn.restype = C.c_void_p
            elif ret == "void":
                fn.restype = None
            else:
                fn.restype = C.c_int
    if missing:
        raise KernelLibraryError(f"{path} does not export: {', '.join(missing)}")
    _lib = L
    return L


def check(rc: int, op: str) -> None:
    """checkKernel (internal/cuda/kernels/elementwise_purego.go:11-16)."""
    if rc != 0:
        msg = load().zb_last_error()
        detail = msg.decode() if msg else ""
        raise RuntimeError(f"{op} kernel failed (cuda error {rc}){': ' + detail if detail else ''}")
