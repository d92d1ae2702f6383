"""Loader for the C-ABI CUDA library (include/u2gnn_b200.h).

The product path has no CPU fallback: if ``libu2gnn_b200.so`` is missing the import fails loudly,
and every call raises ``RuntimeError`` with ``u2gnn_strerror`` on a non-zero status.  Signatures
are parsed from the header so the binding cannot drift from the declared boundary.
"""
from __future__ import annotations

import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libu2gnn_b200.so")
PROBE_LIB_PATH = os.path.join(_HERE, "libu2gnn_b200_probe.so")
HEADER_PATH = os.path.normpath(os.path.join(_HERE, "..", "..", "include", "u2gnn_b200.h"))
PROBE_HEADER_PATH = os.path.normpath(os.path.join(_HERE, "..", "..", "include", "u2gnn_b200_probe.h"))

_SCALARS = {
    "int": ctypes.c_int, "float": ctypes.c_float, "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64,
    "uint32_t": ctypes.c_uint32, "int32_t": ctypes.c_int32, "size_t": ctypes.c_size_t,
    "u2gnn_stream_t": ctypes.c_void_p,
}
_RET = {"int": ctypes.c_int, "size_t": ctypes.c_size_t, "uint32_t": ctypes.c_uint32, "const char*": ctypes.c_char_p}


def parse_header(path=HEADER_PATH):
    """-> {name: (restype, [(ctype, argname), ...])} for every function the header declares."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    out = {}
    for m in re.finditer(r"(const char\*|int|size_t|uint32_t)\s+(u2gnn_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1), m.group(2), " ".join(m.group(3).split())
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    params.append((ctypes.c_void_p, a.split("*")[-1].strip()))
                else:
                    ty, nm = a.rsplit(" ", 1)
                    params.append((_SCALARS[ty.replace("const ", "").strip()], nm))
        out[name] = (_RET[ret], params)
    return out


SIGNATURES = parse_header()


class _Lib:
    def __init__(self, path=LIB_PATH, signatures=None):
        signatures = SIGNATURES if signatures is None else signatures
        if not os.path.exists(path):
            raise ImportError(
                "u2gnn_b200: %s is missing - build it with `python graph-transformer_b200/build.py` "
                "(there is no CPU or PyTorch fallback for this path)" % path)
        self.cdll = ctypes.CDLL(path)
        for name, (ret, params) in signatures.items():
            fn = getattr(self.cdll, name)     # AttributeError if the library lacks a declared symbol
            fn.restype = ret
            fn.argtypes = [t for t, _ in params]
        self._status = {n for n, (r, _) in signatures.items() if r is ctypes.c_int}
        self.launches = 0          # C-ABI calls that launch kernels (bench.py's gpu_launches claim)
        self.timed = None          # {entry point: [(start_event, end_event), ...]} when profiling
        # Call path: cffi in ABI mode on the same declarations when it is importable (2 us per 22-argument call against 6.6 us
        # through ctypes - the launch-bound small configurations issue ~100 calls per 1.8 ms step), ctypes otherwise.
        self._fn = {n: getattr(self.cdll, n) for n in signatures}
        try:
            import cffi
            cmap = {ctypes.c_int: "int", ctypes.c_float: "float", ctypes.c_int64: "int64_t", ctypes.c_uint64: "uint64_t",
                    ctypes.c_uint32: "uint32_t", ctypes.c_int32: "int32_t", ctypes.c_size_t: "size_t", ctypes.c_void_p: "uintptr_t"}
            decl = ["%s %s(%s);" % (cmap[r], n, ", ".join(cmap[t] for t, _ in ps) or "void")
                    for n, (r, ps) in signatures.items() if r in cmap]
            ffi = cffi.FFI()
            ffi.cdef("\n".join(decl))
            self._ffi_lib = ffi.dlopen(path)
            for n, (r, _) in signatures.items():
                if r in cmap:
                    self._fn[n] = getattr(self._ffi_lib, n)
        except Exception:                    # cffi missing / header construct it cannot parse: the ctypes bindings stay
            pass

    def strerror(self, code):
        return self.cdll.u2gnn_strerror(code).decode()

    def call(self, name, *args):
        timed = self.timed
        if timed is not None and name in timed:
            import torch
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rc = self._fn[name](*args)
            b.record()
            timed[name].append((a, b))
        else:
            rc = self._fn[name](*args)
        self.launches += 1
        if name in self._status and rc != 0 and name != "u2gnn_device_check":
            raise RuntimeError("%s failed: %s (%d)" % (name, self.strerror(rc), rc))
        return rc


LIB = _Lib()
_PROBE = None


def probe_lib():
    """The PROBE library (layout self-tests, micro-benchmarks, kernel tracing; include/u2gnn_b200_probe.h): a superset build
    of the product sources loaded side by side with it.  Only tools/ and the hardware-layout tests use it."""
    global _PROBE
    if _PROBE is None:
        sig = dict(SIGNATURES)
        sig.update(parse_header(PROBE_HEADER_PATH))
        _PROBE = _Lib(PROBE_LIB_PATH, sig)
    return _PROBE


def require_device():
    """Raises unless the current CUDA device can run the sm_100a kernels."""
    rc = LIB.cdll.u2gnn_device_check()
    if rc != 0:
        raise RuntimeError("u2gnn_b200 needs an sm_100 (B200) device: " + LIB.strerror(rc))
