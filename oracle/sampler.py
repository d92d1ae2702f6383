"""ctypes front-ends for the sampler checkers (TEST INFRASTRUCTURE ONLY).

OracleSampler -> oracle/liblogu_oracle.so (C restatement, travels to the GPU box)
RefSampler    -> oracle/_ref/liblogu_ref.so (the reference class compiled from its own sources;
                 built only where /root/reference exists, shipped prebuilt to the GPU box)
Both mirror log_uniform.pyx:16-40 (`sample(size, labels) -> (ids, true_freq, sample_freq)`).
"""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(ref=True):
    subprocess.run(["make", "-C", _HERE, "oracle"], check=True, capture_output=True)
    if ref:
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)


def _i64(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


class _Base:
    _prefix = None
    _lib = None

    def __init__(self, n):
        lib = self._load()
        f = lambda name: getattr(lib, self._prefix + name)
        f("new").restype = ctypes.c_void_p
        f("new").argtypes = [ctypes.c_int]
        f("free").argtypes = [ctypes.c_void_p]
        f("probability").restype = ctypes.c_float
        f("probability").argtypes = [ctypes.c_void_p, ctypes.c_int]
        f("sample").argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64),
                                ctypes.POINTER(ctypes.c_int)]
        f("expected_count").argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int64),
                                        ctypes.c_int64, ctypes.POINTER(ctypes.c_float)]
        f("sample_unique").argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64),
                                       ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]
        self._f = f
        self.n = n
        self._h = ctypes.c_void_p(f("new")(n))

    def __del__(self):
        if getattr(self, "_h", None):
            self._f("free")(self._h)
            self._h = None

    def probability(self, idx):
        return float(self._f("probability")(self._h, int(idx)))

    def sample_with_tries(self, size):
        if size > self.n:
            raise ValueError("size > N: the reference loops forever")
        out = np.empty(size, dtype=np.int64)
        tries = ctypes.c_int(0)
        self._f("sample")(self._h, size, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), ctypes.byref(tries))
        return out, tries.value

    def expected_count(self, tries, ids):
        a, p = _i64(ids)
        out = np.empty(a.size, dtype=np.float32)
        self._f("expected_count")(self._h, tries, p, a.size, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
        return out

    def sample(self, size, labels):
        ids, tries = self.sample_with_tries(size)
        return list(ids), list(self.expected_count(tries, labels)), list(self.expected_count(tries, ids))

    def sample_unique(self, size, labels):
        a, p = _i64(list(labels))
        out = np.empty(size, dtype=np.int64)
        self._f("sample_unique")(self._h, size, p, a.size, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
        return list(out)


class OracleSampler(_Base):
    _prefix = "logu_oracle_"

    @classmethod
    def _load(cls):
        if cls._lib is None:
            path = os.path.join(_HERE, "liblogu_oracle.so")
            if not os.path.exists(path):
                build(ref=False)
            cls._lib = ctypes.CDLL(path)
        return cls._lib


class RefSampler(_Base):
    _prefix = "logu_ref_"

    @staticmethod
    def available():
        return os.path.exists(os.path.join(_HERE, "_ref", "liblogu_ref.so"))

    @classmethod
    def _load(cls):
        if cls._lib is None:
            cls._lib = ctypes.CDLL(os.path.join(_HERE, "_ref", "liblogu_ref.so"))
        return cls._lib
