"""ctypes wrapper of the oracle's C++ restatement of the reference CPU path (oracle/cpp/cpu_port.cpp).
TEST / BASELINE INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np
import scipy.sparse as sp

from . import build_oracle

_lib = None


class Result(C.Structure):
    _fields_ = [("f_init", C.c_double), ("gradnorm_init", C.c_double), ("f_opt", C.c_double),
                ("gradnorm_opt", C.c_double), ("relative_change", C.c_double), ("tcg_iterations", C.c_int),
                ("tcg_status", C.c_int), ("outer_iterations", C.c_int), ("rejections", C.c_int), ("spmv", C.c_int),
                ("solves", C.c_int)]


def _load():
    global _lib
    if _lib is None:
        path = build_oracle.LIB
        if not os.path.exists(path):
            path = build_oracle.build()
        lib = C.CDLL(path)
        lib.cpu_port_create.restype = C.c_void_p
        lib.cpu_port_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.cpu_port_destroy.argtypes = [C.c_void_p]
        lib.cpu_port_set_G.argtypes = [C.c_void_p, C.c_void_p]
        lib.cpu_port_nnzL.restype = C.c_long
        lib.cpu_port_nnzL.argtypes = [C.c_void_p]
        lib.cpu_port_optimize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                          C.c_void_p, C.c_void_p, C.POINTER(Result)]
        _lib = lib
    return _lib


def available() -> bool:
    try:
        _load()
        return True
    except Exception:
        return False


class Runner:
    """One agent's problem on the CPU; step() = one optimize() with the updateX constants."""

    def __init__(self, Q, n, d, r, threads: int = 1):
        self.lib = _load()
        Q = sp.csr_matrix(Q)
        Q.sort_indices()
        self.n, self.d, self.r, self.N = n, d, r, (d + 1) * n
        self.threads = threads
        self._rowptr = np.ascontiguousarray(Q.indptr, dtype=np.int32)
        self._colind = np.ascontiguousarray(Q.indices, dtype=np.int32)
        self._val = np.ascontiguousarray(Q.data, dtype=np.float64)
        self.h = self.lib.cpu_port_create(n, d, r, self._rowptr.ctypes.data, self._colind.ctypes.data,
                                          self._val.ctypes.data, threads)
        self.result = Result()
        self.params = dict(algorithm=0, tr_iterations=1, max_inner=10, tol=1e-2, radius=100.0, rgd_step=1e-3)

    def set_G(self, G):
        Gf = np.asfortranarray(np.asarray(G, dtype=np.float64))
        self.lib.cpu_port_set_G(self.h, Gf.ctypes.data)

    def nnzL(self):
        return int(self.lib.cpu_port_nnzL(self.h))

    def step(self, X):
        Xf = np.asfortranarray(np.asarray(X, dtype=np.float64))
        out = np.empty_like(Xf, order="F")
        p = self.params
        self.lib.cpu_port_optimize(self.h, p["algorithm"], p["tr_iterations"], p["max_inner"], p["tol"], p["radius"],
                                   p["rgd_step"], Xf.ctypes.data, out.ctypes.data, C.byref(self.result))
        return out

    def __del__(self):
        try:
            self.lib.cpu_port_destroy(self.h)
        except Exception:
            pass
