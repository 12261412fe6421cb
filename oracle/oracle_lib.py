"""ctypes binding of the CPU oracle (oracle/libg3oracle.so). TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(native=False):
    """Compile the oracle with g++ (seconds). native=True: -O3 -march=native copy under oracle/_ref/."""
    if native:
        os.makedirs(os.path.join(_HERE, "_ref"), exist_ok=True)
        out = os.path.join(_HERE, "_ref", "libg3oracle_native.so")
        cmd = ["g++", "-O3", "-march=native", "-std=gnu++17", "-fPIC", "-shared", "-o", out,
               os.path.join(_HERE, "g3oracle.cpp"), "-lpthread"]
        subprocess.run(cmd, check=True, capture_output=True)
        return out
    subprocess.run(["make", "-C", _HERE, "libg3oracle.so"], check=True, capture_output=True)
    return os.path.join(_HERE, "libg3oracle.so")


def load(path=None):
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = path or os.path.join(_HERE, "libg3oracle.so")
    if not os.path.exists(p):
        build()
    lib = C.CDLL(p)
    ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    lib.g3oracle_eval.argtypes = [ip, dp, dp, C.c_int64, ip, C.c_int32, dp, dp, dp, C.c_int32]
    lib.g3oracle_eval.restype = C.c_int
    lib.g3oracle_report.argtypes = [ip, dp, dp, dp]
    lib.g3oracle_report_len.argtypes = [ip, dp]
    lib.g3oracle_report_len.restype = C.c_int64
    for n, k in (("avoid_zero", 1), ("logspace_add", 2), ("dif_pmin", 3), ("dif_pminmax", 4), ("dnorm", 3),
                 ("ratio_add_pop", 4), ("digamma", 1), ("grow_lengthvbsimple", 4)):
        f = getattr(lib, "g3oracle_" + n)
        f.argtypes = [C.c_double] * k
        f.restype = C.c_double
    lib.g3oracle_suit_andersen.argtypes = [C.c_double] * 7
    lib.g3oracle_suit_andersen.restype = C.c_double
    lib.g3oracle_suitability.argtypes = [C.c_int32, C.POINTER(C.c_double), C.c_double, C.c_double]
    lib.g3oracle_suitability.restype = C.c_double
    lib.g3oracle_migrate_normalize.argtypes = [dp, C.c_int32, C.c_int32, C.c_double]
    lib.g3oracle_growth_bbinom.argtypes = [dp, C.c_int32, C.c_int32, C.c_double, dp]
    lib.g3oracle_grow_matrix_len.argtypes = [dp, C.c_int32, C.c_int32, dp]
    lib.g3oracle_grow_matrix_wgt.argtypes = [dp, C.c_int32, C.c_int32, dp]
    lib.g3oracle_grow_apply.argtypes = [dp, dp, C.c_int32, dp, dp, dp, dp]
    if path is None:
        _LIB = lib
    return lib


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


class Oracle:
    def __init__(self, ints, reals, lib=None):
        self.lib = lib or load()
        self.ints = np.ascontiguousarray(ints, dtype=np.int32)
        self.reals = np.ascontiguousarray(reals, dtype=np.float64)
        from gadget3_b200.spec import K
        self.n_par = int(self.ints[K["G3H_NPAR"]])
        self.n_nll = int(self.ints[K["G3H_NNLL"]])

    def eval_batch(self, par, tangent_idx=None, want_comp=False, nthreads=1):
        par = np.ascontiguousarray(np.atleast_2d(par), dtype=np.float64)
        B = par.shape[0]
        assert par.shape[1] == self.n_par
        nll = np.empty(B)
        comp = np.empty((B, self.n_nll)) if want_comp else None
        grad, tidx, Kt = None, None, 0
        if tangent_idx is not None:
            tidx = np.ascontiguousarray(tangent_idx, dtype=np.int32)
            Kt = len(tidx)
            grad = np.empty((B, Kt))
        rc = self.lib.g3oracle_eval(_ip(self.ints), _dp(self.reals), _dp(par), B,
                                    None if tidx is None else _ip(tidx), Kt, _dp(nll), _dp(comp), _dp(grad),
                                    int(nthreads))
        if rc != 0:
            raise RuntimeError("g3oracle_eval failed: %d" % rc)
        return nll, comp, grad

    def report(self, par):
        par = np.ascontiguousarray(par, dtype=np.float64).ravel()
        n = self.lib.g3oracle_report_len(_ip(self.ints), _dp(self.reals))
        out = np.empty(n)
        rc = self.lib.g3oracle_report(_ip(self.ints), _dp(self.reals), _dp(par), _dp(out))
        if rc != 0:
            raise RuntimeError("g3oracle_report failed: %d" % rc)
        return out
