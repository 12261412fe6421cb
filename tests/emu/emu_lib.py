"""ctypes binding of tests/emu/g3emu.cpp: the tile kernel's BODY run on the CPU (one OS thread per warp).
TEST INFRASTRUCTURE ONLY -- never imported by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "g3emu.cpp")
_OUT = os.path.join(_HERE, "_build", "libg3emu.so")
_CSRC = os.path.join(_HERE, "..", "..", "gadget3_b200", "csrc")
_LIB = None


def build():
    deps = [_SRC] + [os.path.join(_CSRC, f) for f in ("g3_tile.cuh", "g3_tile_plan.h", "g3_common.cuh", "g3_math.cuh")]
    if os.path.exists(_OUT) and all(os.path.getmtime(_OUT) >= os.path.getmtime(d) for d in deps):
        return _OUT
    os.makedirs(os.path.dirname(_OUT), exist_ok=True)
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.run(["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-Wno-attributes", "-I" + cuda_inc, "-o", _OUT, _SRC,
                    "-lpthread"], check=True, capture_output=True)
    return _OUT


def load():
    global _LIB
    if _LIB is None:
        lib = C.CDLL(build())
        ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
        lib.g3emu_eval.argtypes = [ip, dp, dp, C.c_int64, ip, C.c_int32, dp, dp, dp, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_int32, C.c_char_p, C.c_int32, C.c_int32, C.c_int32]
        lib.g3emu_eval.restype = C.c_int
        lib.g3emu_plan_info.argtypes = [ip, dp, C.c_int32, C.c_int32, ip]
        lib.g3emu_plan_info.restype = C.c_int
        _LIB = lib
    return _LIB


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def plan_info(ints, reals, W=0, seg=0):
    ints = np.ascontiguousarray(ints, dtype=np.int32); reals = np.ascontiguousarray(reals, dtype=np.float64)
    out = np.zeros(8 + 16, dtype=np.int32)
    rc = load().g3emu_plan_info(_ip(ints), _dp(reals), W, seg, _ip(out))
    if rc:
        raise RuntimeError("model does not fit the tile kernel")
    d = dict(zip(("W", "n_g", "n_s", "ncols", "maxn", "tb", "units", "nts"), out[:8].tolist()))
    d["seglen"] = out[8:8 + int(ints[9])].tolist()       # G3H_NSTOCK
    return d


def eval_batch(ints, reals, par, tangent_idx=None, want_comp=False, W=0, seg=0, lanes_per_tile=4, smem_state=True, ncta=2,
               meta_smem=True, hot_tables=True, lean=True):
    """`lean`: as the product's dispatch, the instantiation with the model's unused features compiled out (False: the full one)."""
    load().g3emu_set_no_lean(0 if lean else 1)
    ints = np.ascontiguousarray(ints, dtype=np.int32); reals = np.ascontiguousarray(reals, dtype=np.float64)
    par = np.ascontiguousarray(np.atleast_2d(par), dtype=np.float64)
    B = par.shape[0]
    nn = int(ints[32])      # G3H_NNLL
    nll = np.full(B, np.nan)
    comp = np.full((B, nn), np.nan) if want_comp else None
    grad, tidx, K = None, None, 0
    if tangent_idx is not None:
        tidx = np.ascontiguousarray(tangent_idx, dtype=np.int32); K = len(tidx)
        grad = np.full((B, K), np.nan)
    why = C.create_string_buffer(256)
    rc = load().g3emu_eval(_ip(ints), _dp(reals), _dp(par), B, None if tidx is None else _ip(tidx), K, _dp(nll), _dp(comp),
                           _dp(grad), W, seg, lanes_per_tile, int(smem_state), ncta, why, 256, int(meta_smem), int(hot_tables))
    if rc:
        raise RuntimeError("g3emu_eval: " + why.value.decode())
    return nll, comp, grad
