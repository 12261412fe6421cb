"""ctypes binding of the CUDA library (csrc/libg3cuda.so) through the C ABI of include/g3cuda.h.

There is no fallback: if the library is missing, or no CUDA device is usable, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("G3CUDA_LIBRARY") or os.path.join(_HERE, "csrc", "libg3cuda.so")
_LIB = None


class G3CudaError(RuntimeError):
    pass


class _Spec(C.Structure):
    _fields_ = [("version", C.c_int32), ("n_ints", C.c_int64), ("ints", C.POINTER(C.c_int32)),
                ("n_reals", C.c_int64), ("reals", C.POINTER(C.c_double))]


def build():
    """Compile csrc/ with nvcc for sm_100a (cross-compiles without a GPU)."""
    subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "libg3cuda.so"], check=True)
    return LIB_PATH


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise G3CudaError(f"CUDA library not built: {LIB_PATH} (run __graft_entry__.build()); "
                          "there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, ip, dp = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double)
    lib.g3cuda_model_create.argtypes = [C.POINTER(_Spec), C.c_int, C.POINTER(vp)]
    lib.g3cuda_model_create.restype = C.c_int
    lib.g3cuda_model_free.argtypes = [vp]
    lib.g3cuda_model_free.restype = None
    for n in ("g3cuda_n_par", "g3cuda_n_nll", "g3cuda_n_steps"):
        getattr(lib, n).argtypes = [vp]
        getattr(lib, n).restype = C.c_int32
    for n in ("g3cuda_report_len", "g3cuda_launch_count"):
        getattr(lib, n).argtypes = [vp]
        getattr(lib, n).restype = C.c_int64
    lib.g3cuda_eval_batch.argtypes = [vp, dp, C.c_int64, ip, C.c_int32, dp, dp, dp]
    lib.g3cuda_eval_batch.restype = C.c_int
    lib.g3cuda_eval_batch_device.argtypes = [vp, vp, C.c_int64, vp, C.c_int32, vp, vp, vp, vp]
    lib.g3cuda_eval_batch_device.restype = C.c_int
    lib.g3cuda_last_kernel_ms.argtypes = [vp]
    lib.g3cuda_last_kernel_ms.restype = C.c_float
    lib.g3cuda_report.argtypes = [vp, dp, dp]
    lib.g3cuda_report.restype = C.c_int
    lib.g3cuda_last_error.restype = C.c_char_p
    lib.g3cuda_abi_version.restype = C.c_int32
    lib.g3cuda_set_kernel.argtypes = [vp, C.c_int]
    lib.g3cuda_set_kernel.restype = C.c_int
    lib.g3cuda_set_tile_warps.argtypes = [vp, C.c_int]
    lib.g3cuda_set_tile_warps.restype = C.c_int
    lib.g3cuda_set_tile_slots.argtypes = [vp, C.c_int]
    lib.g3cuda_set_tile_slots.restype = C.c_int
    lib.g3cuda_set_tile_partition.argtypes = [vp, C.c_int, C.c_int]
    lib.g3cuda_set_tile_partition.restype = C.c_int
    lib.g3cuda_multi_create.argtypes = [C.POINTER(_Spec), ip, C.c_int32, C.POINTER(vp)]
    lib.g3cuda_multi_create.restype = C.c_int
    lib.g3cuda_multi_free.argtypes = [vp]
    lib.g3cuda_multi_free.restype = None
    lib.g3cuda_multi_n_devices.argtypes = [vp]
    lib.g3cuda_multi_n_devices.restype = C.c_int32
    lib.g3cuda_multi_model.argtypes = [vp, C.c_int32]
    lib.g3cuda_multi_model.restype = vp
    lib.g3cuda_multi_eval_batch.argtypes = [vp, dp, C.c_int64, ip, C.c_int32, dp, dp, dp]
    lib.g3cuda_multi_eval_batch.restype = C.c_int
    lib.g3cuda_tile_info.argtypes = [vp, ip]
    lib.g3cuda_tile_info.restype = C.c_int32
    lib.g3cuda_math_probe.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int64,
                                      C.POINTER(C.c_double)]
    lib.g3cuda_math_probe.restype = C.c_int
    lib.g3cuda_fp64_peak_tflops.argtypes = [C.c_int]
    lib.g3cuda_fp64_peak_tflops.restype = C.c_double
    _LIB = lib
    return lib


EXPORTED = ["g3cuda_model_create", "g3cuda_model_free", "g3cuda_n_par", "g3cuda_n_nll", "g3cuda_n_steps",
            "g3cuda_report_len", "g3cuda_eval_batch", "g3cuda_eval_batch_device", "g3cuda_launch_count",
            "g3cuda_last_kernel_ms", "g3cuda_report", "g3cuda_last_error", "g3cuda_abi_version",
            "g3cuda_fp64_peak_tflops", "g3cuda_set_kernel", "g3cuda_math_probe", "g3cuda_set_tile_warps", "g3cuda_set_tile_slots", "g3cuda_set_tile_partition", "g3cuda_tile_info",
            "g3cuda_multi_create", "g3cuda_multi_free", "g3cuda_multi_n_devices", "g3cuda_multi_model", "g3cuda_multi_eval_batch"]


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


class Handle:
    """Owns one g3cuda_model*."""

    def __init__(self, ints, reals, device=0):
        self.lib = load()
        self.ints = np.ascontiguousarray(ints, dtype=np.int32)
        self.reals = np.ascontiguousarray(reals, dtype=np.float64)
        from .spec import K
        spec = _Spec(K["G3CUDA_SPEC_VERSION"], len(self.ints), self.ints.ctypes.data_as(C.POINTER(C.c_int32)),
                     len(self.reals), self.reals.ctypes.data_as(C.POINTER(C.c_double)))
        h = C.c_void_p()
        rc = self.lib.g3cuda_model_create(C.byref(spec), int(device), C.byref(h))
        if rc != 0:
            raise G3CudaError("g3cuda_model_create: %s (code %d)" % (self.lib.g3cuda_last_error().decode(), rc))
        self.h = h
        self.device = int(device)
        self.n_par = self.lib.g3cuda_n_par(h)
        self.n_nll = self.lib.g3cuda_n_nll(h)
        self.n_steps = self.lib.g3cuda_n_steps(h)

    def close(self):
        if getattr(self, "h", None):
            self.lib.g3cuda_model_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise G3CudaError("%s: %s (code %d)" % (what, self.lib.g3cuda_last_error().decode(), rc))

    def eval_batch(self, par, tangent_idx=None, want_comp=False):
        """HOST buffers in, HOST buffers out (includes the host<->device copies)."""
        par = np.ascontiguousarray(np.atleast_2d(par), dtype=np.float64)
        B = par.shape[0]
        if par.shape[1] != self.n_par:
            raise ValueError(f"expected {self.n_par} parameters per vector, got {par.shape[1]}")
        nll = np.empty(B)
        comp = np.empty((B, self.n_nll)) if want_comp else None
        grad, tidx, Kt = None, None, 0
        if tangent_idx is not None:
            tidx = np.ascontiguousarray(tangent_idx, dtype=np.int32)
            Kt = len(tidx)
            grad = np.empty((B, Kt))
        rc = self.lib.g3cuda_eval_batch(self.h, _dp(par), B,
                                        None if tidx is None else tidx.ctypes.data_as(C.POINTER(C.c_int32)), Kt,
                                        _dp(nll), _dp(comp), _dp(grad))
        self._check(rc, "g3cuda_eval_batch")
        return nll, comp, grad

    def eval_batch_device(self, d_par, B, d_nll, d_comp=0, d_tidx=0, K=0, d_grad=0, stream=0):
        """DEVICE pointers (ints), asynchronous on `stream`."""
        rc = self.lib.g3cuda_eval_batch_device(self.h, d_par, int(B), d_tidx or None, int(K), d_nll,
                                               d_comp or None, d_grad or None, stream or None)
        self._check(rc, "g3cuda_eval_batch_device")

    def set_kernel(self, which: int):
        """0 = automatic, 3 = thread-per-evaluation kernel, 4 = tile kernel, 5 = tile kernel in its full instantiation (tests / profiling)."""
        self._check(self.lib.g3cuda_set_kernel(self.h, int(which)), "g3cuda_set_kernel")

    def set_tile_warps(self, warps: int):
        """Warps per tile of the tile kernel (0 = the planner's choice); tuning hook."""
        self._check(self.lib.g3cuda_set_tile_warps(self.h, int(warps)), "g3cuda_set_tile_warps")

    def set_tile_partition(self, warps: int = 0, seg_rows: int = 0):
        """Re-plan the tile kernel: unit warps (0 = planner's choice), rows per unit (0 = planner's choice, -1 = whole columns)."""
        self._check(self.lib.g3cuda_set_tile_partition(self.h, int(warps), int(seg_rows)), "g3cuda_set_tile_partition")

    def set_tile_slots(self, tiles: int):
        """At most `tiles` tiles of the tile kernel in flight (0 = every slot of the device)."""
        self._check(self.lib.g3cuda_set_tile_slots(self.h, int(tiles)), "g3cuda_set_tile_slots")

    def tile_info(self) -> dict:
        out = np.zeros(6, dtype=np.int32)
        self._check(self.lib.g3cuda_tile_info(self.h, out.ctypes.data_as(C.POINTER(C.c_int32))), "g3cuda_tile_info")
        return dict(zip(("tile_ok", "warps", "table_entries", "state_entries", "state_in_smem", "thread_ok"), out.tolist()))

    def last_kernel_ms(self) -> float:
        return float(self.lib.g3cuda_last_kernel_ms(self.h))

    def launch_count(self) -> int:
        return int(self.lib.g3cuda_launch_count(self.h))

    def report(self, par):
        par = np.ascontiguousarray(par, dtype=np.float64).ravel()
        out = np.empty(self.lib.g3cuda_report_len(self.h))
        self._check(self.lib.g3cuda_report(self.h, _dp(par), _dp(out)), "g3cuda_report")
        return out


def math_probe(which, x, y=None, device=0):
    """The kernels' own logspace_add (0) / avoid_zero (1) / dif_pmin(., ., 1e3) (2) / exact-division avoid_zero (3) /
    exact-division dif_pmin(., ., 1e3) (4) on HOST arrays (test hook)."""
    lib = load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = None if y is None else np.ascontiguousarray(y, dtype=np.float64)
    out = np.empty_like(x)
    rc = lib.g3cuda_math_probe(int(device), int(which), _dp(x), _dp(y), x.size, _dp(out))
    if rc != 0:
        raise G3CudaError("g3cuda_math_probe: %s (code %d)" % (lib.g3cuda_last_error().decode(), rc))
    return out


class _Borrowed(Handle):
    """A device's handle inside a MultiHandle: same calls, owned by the multi-device handle."""

    def __init__(self, lib, h, device):
        self.lib, self.h, self.device = lib, h, int(device)
        self.n_par = lib.g3cuda_n_par(h)
        self.n_nll = lib.g3cuda_n_nll(h)
        self.n_steps = lib.g3cuda_n_steps(h)

    def close(self):
        self.h = None


class MultiHandle:
    """Owns one g3cuda_multi*: the model on several devices, batches cut into one row block per device."""

    def __init__(self, ints, reals, devices):
        self.lib = load()
        self.ints = np.ascontiguousarray(ints, dtype=np.int32)
        self.reals = np.ascontiguousarray(reals, dtype=np.float64)
        from .spec import K
        spec = _Spec(K["G3CUDA_SPEC_VERSION"], len(self.ints), self.ints.ctypes.data_as(C.POINTER(C.c_int32)),
                     len(self.reals), self.reals.ctypes.data_as(C.POINTER(C.c_double)))
        dv = np.ascontiguousarray(list(devices), dtype=np.int32)
        h = C.c_void_p()
        rc = self.lib.g3cuda_multi_create(C.byref(spec), dv.ctypes.data_as(C.POINTER(C.c_int32)), len(dv), C.byref(h))
        if rc != 0:
            raise G3CudaError("g3cuda_multi_create: %s (code %d)" % (self.lib.g3cuda_last_error().decode(), rc))
        self.h = h
        self.devices = [int(d) for d in dv]
        self.handles = [_Borrowed(self.lib, C.c_void_p(self.lib.g3cuda_multi_model(h, i)), d) for i, d in enumerate(self.devices)]
        self.n_par, self.n_nll = self.handles[0].n_par, self.handles[0].n_nll

    def close(self):
        if getattr(self, "h", None):
            self.lib.g3cuda_multi_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def eval_batch(self, par, tangent_idx=None, want_comp=False):
        """HOST buffers in, HOST buffers out; the devices work on their row blocks concurrently."""
        par = np.ascontiguousarray(np.atleast_2d(par), dtype=np.float64)
        B = par.shape[0]
        if par.shape[1] != self.n_par:
            raise ValueError(f"expected {self.n_par} parameters per vector, got {par.shape[1]}")
        nll = np.empty(B)
        comp = np.empty((B, self.n_nll)) if want_comp else None
        grad, tidx, Kt = None, None, 0
        if tangent_idx is not None:
            tidx = np.ascontiguousarray(tangent_idx, dtype=np.int32)
            Kt = len(tidx)
            grad = np.empty((B, Kt))
        rc = self.lib.g3cuda_multi_eval_batch(self.h, _dp(par), B,
                                              None if tidx is None else tidx.ctypes.data_as(C.POINTER(C.c_int32)), Kt,
                                              _dp(nll), _dp(comp), _dp(grad))
        if rc != 0:
            raise G3CudaError("g3cuda_multi_eval_batch: %s (code %d)" % (self.lib.g3cuda_last_error().decode(), rc))
        return nll, comp, grad
