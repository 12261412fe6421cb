"""Driver for the reference's own generated objectives compiled against oracle/tmb_shim (oracle/_ref/).

TEST INFRASTRUCTURE ONLY. `libref_c1.so` / `libref_c2.so` are the reference's checked-in generated C++
(baseline/introduction-single-stock.cpp, baseline/multiple-substocks.cpp) compiled unmodified where
they lie; this module feeds them the model data of our C1/C2 configurations under the names the
generated code declares (DATA_VECTOR / DATA_ARRAY / DATA_IVECTOR / DATA_SCALAR / PARAMETER), runs
`objective_function<double>::operator()` and reads back nll and REPORT()ed arrays.
The libraries can only be BUILT where /root/reference exists (the build container); they travel to
the GPU box as prebuilt files.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
SOURCES = {"C1": "ref_c1.cpp", "C2": "ref_c2.cpp", "LING": "ref_ling.cpp"}


def ling_extras(params):
    """What inttest/codegeneration/ling.cpp declares beyond the generic inputs: two data vectors, two
    PARAMETER_VECTORs (our by-age scalars lingimm.init.<age> / lingmat.init.<age>) and the dummy parameter x."""
    from gadget3_b200.configs import LING_IMM_STDDEV, LING_MAT_STDDEV
    return dict(
        extra_arrays={"ling_imm_stddev": LING_IMM_STDDEV, "ling_mat_stddev": LING_MAT_STDDEV},
        vector_params={"lingimm__init": [f"lingimm.init.{a}" for a in range(3, 11)],
                       "lingmat__init": [f"lingmat.init.{a}" for a in range(5, 16)]},
        extra_params={"x": 0.0})


def lib_path(name):
    return os.path.join(REF_DIR, f"libref_{name.lower()}.so")


def available(name):
    return os.path.exists(lib_path(name))


def build(native=False):
    """make -C oracle ref (needs /root/reference)."""
    subprocess.run(["make", "-C", _HERE, "ref"] + (["REF_CXXFLAGS=-O3 -march=native"] if native else []),
                   check=True, capture_output=True)


def _esc(s):
    return re.sub(r"\W", "__", s)          # cpp_escape_varname (R/run_tmb.R:5)


class RefModel:
    def __init__(self, name, model, params, extra_arrays=None, vector_params=None, extra_params=None):
        from gadget3_b200.run_cuda import _lgmatrix
        self.lib = C.CDLL(lib_path(name))
        L = self.lib
        L.g3ref_new.restype = C.c_void_p
        L.g3ref_free.argtypes = [C.c_void_p]
        L.g3ref_set_scalar.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.g3ref_set_param.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.g3ref_set_array.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int]
        L.g3ref_eval.argtypes = [C.c_void_p]
        L.g3ref_eval.restype = C.c_double
        L.g3ref_n_missing.argtypes = [C.c_void_p]
        L.g3ref_missing.argtypes = [C.c_void_p, C.c_int]
        L.g3ref_missing.restype = C.c_char_p
        L.g3ref_report.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.c_long, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.g3ref_report.restype = C.c_long
        self.h = C.c_void_p(L.g3ref_new())
        self.params = params
        self.model = model
        self.vector_params = vector_params or {}
        self.extra_params = extra_params or {}
        self._set_data(model, params, _lgmatrix)
        for k, v in (extra_arrays or {}).items():
            self._arr(k, v)

    def _arr(self, name, a, dims=None):
        a = np.asarray(a, dtype=np.float64)
        dims = list(a.shape) if dims is None else list(dims)
        flat = np.ascontiguousarray(a.ravel(order="F"))
        d = (C.c_int * len(dims))(*dims)
        self.lib.g3ref_set_array(self.h, name.encode(), flat.ctypes.data_as(C.POINTER(C.c_double)), d, len(dims))

    def _set_data(self, model, params, lgmatrix):
        self.lib.g3ref_set_scalar(self.h, b"reporting_enabled", 0.0)
        stocks = {}
        for a in model.actions:
            for attr in ("stock",):
                s = getattr(a, attr, None)
                if s is not None and getattr(s, "lengthgroups", None) is not None:
                    stocks[s.name] = s
            for s in list(getattr(a, "prey_stocks", [])) + list(getattr(a, "stocks", [])) + list(getattr(a, "output_stocks", [])):
                if getattr(s, "lengthgroups", None) is not None:
                    stocks[s.name] = s
        for n, s in stocks.items():
            self._arr(f"{n}__midlen", s.midlen)
        for a in model.actions:
            if a.kind == "predate" and a.catchability_f.kind == "totalfleet":
                E = a.catchability_f.E           # g3_timeareadata -> intlookup keys year*100 + step (R/timedata.R:113-149)
                keys = [k[0] * 100 + (k[1] or 0) for k in E.table]
                self._arr(E.name + "_keys", keys)
                self._arr(E.name + "_values", list(E.table.values()))
            if a.kind == "distribution":
                ld, nm = a.ld, a.nll_name
                nt = len(ld.times)
                for side in ("model", "obs"):       # R/stock_time.R:46-57: values are 1-based indices
                    self._arr(f"{nm}_{side}__times_keys", ld.times)
                    self._arr(f"{nm}_{side}__times_values", np.arange(1, nt + 1))
                # observation array in the reference's dimension order: length, [age], [stock], time, [area]
                obs = ld.obs.reshape(nt, ld.n_areag, ld.n_stockg, ld.n_age, ld.n_bins)
                dims, axes = [ld.n_bins], [4]
                if ld.minage is not None:
                    dims.append(ld.n_age); axes.append(3)
                if ld.stock_groups is not None:
                    dims.append(ld.n_stockg); axes.append(2)
                dims.append(nt); axes.append(0)
                if ld.area_group is not None:
                    dims.append(ld.n_areag); axes.append(1)
                rest = [ax for ax in range(5) if ax not in axes]
                arr = np.transpose(obs, axes + rest).reshape(dims)
                self._arr(f"{nm}_obs__{ld.unit}", arr)
                for p in a.stocks:
                    M = lgmatrix(p, ld.lenstock)
                    if M is None:
                        M = np.eye(p.L)
                    if a.fleets:
                        for f in a.fleets:
                            self._arr(f"{p.name}_{f.name}_{nm}_model_lgmatrix", M)
                    else:
                        self._arr(f"{p.name}_{nm}_model_lgmatrix", M)
        for i, sw in enumerate(params.switch):
            self.lib.g3ref_set_scalar(self.h, (_esc(sw) + "__lower").encode(), float(params.lower[i]))
            self.lib.g3ref_set_scalar(self.h, (_esc(sw) + "__upper").encode(), float(params.upper[i]))

    def eval(self, full_par, reporting=False):
        self.lib.g3ref_set_scalar(self.h, b"reporting_enabled", 1.0 if reporting else 0.0)
        full_par = np.asarray(full_par, dtype=np.float64)
        for sw, v in zip(self.params.switch, full_par):
            self.lib.g3ref_set_param(self.h, _esc(sw).encode(), float(v))
        for vname, switches in self.vector_params.items():          # PARAMETER_VECTOR
            self._arr(vname, [full_par[self.params.index(sw)] for sw in switches])
        for k, v in self.extra_params.items():
            self.lib.g3ref_set_param(self.h, k.encode(), float(v))
        nll = self.lib.g3ref_eval(self.h)
        nm = self.lib.g3ref_n_missing(self.h)
        if nm:
            raise RuntimeError("reference objective asked for undeclared inputs: "
                               + ", ".join(self.lib.g3ref_missing(self.h, i).decode() for i in range(min(nm, 12))))
        return nll

    def eval_batch(self, full):
        return np.array([self.eval(row) for row in np.atleast_2d(full)])

    def report(self, name):
        dim = (C.c_int * 8)()
        nd = C.c_int(0)
        n = self.lib.g3ref_report(self.h, name.encode(), None, 0, dim, C.byref(nd))
        if n == 0:
            return None
        out = np.empty(n)
        self.lib.g3ref_report(self.h, name.encode(), out.ctypes.data_as(C.POINTER(C.c_double)), n, dim, C.byref(nd))
        return out.reshape([dim[i] for i in range(nd.value)], order="F")

    def __del__(self):
        try:
            self.lib.g3ref_free(self.h)
        except Exception:
            pass
