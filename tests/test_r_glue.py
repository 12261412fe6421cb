"""The R-facing glue exists as files (R/run_cuda.R, src/g3cuda_R.c). R is not in the build image, so what can be checked
here: the C shim compiles (syntax and types) against a minimal stand-in for R's C API (tests/r_stub/) and the real
include/g3cuda.h, it only calls entry points the library exports, and the R file names the .Call routines the shim registers."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_compiles_against_the_abi_header():
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I" + os.path.join(ROOT, "tests", "r_stub"),
                        "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "src", "g3cuda_R.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_shim_calls_only_exported_entry_points():
    from gadget3_b200._lib import EXPORTED
    src = open(os.path.join(ROOT, "src", "g3cuda_R.c")).read()
    used = set(re.findall(r"\b(g3cuda_[a-z_]+)\s*\(", src)) - {"g3cuda_R_finalize", "g3cuda_R_create", "g3cuda_R_dims",
                                                                   "g3cuda_R_eval", "g3cuda_R_report"}
    assert used and used <= set(EXPORTED), used - set(EXPORTED)


def test_r_side_uses_the_registered_routines():
    shim = open(os.path.join(ROOT, "src", "g3cuda_R.c")).read()
    registered = set(re.findall(r'\{"(g3cuda_R_[a-z]+)"', shim))
    rsrc = open(os.path.join(ROOT, "R", "run_cuda.R")).read()
    called = set(re.findall(r'\.Call\("(g3cuda_R_[a-z]+)"', rsrc))
    assert called and called <= registered
    for fn in ("g3_to_cuda", "g3_cuda_adfun", "ut_cuda_tmb_compare", "g3_cuda_unpack_report"):
        assert re.search(r"^%s <- function" % fn, rsrc, re.M), fn


def test_save_spec_round_trip(tmp_path):
    """The files the R loader reads: the packed spec bit for bit, the template, and a report layout whose index vectors
    rebuild every report array from the flat report vector."""
    import json
    import numpy as np
    from gadget3_b200.configs import CONFIGS
    from gadget3_b200.run_cuda import save_spec, unpack_report
    from oracle.oracle_lib import Oracle
    model, p = CONFIGS["C1"]()
    ints, reals = model.pack(p)
    raw = Oracle(ints, reals).report(p.values())
    meta = save_spec(model, tmp_path / "c1", p, report_len=len(raw))
    assert (np.fromfile(str(tmp_path / "c1.ints"), dtype=np.int32) == np.asarray(ints)).all()
    assert np.fromfile(str(tmp_path / "c1.reals"), dtype=np.float64).tobytes() == np.asarray(reals, dtype=np.float64).tobytes()
    assert json.load(open(tmp_path / "c1.json"))["parameter_template"]["switch"] == list(p.switch)
    idx = np.fromfile(str(tmp_path / "c1.idx"), dtype=np.int32)
    want = unpack_report(model, raw)
    for e in meta["report_layout"]:
        n = int(np.prod(e["dim"]))
        at = idx[e["idx_offset"]:e["idx_offset"] + n]
        got = np.where(at < 0, np.nan, raw[np.maximum(at, 0)]).reshape(e["dim"], order="F")
        ref = np.asarray(want[e["name"]], dtype=float).reshape(e["dim"])
        assert ((got == ref) | (np.isnan(got) & np.isnan(ref))).all(), e["name"]
