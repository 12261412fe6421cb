"""The oracle against the REFERENCE'S OWN generated objectives.

baseline/introduction-single-stock.cpp, baseline/multiple-substocks.cpp and inttest/codegeneration/ling.cpp
(the reference's checked-in g3_to_tmb() output) are compiled unmodified against oracle/tmb_shim and run on
our C1/C2/LING model data (LING = configs.build_ling, the mirror of inttest/codegeneration/run.R:17-183:
constant maturity, 15-group growth, two renewals, age -> mature movement, report = FALSE likelihood):
  * tests/golden/reference_{c1,c2,ling}.npz hold their nll and REPORT()ed arrays for seeded vectors
    (tools/make_golden_reference.py) -- checked everywhere, including where /root/reference is absent;
  * where oracle/_ref was built (build container), fresh random vectors are compared live.
This pins the whole-model arithmetic of the oracle (every action of SURVEY section 8a that C1/C2
exercise) to the reference's code, not to a re-reading of it.
"""
import numpy as np
import pytest

from _util import US_WEIGHT, check_detail_reports, golden, us_tolerance
from gadget3_b200.configs import CONFIGS, parameter_batch
from gadget3_b200.run_cuda import unpack_report
from oracle.oracle_lib import Oracle


@pytest.fixture(scope="module")
def setups(oracle_lib):
    out = {}
    for n in ("C1", "C2", "LING"):
        model, p = CONFIGS[n]()
        out[n] = (model, p, Oracle(*model.pack(p)))
    return out


@pytest.mark.parametrize("name", ["C1", "C2", "LING"])
def test_oracle_nll_matches_reference_code(setups, name):
    model, p, orc = setups[name]
    g = golden(name)
    assert list(g["switch"]) == list(p.switch)
    nll, comp, _ = orc.eval_batch(g["par"], want_comp=True, nthreads=4)
    tol = 1e-11 * np.abs(g["nll"]) + US_WEIGHT * us_tolerance(model, p, comp[:, -2])
    assert (np.abs(nll - g["nll"]) <= tol).all()
    assert np.median(np.abs(nll - g["nll"]) / np.abs(g["nll"])) < 1e-14


@pytest.mark.parametrize("vi", [0, 5])
@pytest.mark.parametrize("name", ["C1", "C2"])
def test_oracle_reports_match_reference_code(setups, name, vi):
    """Per-step nll breakdowns, model arrays, regression parameters and stock state histories."""
    model, p, orc = setups[name]
    g = golden(name)
    rep = unpack_report(model, orc.report(g["par"][vi]))
    assert rep["nll"] == pytest.approx(float(g[f"v{vi}__nll"][0]), rel=1e-11)
    for cn, unit, sh in zip(model.comp_names, model.comp_units, model.comp_shapes):
        np.testing.assert_allclose(rep[f"nll_{cn}__{unit}"], g[f"v{vi}__nll_{cn}__{unit}"], rtol=1e-10, atol=1e-300)
        ours = rep[f"{cn}_model__{unit}"]                       # [time][areag][stockg][age][bin]
        ref = g[f"v{vi}__{cn}_model__{unit}"]                   # length, [age], [stock], time, [area]
        nt = sh["n_times"]
        axes, dims = [4], [sh["n_bins"]]
        if ref.ndim >= 3 and sh["n_age"] > 1:
            axes.append(3); dims.append(sh["n_age"])
        if sh["n_stockg"] > 1:
            axes.append(2); dims.append(sh["n_stockg"])
        axes.append(0); dims.append(nt)
        rest = [a for a in range(5) if a not in axes]
        mine = np.transpose(ours, axes + rest).reshape(ref.shape)
        np.testing.assert_allclose(mine, ref, rtol=1e-10, atol=1e-12 * np.abs(ref).max())
        if f"v{vi}__{cn}_model__params" in g.files:
            np.testing.assert_allclose(rep[f"{cn}_model__params"], g[f"v{vi}__{cn}_model__params"], rtol=1e-10)
    us_ref = g[f"v{vi}__nll_understocking__wgt"]
    assert np.abs(rep["nll_understocking__wgt"] - us_ref).sum() <= us_tolerance(model, p, us_ref.sum())
    times = g["report_times"]
    for sn in model.stock_names:
        for what in ("num", "wgt"):
            ref = g[f"v{vi}__dstart_{sn}__{what}"]
            mine = rep[f"dstart_{sn}__{what}"][..., times]
            np.testing.assert_allclose(mine, ref, rtol=1e-10, atol=1e-300)
    check_detail_reports(model, rep, g, vi)


@pytest.mark.parametrize("vi", [0, 5])
def test_oracle_ling_understocking_matches_reference_code(setups, vi):
    """ling.cpp REPORT()s (besides two suitability vectors) only nll_understocking__wgt per step."""
    model, p, orc = setups["LING"]
    g = golden("LING")
    rep = unpack_report(model, orc.report(g["par"][vi]))
    assert rep["nll"] == pytest.approx(float(g["nll"][vi]), rel=1e-11)
    us_ref = g[f"v{vi}__nll_understocking__wgt"]
    assert np.abs(rep["nll_understocking__wgt"] - us_ref).sum() <= us_tolerance(model, p, us_ref.sum())
    assert check_detail_reports(model, rep, g, vi) == 2          # the two suit_*__report vectors


@pytest.mark.parametrize("name", ["C1", "C2", "LING"])
def test_live_reference_when_built(setups, name):
    from oracle import ref_lib
    if not ref_lib.available(name):
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    model, p, orc = setups[name]
    ref = ref_lib.RefModel(name, model, p, **(ref_lib.ling_extras(p) if name == "LING" else {}))
    full = np.tile(p.values(), (12, 1))
    full[:, p.optimised_indices()] = parameter_batch(p, 12, seed=777)
    o, comp, _ = orc.eval_batch(full, want_comp=True)
    r = ref.eval_batch(full)
    tol = 1e-11 * np.abs(r) + US_WEIGHT * us_tolerance(model, p, comp[:, -2])
    assert (np.abs(o - r) <= tol).all()
