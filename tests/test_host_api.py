"""Host-side logic: the g3a_*/g3l_* mirror, g3_to_cuda() lowering and the C ABI surface. No GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gadget3_b200 as g3
from gadget3_b200 import _lib
from gadget3_b200.configs import CONFIGS, parameter_batch
from gadget3_b200.spec import K

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _golden_list(name):
    with open(os.path.join(GOLD, name)) as fh:
        return [ln.strip() for ln in fh if ln.strip() and not ln.startswith("#")]


@pytest.fixture(scope="module")
def cfgs():
    return {n: CONFIGS[n]() for n in ("C1", "C2")}


@pytest.mark.parametrize("name,gold", [("C1", "c1"), ("C2", "c2")])
def test_parameter_order_matches_reference(cfgs, name, gold):
    """PARAMETER() order of the reference's generated objective (baseline/*.cpp) == template order."""
    model, params = cfgs[name]
    ours = [re.sub(r"\W", "__", s) for s in params.switch]         # cpp_escape_varname, R/run_tmb.R:5
    assert ours == _golden_list(f"parameter_order_{gold}.txt")


@pytest.mark.parametrize("name,gold", [("C1", "c1"), ("C2", "c2")])
def test_bounds_penalty_order_matches_reference(cfgs, name, gold):
    """g3l_bounds_penalty blocks run in C-locale name order (baseline/introduction-single-stock.cpp:804-870);
    nll is a floating-point sum, so this is the summation order."""
    model, params = cfgs[name]
    ints, reals = model.pack(params)
    nb, off = ints[K["G3H_NBOUND"]], ints[K["G3H_BOUND_OFF"]]
    ours = [params.switch[i] for i in ints[off:off + nb]]
    gold_order = _golden_list(f"bounds_order_{gold}.txt")
    assert set(ours) <= set(gold_order)
    assert ours == [n for n in gold_order if n in set(ours)]
    # every bounded parameter is finite-bounded in the template
    for n in ours:
        i = params.index(n)
        assert np.isfinite(params.lower[i]) and np.isfinite(params.upper[i])


def test_model_dimensions(cfgs):
    m1, _ = cfgs["C1"]
    m2, _ = cfgs["C2"]
    assert m1.stock_names == ["fish"] and m1.stock_shapes == [(10, 5, 1)]
    assert m1.stock_dims[0] == ["length", "area", "age"]          # baseline/introduction-single-stock.cpp:470
    assert m2.stock_names == ["fish_imm", "fish_mat"] and m2.stock_shapes == [(20, 5, 1), (20, 8, 1)]
    assert m2.stock_dims[0] == ["length", "age", "area"]          # baseline/multiple-substocks.cpp:571
    assert m1.n_steps == m2.n_steps == 136                         # total_steps + 1
    # likelihood components in generated-name order
    assert m2.comp_names == ["adist_surveyindices_log_dist_si_cpue", "cdist_sumofsquares_aldist_f_surv",
                             "cdist_sumofsquares_ldist_f_surv", "cdist_sumofsquares_matp_f_surv"]
    sh = dict(zip(m2.comp_names, m2.comp_shapes))
    assert sh["cdist_sumofsquares_ldist_f_surv"]["n_bins"] == 23
    assert sh["cdist_sumofsquares_aldist_f_surv"]["n_age"] == 10
    assert sh["cdist_sumofsquares_matp_f_surv"]["n_stockg"] == 2


def test_length_groups():
    """g3s_length: midlen / plusdl (R/stock.R:68-112)."""
    s = g3.g3_stock("fish", range(10, 101, 10))
    assert list(s.midlen) == [15, 25, 35, 45, 55, 65, 75, 85, 95, 105]
    assert s.plusdl == 10
    s = g3.g3_stock("x", [5, 10, 15, 25])
    assert s.plusdl == 5 and list(s.midlen) == [7.5, 12.5, 20.0, 27.5]


def test_init_val_wildcards(cfgs):
    _, p = cfgs["C2"]
    i = p.index("fish_mat.M.7")
    assert p.value[i] == 0.15 and p.lower[i] == 0.001 and p.upper[i] == 10 and p.optimise[i]
    assert not p.optimise[p.index("fish_imm.walpha")]
    assert p.value[p.index("fish_imm.rec.scalar")] == 1000 and not p.optimise[p.index("fish_imm.rec.scalar")]
    assert len(p.optimised_indices()) == 71                        # vignettes/multiple-substocks.Rmd:421-439


def test_parameter_batch_inside_bounds(cfgs):
    _, p = cfgs["C2"]
    P = parameter_batch(p, 128)
    idx = p.optimised_indices()
    lo, up = np.asarray(p.lower)[idx], np.asarray(p.upper)[idx]
    assert P.shape == (128, 71)
    assert (P >= lo).all() and (P <= up).all()
    assert np.ptp(P, axis=0).min() > 0


def test_unsupported_actions_raise():
    """Outside the supported subset g3_to_cuda() stops -- there is no CPU fallback."""
    st = g3.g3s_livesonareas(g3.g3s_age(g3.g3_stock("s", range(10, 51, 10)), 1, 3), {"a": 1})
    with pytest.raises(g3.G3Unsupported):
        g3.g3a_mature_continuous(bounded=False)
    with pytest.raises(g3.G3Unsupported):
        g3.g3a_naturalmortality(st, mortality_f=0.2)
    with pytest.raises(ValueError):
        g3.g3a_time(2000, 2001, [3, 3, 3])
    with pytest.raises(ValueError):
        g3.g3_to_cuda([g3.g3a_naturalmortality(st)])                 # no g3a_time


def test_predator_length_suitabilities_need_a_stock_predator():
    """g3_suitability_exponential / richards / andersen read predator_length (R/suitability.R:15-30,69-94): a fleet
    has none, so the lowering refuses them there, as the reference fails on the undefined variable."""
    areas = {"a": 1}
    st = g3.g3s_livesonareas(g3.g3s_age(g3.g3_stock("s", range(10, 51, 10)), 1, 3), areas)
    fl = g3.g3s_livesonareas(g3.g3_fleet("f"), areas)
    land = g3.g3_timeareadata("l", dict(year=[2000], step=[1], area=["a"], w=[1.0]), "w", areas=areas)
    for su in (g3.g3_suitability_richards(1, 1, 1, 1, 1), g3.g3_suitability_exponential(1, 1, 1, 1),
               g3.g3_suitability_andersen(0, 1, 1, 1, 1)):
        acts = [g3.g3a_time(2000, 2001, [6, 6]), g3.g3a_naturalmortality(st), g3.g3a_initialconditions_normalcv(st),
                g3.g3a_predate_fleet(fl, [st], suitabilities=su, catchability_f=g3.g3a_predate_catchability_linearfleet(land))]
        with pytest.raises(g3.G3Unsupported, match="stock predator"):
            g3.g3_to_cuda(acts)
    with pytest.raises(g3.G3Unsupported, match="g3_timeareadata"):
        g3.g3a_predate_catchability_effortfleet(0.1, 5.0)


def test_pack_rejects_reordered_template(cfgs):
    """R/run_tmb.R:1156-1158: 'Parameters not in expected order'."""
    model, p = cfgs["C1"]
    q = p.copy()
    q.switch[1], q.switch[2] = q.switch[2], q.switch[1]
    with pytest.raises(ValueError, match="expected order"):
        model.pack(q)


def test_spec_is_self_consistent(cfgs):
    for name in ("C1", "C2"):
        model, p = cfgs[name]
        ints, reals = model.pack(p)
        assert ints[K["G3H_VERSION"]] == K["G3CUDA_SPEC_VERSION"]
        assert ints[K["G3H_NPAR"]] == len(p.switch)
        assert ints[K["G3H_NNLL"]] == len(model.nll_names)
        assert ints.max() < max(len(ints), len(reals), 1 << 20)
        assert np.isfinite(reals[np.isfinite(reals)]).all()


# ------------------------------------------------------------------------------ C ABI surface
def _declared_functions():
    src = open(os.path.join(ROOT, "include", "g3cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(g3cuda_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    lib = C.CDLL(_lib.LIB_PATH)
    declared = _declared_functions()
    assert len(declared) >= 13
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/g3cuda.h but not exported"
    assert sorted(_lib.EXPORTED) == declared
    lib.g3cuda_abi_version.restype = C.c_int32
    assert lib.g3cuda_abi_version() == K["G3CUDA_SPEC_VERSION"]


def test_no_device_fails_loudly(cfgs):
    """Without a CUDA device every compute entry point fails (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    model, p = cfgs["C1"]
    with pytest.raises(_lib.G3CudaError, match="no usable CUDA device|CUDA"):
        g3.g3_cuda_adfun(model, p)


def test_malformed_spec_is_rejected():
    lib = _lib.load()
    bad = _lib._Spec(999, 0, None, 0, None)
    h = C.c_void_p()
    assert lib.g3cuda_model_create(C.byref(bad), 0, C.byref(h)) < 0
    assert lib.g3cuda_last_error()


def test_bounds_penalty_parameter_template_form_refused():
    """R/likelihood_bounds.R:44-62: the parameter_template form penalises optimise = TRUE rows with the bounds of the moment
    baked in -- not what this backend computes, so it refuses instead of silently returning a different nll."""
    from gadget3_b200 import g3l_bounds_penalty
    from gadget3_b200.configs import CONFIGS
    from gadget3_b200.formula import G3Unsupported
    _, p = CONFIGS["C1"]()
    with pytest.raises(G3Unsupported):
        g3l_bounds_penalty(p)
