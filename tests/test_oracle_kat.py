"""Pin the CPU oracle against the known-answer vectors held by the reference's own unit tests.

Every expected number below is quoted from a reference test (file:line given per test); none is
produced by this repository. The oracle (oracle/g3oracle.cpp) must reproduce them before any CUDA
parity claim that leans on it means anything.
"""
import ctypes as C
import math

import numpy as np
import pytest


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.fixture(scope="module")
def lib(oracle_lib):
    return oracle_lib.load()


def _bbinom_len(lib, midlen, plusdl, linf, kappa, beta, n, dt=1.0):
    """growth_bbinom(avoid_zero(lengthvbsimple / plusdl), n, avoid_zero(beta)) -- R/action_grow.R:214-223."""
    L = len(midlen)
    dl = np.array([lib.g3oracle_avoid_zero(lib.g3oracle_grow_lengthvbsimple(linf, kappa, m, dt) / plusdl)
                   for m in midlen])
    out = np.empty(L * (n + 1))
    lib.g3oracle_growth_bbinom(_dp(dl), L, n, lib.g3oracle_avoid_zero(beta), _dp(out))
    return out.reshape(n + 1, L).T          # column-major L x (n+1)


def test_env_primitives(lib):
    # R/aab_env.R:16-31: avoid_zero(0) = logspace_add(0, 0) / 1000 = log(2) / 1000, not 0
    assert lib.g3oracle_logspace_add(0.0, 0.0) == pytest.approx(math.log(2.0), rel=1e-15)
    assert lib.g3oracle_avoid_zero(0.0) == pytest.approx(math.log(2.0) / 1000.0, rel=1e-15)
    assert lib.g3oracle_avoid_zero(5.0) == 5.0
    assert lib.g3oracle_avoid_zero(-5.0) == pytest.approx(0.0, abs=1e-300)
    # R/env_dif.R:37-47 dif_pmin: smooth minimum
    assert lib.g3oracle_dif_pmin(0.2, 0.95, 1e3) == pytest.approx(0.2, rel=1e-15)
    assert lib.g3oracle_dif_pmin(5.0, 0.95, 1e3) == pytest.approx(0.95, rel=1e-15)
    assert lib.g3oracle_dif_pmin(0.95, 0.95, 1e3) == pytest.approx(0.95 - math.log(2.0) / 1e3, rel=1e-14)
    # R/env_dif.R:49-67
    assert lib.g3oracle_dif_pminmax(0.0, 0.0, 1.0, 1e5) == pytest.approx(math.log(2.0) / 1e5, rel=1e-10)
    # R/aab_env.R:133-153
    assert lib.g3oracle_ratio_add_pop(2.0, 10.0, 4.0, 30.0) == pytest.approx((2 * 10 + 4 * 30) / 40.0, rel=1e-15)
    # TMB dnorm == R dnorm
    assert lib.g3oracle_dnorm(1.3, 0.7, 2.1) == pytest.approx(
        math.exp(-0.5 * ((1.3 - 0.7) / 2.1) ** 2) / (2.1 * math.sqrt(2 * math.pi)), rel=1e-14)


def test_grow_lengthvbsimple(lib):
    """tests/test-action_grow.R:9-23: g3_stock("s", seq(10, 100, 10) - 5), kappa = 100, cur_step_size = 1."""
    midlen = np.arange(10, 101, 10, dtype=float)
    got = np.array([lib.g3oracle_grow_lengthvbsimple(150.0, 100.0, m, 1.0) for m in midlen])
    np.testing.assert_allclose(got, [140, 130, 120, 110, 100, 90, 80, 70, 60, 50], rtol=1.5e-8)
    got = np.array([lib.g3oracle_grow_lengthvbsimple(55.0, 100.0, m, 1.0) for m in midlen])
    np.testing.assert_allclose(got, [45, 35, 25, 15, 5, 0, 0, 0, 0, 0], atol=1e-7)


GIM_120 = """
  8.29  8.44  0.40  0.17 0.12 0.15
  9.00 10.00  0.00  0.00 0.00 0.00
 11.43 14.29  1.02  0.36 0.22 0.26
 26.56 39.35  7.50  2.00 1.11 1.18
 26.18 50.91 21.82  2.18 0.91 0.82
 21.00 70.00 80.00 32.00 0.00 0.00
  0.00  0.00  0.00  0.00 0.00 1.00
  0.02  0.05  0.08  0.13 0.21 0.51
  0.14  0.12  0.12  0.13 0.17 0.32
  0.37  0.14  0.10  0.10 0.11 0.19
"""
GIM_40 = """
  0.14 0.12 0.12 0.13 0.17 0.32
  0.37 0.14 0.10 0.10 0.11 0.19
  0.66 0.09 0.06 0.05 0.05 0.09
  1.00 0.00 0.00 0.00 0.00 0.00
  1.00 0.00 0.00 0.00 0.00 0.00
  1.00 0.00 0.00 0.00 0.00 0.00
  1.00 0.00 0.00 0.00 0.00 0.00
  1.00 0.00 0.00 0.00 0.00 0.00
  1.00 0.00 0.00 0.00 0.00 0.00
  1.00 0.00 0.00 0.00 0.00 0.00
"""


def _mat(txt):
    return np.array([[float(x) for x in ln.split()] for ln in txt.strip().splitlines()])


@pytest.mark.parametrize("linf,expected", [(120.0, GIM_120), (40.0, GIM_40)])
def test_growth_bbinom_tables(lib, linf, expected):
    """tests/test-action_grow.R:25-63: kappa = 20, beta = 0.5, maxlengthgroupgrowth = 5, 2 decimals.
    Includes rows where the mean jump exceeds n (alpha < 0): lgamma of negative arguments."""
    midlen = np.arange(10, 101, 10, dtype=float)
    got = _bbinom_len(lib, midlen, 10.0, linf, 20.0, 0.5, 5)
    np.testing.assert_array_equal(np.round(got, 2), _mat(expected))
    if linf == 40.0:
        np.testing.assert_allclose(got.sum(axis=1), np.ones(10), rtol=1e-2)


def test_growth_l_baseline(lib):
    """tests/test-action_grow.R:96-105: teststock = seq(10, 35, 5), linf 160, kappa 90, beta 30,
    maxlengthgroupgrowth 4, one 12-month step; teststock__growth_l at tolerance 1e-5."""
    expected = np.array([
        12199.8976226175, 9352.23228375502, 7157.83523789257, 5463.49450549465, 4154.31593595054, 3143.22179218075,
        51322.2074676947, 39560.4630926918, 30458.8733527337, 23399.2087912083, 17917.1342692155, 13660.121314133,
        81087.2009530949, 62860.2639001552, 48695.7187344743, 37658.1016483506, 29043.7267350904, 22317.4314305221,
        57032.8699935755, 44472.5676572519, 34669.6583623158, 26995.0549450546, 20974.8144063369, 16247.8860873661,
        15068.9788855573, 11821.5345660349, 9275.97774268346, 7273.66758241751, 5694.90600451131, 4448.35417879727,
    ]).reshape(5, 6).T
    midlen = np.arange(10, 36, 5, dtype=float) + 2.5
    got = _bbinom_len(lib, midlen, 5.0, 160.0, 90.0, 30.0, 4)
    np.testing.assert_allclose(got, expected, rtol=1e-5)


def test_grow_apply_handmade_matrix(lib):
    """tests/test-action_grow.R:121-182: hand-made 6 x 7 growth matrix, zero weight deltas;
    plus group cannot be escaped; 0 / avoid_zero(0) stays finite."""
    gm = np.array([
        0, 0.25, 1, 0.5, 0.5, 1,
        0, 0, 0, 0, 0, 0,
        1, 0.75, 0, 0, 0, 0,
        0, 0, 0, 0, 0, 0,
        0, 0, 0, 0, 0.5, 0,
        0, 0, 0, 0, 0, 0,
        0, 0, 0, 0.5, 0, 0], dtype=float)            # column-major 6 x 7
    L, D = 6, 7
    G = np.empty(L * L)
    W = np.empty(L * L)
    lib.g3oracle_grow_matrix_len(_dp(gm), L, D, _dp(G))
    wm = np.zeros(L * D)
    lib.g3oracle_grow_matrix_wgt(_dp(wm), L, D, _dp(W))
    num = np.array([10, 100, 1000, 1000, 10000, 100000], dtype=float)
    wgt = np.array([100, 200, 300, 400, 500, 600], dtype=float)
    on, ow = np.empty(L), np.empty(L)
    lib.g3oracle_grow_apply(_dp(G), _dp(W), L, _dp(num), _dp(wgt), _dp(on), _dp(ow))
    assert on.tolist() == [0, 25, 1010, 575, 5000, 105500]            # ut_cmp_identical
    lsa00 = lib.g3oracle_logspace_add(0.0, 0.0)
    expected_w = [0 / lsa00,
                  (200 * 100 * 0.25) / 25,
                  (300 * 1000 + 100 * 10) / 1010,
                  (400 * 1000 * 0.5 + 200 * 100 * 0.75) / 575,
                  (500 * 10000 * 0.5) / 5000,
                  (600 * 100000 + 500 * 10000 * 0.5 + 400 * 1000 * 0.5) / 105500]
    np.testing.assert_allclose(ow, expected_w, rtol=1e-5)


def test_digamma_matches_lgamma_slope(lib):
    # digamma drives the lgamma tangent of the gradient oracle; check against a central difference
    for x in (0.3, 1.0, 2.5, 17.0, 1500.0):
        h = 1e-6 * max(1.0, x)
        fd = (math.lgamma(x + h) - math.lgamma(x - h)) / (2 * h)
        assert lib.g3oracle_digamma(x) == pytest.approx(fd, rel=1e-6)


def test_suitability_andersen(lib):
    """tests/test-eval.R:5-21: g3_suitability_andersen(0, 1, 2, 3, 4), predator_length = 99,
    prey g3_stock('prey', 1:10 * 10)."""
    expected = [1.5385642245386, 1.90781891144376, 1.99894574784489, 1.97774955212781, 1.91681934676797,
                1.83906905397844, 1.75539377253919, 1.67124659054669, 1.58937906439558, 1.51113553695681]
    midlen = np.arange(10, 101, 10, dtype=float) + 5.0
    got = [lib.g3oracle_suit_andersen(0.0, 1.0, 2.0, 3.0, 4.0, 99.0, m) for m in midlen]
    np.testing.assert_allclose(got, expected, rtol=1.5e-8)


@pytest.mark.parametrize("vec,src,row_total,expected", [
    ([np.sqrt(0.2), 0, np.sqrt(0.2)], 0, 1.0, [0.8, 0, 0.2]),
    ([0, np.sqrt(0.2), 0], 1, 1.0, [0, 1, 0]),
    ([np.sqrt(0.2), 0, 0], 2, 1.0, [0.2, 0, 0.8]),
    ([np.sqrt(0.2), 0, np.sqrt(0.2)], 0, 10.0, [0.98, 0, 0.02]),
])
def test_migrate_normalize(lib, vec, src, row_total, expected):
    """tests/test-action_migrate.R:19-31 (the selected column of array(c(sqrt(0.2), 0[, 0]), dim = c(3,3)))."""
    v = np.array(vec, dtype=float)
    lib.g3oracle_migrate_normalize(_dp(v), 3, src, row_total)
    np.testing.assert_allclose(v, expected, rtol=1e-12, atol=1e-15)


def _suit(lib, kind, par, ml, pl=0.0):
    p = np.zeros(5)
    p[:len(par)] = par
    return lib.g3oracle_suitability(kind, _dp(p), float(ml), float(pl))


@pytest.mark.parametrize("p3_exp,p4_exp", [(0.1, 0.1), (-np.inf, 0.1), (0.1, 0.999)])
def test_suitability_andersenfleet(lib, p3_exp, p4_exp):
    """tests/test-suitability.R:62-108: lg = seq(100, 500, 100), defaults p0 = 0, p1 = log(2), p2 = 1,
    p3 = exp(p3_exp), p4 = exp(p4_exp), p5 = largest midlen; against the test's non-differentiable form, 1e-6."""
    from gadget3_b200.spec import K
    lg = np.arange(100.0, 501.0, 100.0)
    midlen = lg + 50.0
    p0, p1, p2, p3, p4, p5 = 0.0, math.log(2.0), 1.0, math.exp(p3_exp), math.exp(p4_exp), midlen.max()
    x = np.log(p5 / midlen)
    with np.errstate(divide="ignore", invalid="ignore"):
        want = p0 + p2 * np.where(x <= p1, np.exp(-(x - p1) ** 2 / p4), np.exp(-(x - p1) ** 2 / p3))
    got = [_suit(lib, K["G3SUIT_ANDERSENFLEET"], [p0, p1, p2, p3, p4], m, p5) for m in midlen]
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-12)


def test_suitability_closed_forms(lib):
    """R/suitability.R:60-94: gamma, exponential, straightline, richards against the formulas as documented."""
    from gadget3_b200.spec import K
    for ml in (12.5, 40.0, 87.5):
        a, b, g = 3.2, 4.0, 2.5
        assert _suit(lib, K["G3SUIT_GAMMA"], [a, b, g], ml) == pytest.approx(
            (ml / ((a - 1) * b * g)) ** (a - 1) * math.exp(a - 1 - ml / (b * g)), rel=1e-13)
        a, b, g, d, pl = -3.0, 0.08, 0.01, 0.9, 55.0
        e = d / (1 + math.exp(-a - b * ml - g * pl))
        assert _suit(lib, K["G3SUIT_EXPONENTIAL"], [a, b, g, d], ml, pl) == pytest.approx(e, rel=1e-14)
        assert _suit(lib, K["G3SUIT_RICHARDS"], [a, b, g, d, 1.7], ml, pl) == pytest.approx(e ** (1 / 1.7), rel=1e-13)
        assert _suit(lib, K["G3SUIT_STRAIGHTLINE"], [0.1, 0.004], ml) == pytest.approx(0.1 + 0.004 * ml, rel=1e-15)


def test_mature_continuous_age_term_hand_value(oracle_lib):
    """g3a_mature_continuous with its age term (R/action_mature.R:1-42): per step the share
    dif_pminmax(m0 * (alpha * ldelta + beta * cur_step_size), 0, 1, 1e5), m0 = 1 / (1 + exp(-alpha (l - l50) - beta (age - a50)))
    of every length jump matures. With alpha = 0 that share no longer depends on length or jump, so -- growth keeps every
    fish (plus group), mortality switched off, no landings and no ageing inside the first steps of the year -- an age
    column of the immature stock that has a destination loses exactly that share of its fish per step, and a column
    without one (ages 1-2: mat starts at age 3) gets the remainder back and loses nothing. The expected numbers are
    computed here from the formula, not by the oracle."""
    from gadget3_b200.configs import CONFIGS
    from gadget3_b200.run_cuda import unpack_report
    from oracle.oracle_lib import Oracle
    model, p = CONFIGS["C2_MATAGE"]()
    v = p.values().copy()
    beta, a50, dt = 0.4, 3.0, 0.25
    v[p.index("fish_imm.mat.alpha")] = 0.0
    v[p.index("fish_imm.mat.beta")] = beta
    v[p.index("fish_imm.mat.a50")] = a50
    for a in range(1, 6):
        v[p.index(f"fish_imm.M.{a}")] = 0.0
    rep = unpack_report(model, Oracle(*model.pack(p)).report(v))
    num = np.asarray(rep["dstart_fish_imm__num"])          # length x age x area x time
    tot = num.sum(axis=0)[:, 0, :]                         # age x time
    for ai, age in enumerate(range(1, 6)):
        m0 = 1.0 / (1.0 + math.exp(-beta * (age - a50)))
        share = min(max(m0 * beta * dt, 0.0), 1.0) if age >= 3 else 0.0
        if age == 1:
            continue            # (the renewal lands in age 1 at step 1)
        # the first step of the run: no landings (they come at step 2), no ageing (step 4)
        assert tot[ai, 1] / tot[ai, 0] == pytest.approx(1.0 - share, rel=1e-9), age


def test_migrate_twelve_step_recurrence(oracle_lib):
    """tests/test-action_migrate.R:63-102: a stock on areas a, c, d (not b) with a spring migration d -> a (0.4 = the
    parameter squared) on step 2 and a winter migration c -> d, a -> c (0.6) on step 4; the test applies the moves by hand,
    numbers and ratio_add_pop weights, step after step, and compares every step with the model. Same recurrence here on the
    oracle's state at the start of the following step, for every (length, age) -- the reference checks one pair and that the
    ages agree."""
    from gadget3_b200.configs import CONFIGS
    from gadget3_b200.run_cuda import unpack_report
    from oracle.oracle_lib import Oracle
    model, p = CONFIGS["MINI_MIGRATE"]()
    rep = unpack_report(model, Oracle(*model.pack(p)).report(p.values()))
    num, wgt = np.asarray(rep["dstart_stock_acd__num"]), np.asarray(rep["dstart_stock_acd__wgt"])     # [length][age][area a,c,d][t]
    assert num.shape == (4, 5, 3, 20)
    rap = lambda ov, oa, nv, na: (ov * oa + nv * na) / (oa + na)
    a, c, d = 0, 1, 2
    en, ew = num[..., 0].copy(), wgt[..., 0].copy()
    assert en.min() > 0 and len({float(x) for x in ew[0, 0]}) == 3          # weights differ by area: the mixing is visible
    for t in range(19):
        if t % 4 == 1:          # spring
            ew[:, :, a] = rap(ew[:, :, a], en[:, :, a], ew[:, :, d], en[:, :, d] * 0.4)
            en[:, :, a] = en[:, :, a] + en[:, :, d] * 0.4
            en[:, :, d] = en[:, :, d] - en[:, :, d] * 0.4
        if t % 4 == 3:          # winter: c -> d, then a -> c; nothing goes a -> d directly
            ew[:, :, d] = rap(ew[:, :, d], en[:, :, d], ew[:, :, c], en[:, :, c] * 0.6)
            en[:, :, d] = en[:, :, d] + en[:, :, c] * 0.6
            en[:, :, c] = en[:, :, c] - en[:, :, c] * 0.6
            ew[:, :, c] = rap(ew[:, :, c], en[:, :, c], ew[:, :, a], en[:, :, a] * 0.6)
            en[:, :, c] = en[:, :, c] + en[:, :, a] * 0.6
            en[:, :, a] = en[:, :, a] - en[:, :, a] * 0.6
        np.testing.assert_allclose(num[..., t + 1], en, rtol=1e-12, err_msg=f"numbers, t = {t}")
        np.testing.assert_allclose(wgt[..., t + 1], ew, rtol=1e-12, err_msg=f"weights, t = {t}")
    assert np.abs(num[..., 19] - num[..., 0]).max() > 1.0       # (the fish did move)


def test_time_varying_natural_mortality(oracle_lib):
    """SURVEY section 8f rank 4, first slice: M from a by_year x by_step x by_age g3_param_table (R/params.R:122-129) and from a
    g3_timevariable (R/timevariable.R:1-31: 'init' until 2001 step 2, the second formula until 2002 step 1, then the third).
    The stocks of TIMEVAR_RIG only die, so num(t + 1) / num(t) = exp(-M(t) * cur_step_size) (R/action_naturalmortality.R:1-9)
    with the M in force at step t, for every length group."""
    from gadget3_b200.configs import CONFIGS
    from gadget3_b200.run_cuda import unpack_report
    from oracle.oracle_lib import Oracle
    model, p = CONFIGS["TIMEVAR_RIG"]()
    v = p.values()
    g = lambda n: float(v[p.index(n)])
    rep = unpack_report(model, Oracle(*model.pack(p)).report(v))
    dt = [3 / 12, 3 / 12, 6 / 12]
    ny = np.asarray(rep["dstart_st_year__num"])          # [length][age][area][t], 9 steps
    nv = np.asarray(rep["dstart_st_tv__num"])
    assert ny.shape == (4, 2, 1, 9) and ny.min() > 0
    for t in range(8):
        y, st = 2000 + t // 3, t % 3 + 1
        for a in (1, 2):
            want = np.exp(-g(f"st_year.Mt.{y}.{st}.{a}") * dt[st - 1])
            np.testing.assert_allclose(ny[:, a - 1, 0, t + 1] / ny[:, a - 1, 0, t], want, rtol=1e-13, err_msg=f"by_year/by_step, t = {t}")
        m = g("tvM.init") if (y, st) < (2001, 2) else g("tvM.a") if (y, st) < (2002, 1) else g("tvM.b") * 1.5
        np.testing.assert_allclose(nv[:, 0, 0, t + 1] / nv[:, 0, 0, t], np.exp(-m * dt[st - 1]), rtol=1e-13, err_msg=f"g3_timevariable, t = {t}")
    # all three formulas of the time variable are parameters of the model, whatever step they come into force at
    assert {"tvM.init", "tvM.a", "tvM.b"} <= set(p.switch)
    # g3_timevariable's own checks (R/timevariable.R:8,16)
    from gadget3_b200 import g3_timevariable
    with pytest.raises(ValueError, match="init"):
        g3_timevariable("x", {"2000": 1})
    with pytest.raises(ValueError, match="Invalid formula identifier"):
        g3_timevariable("x", {"init": 1, "20x0": 2})


def test_time_varying_selectivity(oracle_lib):
    """a by_year table inside a suitability function (R/params.R:122-129). MINI_TIMEVAR's fleet has
    l50 by year; the suitable biomass the oracle reports for a step, over the biomass at the start of that step, must be
    g3_suitability_exponentiall50 (R/suitability.R:5-13) with THAT year's l50 -- and the final suit report the last year's."""
    from gadget3_b200.configs import CONFIGS
    from gadget3_b200.run_cuda import unpack_report
    from oracle.oracle_lib import Oracle
    model, p = CONFIGS["MINI_TIMEVAR"]()
    v = p.values()
    g = lambda n: float(v[p.index(n)])
    rep = unpack_report(model, Oracle(*model.pack(p)).report(v))
    midlen = np.arange(12.5, 60, 5.0)
    alpha = g("fish.f_comm.alpha")
    suit = np.asarray(rep["detail_fish_mat_f_comm__suit"])        # [length][age][area][t]
    bio = np.asarray(rep["dstart_fish_mat__num"]) * np.asarray(rep["dstart_fish_mat__wgt"])
    for t in range(20):
        want = 1 / (1 + np.exp(-alpha * (midlen - g(f"fish.f_comm.l50.{2000 + t // 4}"))))
        got = suit[:, 1, 0, t] / bio[:, 1, 0, t]
        np.testing.assert_allclose(got, want, rtol=1e-12, err_msg=f"t = {t}")
    last = 1 / (1 + np.exp(-alpha * (midlen - g("fish.f_comm.l50.2004"))))
    np.testing.assert_allclose(np.asarray(rep["suit_fish_mat_f_comm__report"]).ravel(), last, rtol=1e-13)
    assert len({g(f"fish.f_comm.l50.{y}") for y in range(2000, 2005)}) == 5


def test_time_varying_growth(oracle_lib):
    """growth parameters that vary with time (Linf by year, K a g3_timevariable). The reference
    recomputes its growth matrices when their inputs change (R/action_grow.R:391-410). st_grow of TIMEVAR_RIG only grows,
    so its numbers after a step are the beta-binomial redistribution (growth_bbinom, R/action_grow.R:155-213, nine-lgamma
    form; delta from g3a_grow_lengthvbsimple, R/action_grow.R:44-55) of the numbers before it -- with THAT step's Linf and K --
    re-derived here with scipy's gammaln."""
    from scipy.special import gammaln
    from gadget3_b200.configs import CONFIGS
    from gadget3_b200.run_cuda import unpack_report
    from oracle.oracle_lib import Oracle
    model, p = CONFIGS["TIMEVAR_RIG"]()
    v = p.values()
    g = lambda n: float(v[p.index(n)])
    rep = unpack_report(model, Oracle(*model.pack(p)).report(v))
    num = np.asarray(rep["dstart_st_grow__num"])          # [length][age][area][t]
    L, n, dl = num.shape[0], 3, 5.0
    midlen = np.arange(12.5, 60, 5.0)
    az = lambda a: (np.maximum(a * 1000.0, 0.0) + np.log1p(np.exp(-np.abs(a * 1000.0)))) / 1000.0
    dt = [3 / 12, 3 / 12, 6 / 12]
    beta = az(g("st_grow.bbin"))
    for t in range(8):
        y, st = 2000 + t // 3, t % 3 + 1
        K = g("tvK.init") if (y, st) < (2001, 1) else g("tvK.a") if (y, st) < (2002, 3) else g("tvK.b")
        delta = az(az((g(f"st_grow.Linf_t.{y}") - midlen) * (1 - np.exp(-K * dt[st - 1]))) / dl)
        alpha = (beta * delta) / (n - delta)
        P = np.zeros((L, L))
        for x in range(n + 1):
            pm = np.exp(gammaln(n + 1) + gammaln(alpha + beta) + gammaln(n - x + beta) + gammaln(x + alpha) - gammaln(n - x + 1)
                        - gammaln(x + 1) - gammaln(n + alpha + beta) - gammaln(beta) - gammaln(alpha))
            for l in range(L):
                P[min(l + x, L - 1), l] += pm[l]
        for a in range(2):
            want = P @ num[:, a, 0, t]
            np.testing.assert_allclose(num[:, a, 0, t + 1], want, rtol=1e-9, atol=1e-12 * want.max(), err_msg=f"t = {t}, age idx {a}")
    assert len({g(f"st_grow.Linf_t.{y}") for y in (2000, 2001, 2002)}) == 3


def test_time_varying_renewal_shape(oracle_lib):
    """Renewal mean length by year and standard deviation through a g3_timevariable (SURVEY section 8f rank 4;
    R/params.R:122-129, R/timevariable.R:1-31): the host lowers one renewal record per set of years that share their shape
    (gadget3_b200/run_cuda.py), the reported renewal of every year is the normal density of THAT year's parameters on the
    length grid, normalised, times 1e4 * rec.<year> (R/action_renewal.R:56-80), worked out by hand."""
    import math
    from gadget3_b200.configs import CONFIGS
    from gadget3_b200.run_cuda import unpack_report
    from gadget3_b200.spec import K
    from oracle.oracle_lib import Oracle
    model, p = CONFIGS["RENEWAL_RIG"]()
    ints, reals = model.pack(p)
    assert ints[ints[K["G3H_STOCK_OFF"]] + K["G3S_N_RENEW"]] == 4          # four years, four shapes
    v = dict(zip(p.switch, p.values()))
    rn = unpack_report(model, Oracle(ints, reals).report(p.values()))["detail_st__renewalnum"]     # [length, age, area, time]
    mid = np.arange(12.5, 55, 5.0)
    for i, y in enumerate((2000, 2001, 2002, 2003)):
        mean, sd = v[f"st.recl.{y}"], (v["recsd.early"] if y < 2002 else v["recsd.late"])
        d = np.exp(-0.5 * ((mid - mean) / sd) ** 2) / (sd * math.sqrt(2 * math.pi))
        np.testing.assert_allclose(rn[:, 0, 0, 2 * i], d / d.sum() * 1e4 * v[f"st.rec.{y}"], rtol=1e-13)
        assert (rn[:, 1:, 0, 2 * i] == 0).all()                             # into the youngest age only
    # the same shape every year: one record, and the model equals the by-year model at equal values to the bit
    q = p.copy()
    for y in (2000, 2001, 2002, 2003):
        q.value[q.index(f"st.recl.{y}")] = 20.0
    q.value[q.index("recsd.late")] = q.value[q.index("recsd.early")]
    a = Oracle(ints, reals).eval_batch(q.values()[None, :])[0][0]
    assert np.isfinite(a)
