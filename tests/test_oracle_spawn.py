"""g3a_spawn in the oracle (R/action_spawn.R:1-233) pinned the way the reference's own test does it
(tests/test-action_spawn.R:41-190): a rig in which, for every recruitment function, a mature stock with assigned number and
mean weight spawns into a stock nothing else touches; the total that arrives each step must equal the recruitment formula
evaluated by hand on the parents. Then the pieces the rig does not exercise -- proportion spawning by length, spawning
mortality, run_step, the length distribution and mean weight of the offspring -- on the closed life cycle of MINI_SPAWN.
CPU only."""
import numpy as np
import pytest

from gadget3_b200.configs import CONFIGS, SPAWN_KINDS
from gadget3_b200.run_cuda import unpack_report
from oracle.oracle_lib import Oracle

MIDLEN = np.array([15.0, 25.0, 35.0, 45.0])


@pytest.fixture(scope="module")
def rig(oracle_lib):
    model, p = CONFIGS["SPAWN_RIG"]()
    rep = unpack_report(model, Oracle(*model.pack(p)).report(p.values()))
    return model, p, rep


def _par(p, name):
    return float(p.values()[p.index(name)])


def test_rig_produces_the_assigned_parents(rig):
    """tests/test-action_spawn.R:100-109: 'Test rig produced increasing abundance / mean weight' -- here abundance by year x step."""
    model, p, rep = rig
    for k in SPAWN_KINDS:
        num = np.asarray(rep[f"dstart_mat_{k}__num"])          # [length][age][area][t]
        for t in range(16):
            want = _par(p, f"of_abund.{1982 + t // 4}") * _par(p, f"of_abund.step.{t % 4 + 1}")
            assert (num[..., t] == want).all()
        assert (np.asarray(rep[f"dstart_mat_{k}__wgt"]) == _par(p, "of_meanwgt")).all()


def _arrivals(rep, k):
    n = np.asarray(rep[f"dstart_imm_{k}__num"]).sum(axis=(0, 1, 2))      # total by t
    return np.diff(n)           # what arrived in steps 0 .. 14 (the run has 16 steps; the last one is not observed)


def _parents(rep, k):
    num = np.asarray(rep[f"dstart_mat_{k}__num"])[:, 0, 0, :15]
    wgt = np.asarray(rep[f"dstart_mat_{k}__wgt"])[:, 0, 0, :15]
    return num, wgt


def test_recruitment_functions_match_their_formulas(rig):
    """tests/test-action_spawn.R:111-190, one assertion per recruitment function ('Matches formula', 'Total biomass * mu', ...)."""
    model, p, rep = rig
    age = 5.0
    for k in SPAWN_KINDS:
        got = _arrivals(rep, k)
        num, wgt = _parents(rep, k)
        s = (num * wgt).sum(axis=0)
        if k == "fecundity":
            p0, p1, p2, p3, p4 = (_par(p, f"fecundity_p{i}") for i in range(5))
            want = (MIDLEN[:, None] ** p1 * age ** p2 * num ** p3 * wgt ** p4).sum(axis=0) * p0
        elif k == "simplessb":
            want = s * _par(p, "simplessb_mu")
        elif k == "ricker":
            want = _par(p, "ricker_mu") * s * np.exp(-_par(p, "ricker_lambda") * s)
        elif k == "bevertonholt":
            want = _par(p, "bevertonholt_mu") * s / (_par(p, "bevertonholt_lambda") + s)
        elif k == "bevertonholt_ss3":
            h, B0, R0 = (_par(p, f"mat_{k}.spawn.{x}") for x in ("h", "B0", "R0"))
            R = np.array([np.exp(_par(p, f"mat_{k}.spawn.R_exp.{1982 + t // 4}")) * R0 for t in range(15)])
            want = 4 * h * s * R / (B0 * (1 - h) + s * (5 * h - 1))
        else:
            r0, blim = _par(p, "hockeystick_r0"), _par(p, "hockeystick_blim")
            want = r0 * np.minimum(s / blim, 1.0)
            assert (s < blim).any() and (s > blim).any()            # both arms of the stick
            # the differentiable form, R/action_spawn.R:81: r0 * logspace_add(-1000 s / blim, -1000) / -1000
            want = r0 * np.logaddexp(-1000 * s / blim, -1000.0) / -1000
        assert (got > 0).all(), k
        np.testing.assert_allclose(got, want, rtol=1e-11, err_msg=k)


def test_offspring_length_distribution_and_weight(rig):
    """R/action_spawn.R:197-209: the arrivals are spread as exp(-((midlen - mean) / sd)^2 / 2), normalised over the output
    stock, and carry the mean weight alpha * midlen^beta (the first arrival meets an empty stock: ratio_add_pop gives it)."""
    model, p, rep = rig
    k = "simplessb"
    n1 = np.asarray(rep[f"dstart_imm_{k}__num"])[:, 0, 0, 1]
    w1 = np.asarray(rep[f"dstart_imm_{k}__wgt"])[:, 0, 0, 1]
    linf, kk, t0, sd = (_par(p, f"imm_{k}.{x}") for x in ("Linf", "K", "t0", "rec.sd"))
    mean = linf * (1 - np.exp(-kk * (0 - t0)))          # g3a_renewal_vonb_t0 at the output stock's youngest age, 0
    shape = np.exp(-(((MIDLEN - mean) * (1 / sd)) ** 2) * 0.5)
    np.testing.assert_allclose(n1 / n1.sum(), shape / shape.sum(), rtol=1e-12)
    np.testing.assert_allclose(w1, _par(p, f"imm_{k}.walpha") * MIDLEN ** _par(p, f"imm_{k}.wbeta"), rtol=1e-9)


def test_proportion_mortality_and_run_step(oracle_lib):
    """MINI_SPAWN: spawning only on step 2; the parents lose spawningnum * (1 - exp(-mortality)) with spawningnum = num *
    proportion(length); with the recruitment switched off (mu -> 0) the immature stock's youngest age stays empty after
    the first ageing; the arrivals scale with mu at small lambda."""
    model, p = CONFIGS["MINI_SPAWN"]()
    orc = Oracle(*model.pack(p))
    v = p.values().copy()
    rep = unpack_report(model, orc.report(v))
    young = np.asarray(rep["dstart_fish_imm__num"])[:, 0, :, :].sum(axis=(0, 1))          # age 1 by t
    # year 2 (t = 4 .. 7): empty after the ageing, filled by the spawning of step 2 (t = 5), seen from t = 6
    assert young[4] == 0 and young[5] == 0 and young[6] > 1e4 and young[7] < young[6]
    v0 = v.copy(); v0[p.index("fish_mat.spawn.mu")] = 0.0
    rep0 = unpack_report(model, orc.report(v0))
    young0 = np.asarray(rep0["dstart_fish_imm__num"])[:, 0, :, :].sum(axis=(0, 1))
    assert (young0[4:8] == 0).all()
    # spawning mortality: compare the parents with and without it, one step after the first spawning. Without fishing and
    # with mortality_f = m the surviving parents are num * (1 - prop (1 - exp(-m))) before the matured fish arrive; here we
    # check the sign and the size of the total effect against the length-resolved bound
    vm = v.copy(); vm[p.index("fish_mat.spawn.mort")] = 0.0
    repm = unpack_report(model, orc.report(vm))
    a = np.asarray(rep["dstart_fish_mat__num"])[..., 2].sum()
    b = np.asarray(repm["dstart_fish_mat__num"])[..., 2].sum()
    m = float(v[p.index("fish_mat.spawn.mort")])
    assert b * (1 - (1 - np.exp(-m))) < a < b           # some, but not all, parents spawn (proportion by length < 1)
    # nothing happens to the parents on the other steps
    np.testing.assert_array_equal(np.asarray(rep["dstart_fish_mat__num"])[..., 1], np.asarray(repm["dstart_fish_mat__num"])[..., 1])
    # arrivals proportional to mu (lambda small: exp(-lambda s) ~ 1 to a few 1e-1; compare exactly through the formula)
    v2 = v.copy(); v2[p.index("fish_mat.spawn.mu")] *= 2; v2[p.index("fish_mat.spawn.lambda")] = 0.0
    v1 = v.copy(); v1[p.index("fish_mat.spawn.lambda")] = 0.0
    y1 = np.asarray(unpack_report(model, orc.report(v1))["dstart_fish_imm__num"])[:, 0, :, 2]
    y2 = np.asarray(unpack_report(model, orc.report(v2))["dstart_fish_imm__num"])[:, 0, :, 2]
    y0 = np.asarray(rep0["dstart_fish_imm__num"])[:, 0, :, 2]
    np.testing.assert_allclose(y2 - y0, 2 * (y1 - y0), rtol=1e-9)


def test_spawn_api_errors():
    """tests/test-action_spawn.R:31-36: 'Length of output_ratios must match', 'output_ratios must sum to 1'."""
    from gadget3_b200 import g3_stock, g3a_spawn, g3a_spawn_recruitment_simplessb, g3s_age
    mat = g3s_age(g3_stock("stock_mat", [30, 40]), 5, 7)
    i1, i2 = g3s_age(g3_stock("stock_imm1", range(10, 50, 10)), 3, 7), g3s_age(g3_stock("stock_imm2", range(10, 50, 10)), 4, 7)
    rf = g3a_spawn_recruitment_simplessb()
    with pytest.raises(ValueError, match="output_ratios"):
        g3a_spawn(mat, rf, output_stocks=[i1, i2], output_ratios=[9, 9, 9])
    with pytest.raises(ValueError, match="output_ratios"):
        g3a_spawn(mat, rf, output_stocks=[i1, i2], output_ratios=[9, 9])


def test_spawning_weight_loss_by_hand(oracle_lib):
    """g3a_spawn(weightloss_f) / weightloss_args (R/action_spawn.R:166-181 -> g3a_weightloss with apply_to_pop =
    stock__spawningnum, R/action_weightloss.R:19-44): on parents nothing else touches, number and mean weight at the start
    of step 3 follow from those at the start of step 2 by the spawning formulas alone --
    spawningnum = num * proportion(l); num' = num - spawningnum (1 - exp(-mortality));
    wgt' = ratio_add_pop(wgt, num' - spawningnum, (wgt - min_weight) (1 - rel_loss) + min_weight, spawningnum)."""
    import math
    from gadget3_b200 import G3Unsupported, g3_parameterized, g3a_spawn, g3a_spawn_recruitment_simplessb
    from gadget3_b200.configs import CONFIGS
    from gadget3_b200.run_cuda import unpack_report
    from oracle.oracle_lib import Oracle
    model, p = CONFIGS["WLOSS_RIG"]()
    rep = unpack_report(model, Oracle(*model.pack(p)).report(p.values()))
    v = dict(zip(p.switch, p.values()))
    mid = np.arange(15, 60, 10.0)
    az = lambda x: np.logaddexp(1000 * x, 0) / 1000
    for k in ("rel", "min"):
        n, w = rep[f"dstart_mat_{k}__num"], rep[f"dstart_mat_{k}__wgt"]        # [length, age, area, time]
        for t in (1, 5):                                                      # step 2 of both years
            prop = (1 / (1 + np.exp(-v[f"mat_{k}.spawn.alpha"] * (mid - v[f"mat_{k}.spawn.l50"]))))[:, None, None]
            spn = n[..., t] * prop
            num2 = n[..., t] - spn * (1 - math.exp(-v[f"mat_{k}.spawn.mort"]))
            mw = v.get(f"mat_{k}.spawn.minwgt", 0.0)
            lost = (w[..., t] - mw) * (1 - v[f"mat_{k}.spawn.wloss"]) + mw
            np.testing.assert_allclose(n[..., t + 1], num2, rtol=1e-14)
            np.testing.assert_allclose(w[..., t + 1], (w[..., t] * (num2 - spn) + lost * spn) / az(num2), rtol=1e-14)
            assert (w[..., t + 1] < w[..., t]).all()
        assert np.array_equal(w[..., 1], w[..., 0]) and np.array_equal(w[..., 3], w[..., 2])      # no other step touches them
    with pytest.raises(G3Unsupported, match="abs_loss"):
        g3a_spawn(model.actions[1].stock, g3a_spawn_recruitment_simplessb(g3_parameterized("mu")), weightloss_args=dict(abs_loss=0.1))
