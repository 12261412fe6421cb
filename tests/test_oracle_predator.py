"""The oracle's stock-predator path (g3a_predate + g3a_predate_catchability_predator + g3a_predate_maxconsumption,
R/action_predate.R:190-238, 240-420) re-derived in numpy from its own REPORTed arrays -- the way the reference's tests
pin the likelihoods (tests/test-likelihood_distribution.R:419-700) -- and the two scaling properties the reference's
predator test asserts (tests/test-action_predate-predator.R:94-135): suitable biomass scales with energycontent,
consumption with predator_length^m3 when half-feeding is far above the available food. CPU only."""
import numpy as np
import pytest

from gadget3_b200.configs import CONFIGS
from gadget3_b200.run_cuda import unpack_report
from oracle.oracle_lib import Oracle


@pytest.fixture(scope="module")
def mini(oracle_lib):
    model, p = CONFIGS["MINI_ANDERSEN"]()
    return model, p, Oracle(*model.pack(p))


def _stocks(model):
    out = {}
    for a in model.actions:
        s = getattr(a, "stock", None)
        if s is not None and getattr(s, "midlen", None) is not None:
            out[s.name] = s
    return out


def _values(p, **over):
    v = p.values().copy()
    for n in p.switch:                      # no migration: every fish stays (the report's dstart_* is then the state predation sees)
        if ".migrate." in n:
            v[p.index(n)] = 0.0
    for k, x in over.items():
        v[p.index(k.replace("__", "."))] = x
    return v


def _expected_consumption(model, p, v, rep, t):
    """cons[l, age, area] by eqn 4.22-4.23: predN * M_L * psi_L * suit / (energycontent * total_predsuit), summed over predator lengths."""
    st = _stocks(model)
    g = lambda n: float(v[p.index(n)])
    dt = 3 / 12.0
    predlen = np.asarray(st["pred"].midlen, dtype=float)
    m0, m1, m2, m3 = (g(f"pred.consumption.m{i}") for i in range(4))
    temp = 0.0
    maxcons = m0 * dt * np.exp(m1 * temp - m2 * temp ** 3) * predlen ** m3                 # M_L
    hf, E = g("pred.halffeeding"), g("prey.energycontent")
    S = np.asarray(rep["suit_prey_pred__report"], dtype=float)                          # [prey length][predator length]
    B0 = (np.asarray(rep["dstart_prey__num"]) * np.asarray(rep["dstart_prey__wgt"]))[..., t]        # [l][age][area]
    predN = np.asarray(rep["dstart_pred__num"])[..., t].sum(axis=1)                     # [pl][area]
    cons = np.zeros_like(B0)
    for r in range(B0.shape[2]):
        suit = S[:, None, :] * E * B0[:, :, r][:, :, None]                              # [l][age][pl]
        ts = suit.sum(axis=(0, 1))                                                      # total_predsuit per predator length
        psi = ts / (hf * dt + ts)                                                       # feeding level
        cons[:, :, r] = (suit * (predN[:, r] * maxcons * psi / (E * ts))[None, None, :]).sum(axis=2)
    # the overconsumption block rescales every predator's share by consconv = tp' / avoid_zero(tp) with
    # tp' = B0 * dif_pmin(tp / avoid_zero(B0), 0.95, 1e3) (R/action_predate.R:381-416): the identity for cells with real
    # consumption, a squeeze towards zero for the (near-)empty ones
    return _overconsumption_squeeze(cons, B0)


def _overconsumption_squeeze(cons, B0):
    az = lambda a: (np.maximum(a * 1000.0, 0.0) + np.log1p(np.exp(-np.abs(a * 1000.0)))) / 1000.0
    ratio = cons / az(B0)
    ratio = 0.95 - (np.maximum(1e3 * (0.95 - ratio), 0.0) + np.log1p(np.exp(-np.abs(1e3 * (0.95 - ratio))))) / 1e3
    return cons * (B0 * ratio) / az(cons)


def test_consumption_rederived_from_reported_arrays(mini):
    model, p, orc = mini
    v = _values(p)
    rep = unpack_report(model, orc.report(v))
    got = np.asarray(rep["detail_prey_pred__cons"], dtype=float)
    for t in (0, 1, 2):         # before the first ageing; growth and mortality come after predation in the step
        want = _expected_consumption(model, p, v, rep, t)
        assert want.max() > 0
        assert np.abs(got[..., t] - want).max() <= 1e-9 * want.max(), t


def test_energycontent(mini):
    """tests/test-action_predate-predator.R:94-106 (suitable biomass doubles with energycontent = 2): suit / total_predsuit does
    not change, the feeding level psi = total / (half_feeding dt + total) rises, and the consumption in biomass is divided by
    energycontent (R/action_predate.R:219-237)."""
    model, p, orc = mini
    v = _values(p, prey__energycontent=2.0)
    rep = unpack_report(model, orc.report(v))
    want = _expected_consumption(model, p, v, rep, 0)
    got = np.asarray(rep["detail_prey_pred__cons"], dtype=float)[..., 0]
    assert np.abs(got - want).max() <= 1e-9 * want.max()
    base = np.asarray(unpack_report(model, orc.report(_values(p)))["detail_prey_pred__cons"], dtype=float)[..., 0]
    assert 0.5 < got.sum() / base.sum() < 1.0           # twice the energy per kg: (slightly more than) half the biomass eaten


def test_consumption_scales_with_predator_length_power_m3(mini):
    """tests/test-action_predate-predator.R:114-135: with half-feeding far above the food, consumption by predator length
    goes as length^m3. The report sums over predator lengths, so: the same run with m3 + 1 equals the re-derivation in which
    every predator length's share is multiplied by its length (first step: same state)."""
    model, p, orc = mini
    m3 = float(p.values()[p.index("pred.consumption.m3")])
    base = dict(pred__halffeeding=1e7)          # far above the suitable biomass (~1e4), consumption still well above avoid_zero's scale
    va, vb = _values(p, **base), _values(p, pred__consumption__m3=m3 + 1.0, **base)
    ra, rb = unpack_report(model, orc.report(va)), unpack_report(model, orc.report(vb))
    want_b = _expected_consumption(model, p, vb, ra, 0)         # state of run a at t = 0 == state of run b at t = 0
    got_b = np.asarray(rb["detail_prey_pred__cons"], dtype=float)[..., 0]
    assert np.abs(got_b - want_b).max() <= 1e-9 * want_b.max()
    got_a = np.asarray(ra["detail_prey_pred__cons"], dtype=float)[..., 0]
    st = _stocks(model)
    lo, hi = min(st["pred"].midlen), max(st["pred"].midlen)
    ratio = got_b.sum() / got_a.sum()                           # a length-weighted mean of the predator lengths
    assert lo < ratio < hi


# ---------------------------------------------------------------- g3a_otherfood (R/action_renewal.R:276-308)
@pytest.fixture(scope="module")
def mini_of(oracle_lib):
    model, p = CONFIGS["MINI_OTHERFOOD"]()
    return model, p, Oracle(*model.pack(p))


def test_otherfood_is_assigned_every_step(mini_of):
    """tests/test-action_predate-predator.R:39-45 sets up g3a_otherfood(num_f = of_abund by year x of_abund.step by step,
    wgt_f = of_meanwgt): the state at the start of EVERY step is the formula's value in every cell, whatever was eaten in
    the step before (the stock has no dynamics of its own)."""
    model, p, orc = mini_of
    v = _values(p)
    years = [2000, 2001, 2002, 2003]
    for i, y in enumerate(years):
        v[p.index(f"otherfood.of_abund.{y}")] = 4e4 + 1e4 * i
    for s in range(1, 5):
        v[p.index(f"otherfood.of_abund.step.{s}")] = 0.5 + 0.25 * s
    v[p.index("otherfood.of_meanwgt")] = 0.07
    rep = unpack_report(model, orc.report(v))
    num, wgt = np.asarray(rep["dstart_otherfood__num"]), np.asarray(rep["dstart_otherfood__wgt"])
    assert num.shape[-1] == 16
    for t in range(16):
        want = (4e4 + 1e4 * (t // 4)) * (0.5 + 0.25 * (t % 4 + 1))
        assert (num[..., t] == want).all(), t
        assert (wgt[..., t] == 0.07).all()
    assert np.asarray(rep["detail_otherfood_pred__cons"]).min() > 0          # and it IS eaten, every step, in both areas


def _expected_two_prey_consumption(model, p, v, rep, t, pref=None):
    """As _expected_consumption, with the otherfood stock in the predator's total suitable biomass (R/action_predate.R:314-333:
    total_predsuit sums over every prey of the predator) -- returns the consumption of `prey` and of `otherfood`."""
    st = _stocks(model)
    g = lambda n: float(v[p.index(n)])
    dt = 3 / 12.0
    predlen = np.asarray(st["pred"].midlen, dtype=float)
    m0, m1, m2, m3 = (g(f"pred.consumption.m{i}") for i in range(4))
    maxcons = m0 * dt * predlen ** m3
    hf = g("pred.halffeeding")
    E = {"prey": g("prey.energycontent"), "otherfood": g("otherfood.energycontent")}
    S = {k: np.asarray(rep[f"suit_{k}_pred__report"], dtype=float) for k in E}
    B0 = {k: (np.asarray(rep[f"dstart_{k}__num"]) * np.asarray(rep[f"dstart_{k}__wgt"]))[..., t] for k in E}
    predN = np.asarray(rep["dstart_pred__num"])[..., t].sum(axis=1)
    cons = {k: np.zeros_like(B0[k]) for k in E}
    for r in range(predN.shape[1]):
        suit = {k: S[k][:, None, :] * E[k] * B0[k][:, :, r][:, :, None] for k in E}
        if pref is not None:            # (suit_f * energycontent * num * wgt)^prey_preference, R/action_predate.R:220-225
            suit = {k: suit[k] ** pref[k] for k in E}
        ts = sum(suit[k].sum(axis=(0, 1)) for k in E)
        psi = ts / (hf * dt + ts)
        for k in E:
            cons[k][:, :, r] = (suit[k] * (predN[:, r] * maxcons * psi / (E[k] * ts))[None, None, :]).sum(axis=2)
    return {k: _overconsumption_squeeze(cons[k], B0[k]) for k in E}


def test_otherfood_shares_the_predators_appetite(mini_of):
    model, p, orc = mini_of
    v = _values(p)
    rep = unpack_report(model, orc.report(v))
    for t in (0, 1, 2):
        want = _expected_two_prey_consumption(model, p, v, rep, t)
        for k in ("prey", "otherfood"):
            got = np.asarray(rep[f"detail_{k}_pred__cons"], dtype=float)[..., t]
            assert want[k].max() > 0
            assert np.abs(got - want[k]).max() <= 1e-9 * want[k].max(), (k, t)
    # more other food => less of the prey eaten (same predators, same appetite)
    v2 = v.copy()
    for y in (2000, 2001, 2002, 2003):
        v2[p.index(f"otherfood.of_abund.{y}")] *= 10
    rep2 = unpack_report(model, orc.report(v2))
    assert np.asarray(rep2["detail_prey_pred__cons"])[..., 0].sum() < np.asarray(rep["detail_prey_pred__cons"])[..., 0].sum()


def test_prey_preferences(oracle_lib):
    """R/action_predate.R:208,220-225: the suitable biomass of each prey is raised to its prey_preference before it enters the
    predator's total and the split of the consumption. MINI_PREFS: 0.9 (a parameter) for the prey, the constant 1.3 for the
    other food; consumption re-derived from the reported arrays, and preference 1 for both gives MINI_OTHERFOOD back."""
    model, p = CONFIGS["MINI_PREFS"]()
    orc = Oracle(*model.pack(p))
    assert "pref.prey" in p.switch
    v = _values(p)
    rep = unpack_report(model, orc.report(v))
    for t in (0, 1, 2):
        want = _expected_two_prey_consumption(model, p, v, rep, t, pref={"prey": 0.9, "otherfood": 1.3})
        for k in ("prey", "otherfood"):
            got = np.asarray(rep[f"detail_{k}_pred__cons"], dtype=float)[..., t]
            assert np.abs(got - want[k]).max() <= 1e-9 * want[k].max(), (k, t)
    plain = _expected_two_prey_consumption(model, p, v, rep, 0)
    want = _expected_two_prey_consumption(model, p, v, rep, 0, pref={"prey": 0.9, "otherfood": 1.3})
    assert np.abs(plain["prey"] - want["prey"]).max() > 1e-3 * plain["prey"].max()          # the exponents do something


def test_stomach_content_distribution(oracle_lib):
    """g3l_catchdistribution(predators = list(pred), stocks = list(prey)) (R/likelihood_distribution.R:200,275-292: stock
    predators are collected exactly as fleets are, predprey__cons summed over the predator's own dimensions): the model
    array in biomass is the reported consumption detail_prey_pred__cons of that step, summed into the data's length bins;
    data with predator_length columns are refused."""
    from gadget3_b200 import G3Unsupported, g3l_catchdistribution, g3l_distribution_sumofsquares
    from gadget3_b200.configs import CONFIGS
    from gadget3_b200.run_cuda import unpack_report
    from oracle.oracle_lib import Oracle
    model, p = CONFIGS["MINI_STOMACH"]()
    rep = unpack_report(model, Oracle(*model.pack(p)).report(p.values()))
    cons = rep["detail_prey_pred__cons"]                   # [length (5..50 by 5), age, area, time]
    got = rep["cdist_sumofsquares_stomach_pred_model__wgt"]        # [time, areagroup, stockgroup, age, bin]
    assert got.shape[0] == 4 and got.shape[-1] == 3
    bins = [slice(0, 3), slice(3, 6), slice(6, 9)]          # [5,20) [20,35) [35,Inf)
    for i in range(4):
        t = 4 * i + 1                                       # step 2 of every year
        hand = np.array([cons[b, :, :, t].sum() for b in bins])
        assert hand.min() > 0
        np.testing.assert_allclose(got[i].ravel(), hand, rtol=1e-12)
    pred, prey = model.actions[2].stock, model.actions[1].stock
    with pytest.raises(G3Unsupported, match="predator_length"):
        g3l_catchdistribution("x", dict(year=[2000], step=[2], predator_length=["[20,30)"], weight=[1.0]), predators=[pred],
                              stocks=[prey], function_f=g3l_distribution_sumofsquares())
