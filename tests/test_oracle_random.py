"""g3l_random_dnorm / g3l_random_walk (R/likelihood_random.R:14-109) in the oracle, pinned the way the reference's own test
pins them (tests/test-likelihood_random.R:36-150): the reported terms and the objective against the normal density worked
out by hand from the parameter values."""
import math

import numpy as np
import pytest

from gadget3_b200 import G3Unsupported, g3_parameterized, g3l_random_dnorm, g3l_random_walk
from gadget3_b200.configs import CONFIGS, parameter_batch
from gadget3_b200.run_cuda import unpack_report
from oracle.oracle_lib import Oracle


def _ldnorm(x, mu, sd):
    return -0.5 * math.log(2 * math.pi) - math.log(sd) - 0.5 * ((x - mu) / sd) ** 2


def _walk(vals, sd):
    """sum_walk_dnorm of tests/test-likelihood_random.R:119-128."""
    return sum(-_ldnorm(vals[i + 1], vals[i], sd) for i in range(len(vals) - 1))


def _hand(v):
    years, steps = range(1990, 1995), range(1, 5)
    return dict(
        dnorm_log=-_ldnorm(v["dnorm_log"], v["dnorm_log_mean"], v["dnorm_log_sigma"]),
        dnorm_lin=math.exp(_ldnorm(v["dnorm_lin"], v["dnorm_lin_mean"], v["dnorm_lin_sigma"])),
        walk_year=_walk([v[f"walk_year.{y}"] for y in years], v["walk_year_sigma"]),
        walk_step=_walk([v[f"walk_step.{y}.{s}"] for y in years for s in steps], v["walk_step_sigma"]))


def test_rig_against_the_normal_density_by_hand(oracle_lib):
    model, p = CONFIGS["RANDOM_RIG"]()
    assert model.nll_names[:4] == ["nll_random_dnorm_dnorm_lin", "nll_random_dnorm_dnorm_log",
                                   "nll_random_walk_walk_step", "nll_random_walk_walk_year"]
    orc = Oracle(*model.pack(p))
    opt = p.optimised_indices()
    P = parameter_batch(p, 5, seed=3)
    full = np.tile(p.values(), (5, 1))
    full[:, opt] = P
    nll, comp, _ = orc.eval_batch(full, want_comp=True)
    names = list(p.switch)
    for b in range(5):
        h = _hand(dict(zip(names, full[b])))
        got = dict(dnorm_lin=comp[b, 0], dnorm_log=comp[b, 1], walk_step=comp[b, 2], walk_year=comp[b, 3])
        for k in h:
            assert got[k] == pytest.approx(h[k], rel=1e-12), k
        assert comp[b, -2] == 0.0                                # no understocking term
        assert nll[b] == pytest.approx(sum(h.values()), rel=1e-12)   # "nll matches (and has been included only once)"
    # the reported per-step rows: single-parameter terms score once, at the last step; the yearly walk at the last
    # step of 1991 .. 1994; the step walk on every step but the first
    rep = unpack_report(model, orc.report(full[0]))
    assert np.flatnonzero(rep["nll_random_dnorm_dnorm_log__dnorm"]).tolist() == [19]
    assert np.flatnonzero(rep["nll_random_dnorm_dnorm_lin__dnorm"]).tolist() == [19]
    assert np.flatnonzero(rep["nll_random_walk_walk_year__dnorm"]).tolist() == [7, 11, 15, 19]
    assert np.flatnonzero(rep["nll_random_walk_walk_step__dnorm"]).tolist() == list(range(1, 20))
    v = dict(zip(names, full[0]))
    assert rep["nll_random_walk_walk_year__dnorm"][7] == pytest.approx(-_ldnorm(v["walk_year.1991"], v["walk_year.1990"], v["walk_year_sigma"]), rel=1e-12)


def test_weights_switch_the_terms_off(oracle_lib):
    model, p = CONFIGS["RANDOM_RIG"]()
    v = p.values().copy()
    names = list(p.switch)
    for n in ("dnorm_log_enabled", "dnorm_lin_enabled", "walk_year_enabled", "walk_step_enabled"):
        v[names.index(n)] = 0.0
    nll, comp, _ = Oracle(*model.pack(p)).eval_batch(v[None, :], want_comp=True)
    assert nll[0] == 0.0 and (comp[0, :4] != 0.0).all()         # the terms are still reported, the objective ignores them


def test_vignette_model_with_recruitment_deviations(oracle_lib):
    """C1 with a dnorm and a random walk on fish.rec.<year> * 1e4: the objective grows by exactly those terms."""
    m0, p0 = CONFIGS["C1"]()
    m1, p1 = CONFIGS["C1_RANDOM"]()
    v1 = p1.values()
    d1 = dict(zip(p1.switch, v1))
    v0 = np.array([d1[n] for n in p0.switch])
    n0 = Oracle(*m0.pack(p0)).eval_batch(v0[None, :])[0][0]
    nll, comp, _ = Oracle(*m1.pack(p1)).eval_batch(v1[None, :], want_comp=True)
    rec = [d1[f"fish.rec.{y}"] * 1e4 for y in range(1990, 2024)]
    dn = sum(-_ldnorm(r, d1["rec_mean"], d1["rec_sigma"]) for r in rec)
    wk = _walk(rec, d1["rec_sigma"])
    i = m1.nll_names.index("nll_random_dnorm_rec_dnorm")
    assert comp[0, i] == pytest.approx(dn, rel=1e-12) and comp[0, i + 1] == pytest.approx(wk, rel=1e-12)
    # (the two new bounded parameters add their own, tiny, bounds penalty)
    assert nll[0] - n0 == pytest.approx(dn + wk + (comp[0, -1] - Oracle(*m0.pack(p0)).eval_batch(v0[None, :], want_comp=True)[1][0, -1]), rel=1e-9)


def test_gradient_of_the_terms(oracle_lib):
    """d/dx of -log dnorm(x, mu, sd) = (x - mu) / sd^2: the oracle's duals on the rig."""
    model, p = CONFIGS["RANDOM_RIG"]()
    names = list(p.switch)
    v = p.values()
    d = dict(zip(names, v))
    tix = np.array([names.index("dnorm_log"), names.index("walk_year.1990"), names.index("walk_year.1992")], dtype=np.int32)
    _, _, g = Oracle(*model.pack(p)).eval_batch(v[None, :], tangent_idx=tix)
    s2 = d["walk_year_sigma"] ** 2
    assert g[0, 0] == pytest.approx((d["dnorm_log"] - d["dnorm_log_mean"]) / d["dnorm_log_sigma"] ** 2, rel=1e-12)
    assert g[0, 1] == pytest.approx(-(d["walk_year.1991"] - d["walk_year.1990"]) / s2, rel=1e-12)
    assert g[0, 2] == pytest.approx((d["walk_year.1992"] - d["walk_year.1991"]) / s2 - (d["walk_year.1993"] - d["walk_year.1992"]) / s2, rel=1e-12)


def test_period_guessing_and_argument_checks():
    gp = g3_parameterized
    assert g3l_random_dnorm("a", gp("x")).period == "single"
    assert g3l_random_dnorm("a", gp("x", by_year=True)).period == "year"
    assert g3l_random_walk("a", gp("x", by_year=True, by_step=True, exponentiate=True) * 2.0).period == "step"
    with pytest.raises(ValueError, match="Seasonal random parameters not supported yet"):
        g3l_random_dnorm("a", gp("x", by_step=True))
    with pytest.raises(ValueError, match="Per-stock random parameters not supported yet"):
        g3l_random_dnorm("a", gp("x", by_stock=True, by_year=True))
    with pytest.raises(ValueError):
        g3l_random_dnorm("a", gp("x"), period="decade")
    with pytest.raises(ValueError):
        g3l_random_dnorm("a", gp("x"), log_f=1)
    assert g3l_random_walk("w", gp("x", by_year=True)).report_name == "nll_random_walk_w"


def test_single_terms_follow_retro_years(oracle_lib):
    """period = 'single' scores at cur_time == total_steps (R/likelihood_random.R:46): with retro_years = 2 the run ends
    two years earlier and the term is still added, once, at ITS last step; the yearly walk loses its last two steps."""
    model, p = CONFIGS["RANDOM_RIG"]()
    orc = Oracle(*model.pack(p))
    v = p.values().copy()
    v[p.index("retro_years")] = 2.0
    nll, comp, _ = orc.eval_batch(v[None, :], want_comp=True)
    d = dict(zip(p.switch, v))
    h = _hand(d)
    assert comp[0, 1] == pytest.approx(h["dnorm_log"], rel=1e-12) and comp[0, 0] == pytest.approx(h["dnorm_lin"], rel=1e-12)
    wy = [d[f"walk_year.{y}"] for y in range(1990, 1993)]
    assert comp[0, 3] == pytest.approx(_walk(wy, d["walk_year_sigma"]), rel=1e-12)
    ws = [d[f"walk_step.{y}.{s}"] for y in range(1990, 1993) for s in range(1, 5)]
    assert comp[0, 2] == pytest.approx(_walk(ws, d["walk_step_sigma"]), rel=1e-12)
    rep = unpack_report(model, orc.report(v))
    assert np.flatnonzero(rep["nll_random_dnorm_dnorm_log__dnorm"]).tolist() == [11]
