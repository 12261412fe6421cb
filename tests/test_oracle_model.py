"""Whole-model behaviour of the CPU oracle on the C1 / C2 configurations (CPU only).

These are internal-consistency properties (the oracle's arithmetic itself is pinned by
test_oracle_kat.py and, where built, by the reference's generated objective in oracle/_ref).
"""
import numpy as np
import pytest

from gadget3_b200.configs import CONFIGS, parameter_batch
from gadget3_b200.run_cuda import unpack_report
from gadget3_b200.spec import K
from oracle.oracle_lib import Oracle


@pytest.fixture(scope="module")
def setups(oracle_lib):
    out = {}
    for n in ("C1", "C2"):
        model, p = CONFIGS[n]()
        ints, reals = model.pack(p)
        out[n] = (model, p, Oracle(ints, reals))
    return out


def _full(p, P):
    full = np.tile(p.values(), (P.shape[0], 1))
    full[:, p.optimised_indices()] = P
    return full


@pytest.mark.parametrize("name", ["C1", "C2"])
def test_nll_is_weighted_sum_of_components(setups, name):
    model, p, orc = setups[name]
    full = _full(p, parameter_batch(p, 8))
    nll, comp, _ = orc.eval_batch(full, want_comp=True)
    assert np.isfinite(nll).all() and (nll > 0).all()
    w = np.array([p.value[p.index(n + "_weight")] for n in model.comp_names] + [1e8, 1.0])
    np.testing.assert_allclose((comp * w).sum(axis=1), nll, rtol=1e-12)


@pytest.mark.parametrize("name", ["C1", "C2"])
def test_report_rows_sum_to_components(setups, name):
    model, p, orc = setups[name]
    full = _full(p, parameter_batch(p, 2, seed=5))
    nll, comp, _ = orc.eval_batch(full, want_comp=True)
    rep = unpack_report(model, orc.report(full[1]))
    assert rep["nll"] == pytest.approx(nll[1], rel=1e-14)
    for i, n in enumerate(model.comp_names):
        assert rep[f"nll_{n}__{model.comp_units[i]}"].sum() == pytest.approx(comp[1, i], rel=1e-12)
    assert rep["nll_understocking__wgt"].sum() == pytest.approx(comp[1, -2], rel=1e-12, abs=1e-300)
    # dstart arrays come back in the model's own dimension order, time last
    for sname, (L, A, R), dims in zip(model.stock_names, model.stock_shapes, model.stock_dims):
        arr = rep[f"dstart_{sname}__num"]
        want = (L, A, R, model.n_steps) if dims[1] == "age" else (L, R, A, model.n_steps)
        assert arr.shape == want
        assert np.isfinite(arr).all() and (arr >= 0).all()


def test_initial_conditions_block(setups):
    """dstart at cur_time == 0 is the g3a_initialconditions_normalcv formula (R/action_renewal.R:139-163)."""
    model, p, orc = setups["C1"]
    v = p.values()
    rep = unpack_report(model, orc.report(v))
    num0 = rep["dstart_fish__num"][:, 0, :, 0]          # length, area, age, time
    wgt0 = rep["dstart_fish__wgt"][:, 0, :, 0]
    g = lambda n: v[p.index(n)]
    midlen = np.arange(15, 106, 10, dtype=float)
    az = lambda x: (np.maximum(1000 * x, 0) + np.log1p(np.exp(-np.abs(1000 * x)))) / 1000
    for ai, age in enumerate(range(1, 6)):
        mu = g("fish.Linf") * (1 - np.exp(-g("fish.K") * ((age - 0.25) - g("fish.t0"))))
        sd = az(mu * g("fish.lencv"))
        d = np.exp(-0.5 * ((midlen - mu) / sd) ** 2) / (sd * np.sqrt(2 * np.pi))
        factor = g("fish.init.scalar") * g(f"fish.init.{age}") * np.exp(-(g(f"fish.M.{age}") + g("init.F")) * (age - g("recage")))
        np.testing.assert_allclose(num0[:, ai], d / az(d.sum()) * 10000 * factor, rtol=1e-12)
        np.testing.assert_allclose(wgt0[:, ai], g("fish.walpha") * midlen ** g("fish.wbeta"), rtol=1e-13)


def test_negative_retro_years_is_nan(setups):
    """R/action_time.R:78-79: warn and return NaN."""
    model, p, orc = setups["C1"]
    v = p.values().copy()
    v[p.index("retro_years")] = -1
    nll, _, _ = orc.eval_batch(v[None, :])
    assert np.isnan(nll[0])


def test_retro_years_shortens_the_run(setups):
    model, p, orc = setups["C1"]
    v = p.values().copy()
    a = orc.eval_batch(v[None, :])[0][0]
    v[p.index("retro_years")] = 3
    b = orc.eval_batch(v[None, :])[0][0]
    assert np.isfinite(b) and b != a


@pytest.mark.parametrize("name", ["C1", "C2"])
def test_dual_gradient_matches_central_differences(setups, name):
    """The reference has no gradient test; the Dual<K> oracle is checked against
    central differences of its own value."""
    model, p, orc = setups[name]
    x0 = _full(p, parameter_batch(p, 1, seed=11))[0]
    opt = p.optimised_indices()
    rng = np.random.default_rng(3)
    lo, up = np.asarray(p.lower), np.asarray(p.upper)
    # the bounds penalty (scale 1e6) is too stiff for finite differences right at a bound
    inside = [i for i in opt if min(x0[i] - lo[i], up[i] - x0[i]) > 1e-3 * (up[i] - lo[i])]
    idx = np.sort(rng.choice(inside, size=10, replace=False)).astype(np.int32)
    _, _, g = orc.eval_batch(x0[None, :], tangent_idx=idx)
    def cdiff(i, h):
        xp, xm = x0.copy(), x0.copy()
        xp[i] += h
        xm[i] -= h
        return (orc.eval_batch(xp[None, :])[0][0] - orc.eval_batch(xm[None, :])[0][0]) / (2 * h)

    for k, i in enumerate(idx):
        # The 1e8-weighted understocking term (a squared difference of near-equal sums) puts noise
        # of about 1e-9 / h on a difference quotient, so use a large step plus Richardson extrapolation.
        h = 2e-3 * max(abs(x0[i]), 1e-3 * (up[i] - lo[i]))
        fd = (4.0 * cdiff(i, h / 2) - cdiff(i, h)) / 3.0
        assert g[0, k] == pytest.approx(fd, rel=2e-4, abs=1e-5 * abs(g[0]).max()), p.switch[i]


def test_threads_do_not_change_results(setups):
    model, p, orc = setups["C2"]
    full = _full(p, parameter_batch(p, 13, seed=9))
    a = orc.eval_batch(full, nthreads=1)[0]
    b = orc.eval_batch(full, nthreads=4)[0]
    assert (a == b).all()


def test_bounds_penalty_zero_inside_huge_outside(setups):
    """tests/test-likelihood_bounds.R:20-41: inside its bounds a parameter adds exactly nothing; one unit outside a
    60-wide range pushes nll beyond 1e8; a new bound that excludes the current value does the same."""
    model, p, orc = setups["C1"]
    full = p.values()[None, :].copy()
    nll0, comp0, _ = orc.eval_batch(full, want_comp=True)
    assert comp0[0, -1] == 0.0                                   # "nll: start off inside bounds"
    i = p.index("fish.K")                                        # bounds 0.04 .. 1.2
    out = full.copy(); out[0, i] = p.upper[i] + (p.upper[i] - p.lower[i]) / 60.0
    nll1, comp1, _ = orc.eval_batch(out, want_comp=True)
    assert nll1[0] - nll0[0] > 1e8 and comp1[0, -1] > 1e8        # "outside initial upper bound"
    q = p.copy(); q.upper[i] = p.value[i] - 0.01                 # "outside new upper bound"
    nll2, comp2, _ = Oracle(*model.pack(q)).eval_batch(full, want_comp=True)
    assert comp2[0, -1] > 1e8 and nll2[0] > 1e8
    q = p.copy(); q.lower[i] = p.value[i] + 0.01                 # "outside new lower bound"
    assert Oracle(*model.pack(q)).eval_batch(full, want_comp=True)[1][0, -1] > 1e8
