"""The oracle's g3l_distribution_multinomial and g3l_distribution_surveyindices_linear, re-derived from its own
REPORTed model arrays the way the reference's tests do it (tests/test-likelihood_distribution.R:599-603: the
multinomial formula written out in R over the reported `..._model__num` and `..._obs__num`;
tests/test-likelihood_distribution-surveyindices.R:44-61 for the regression). Formulas:
R/likelihood_distribution.R:41-60 and :82-125. CPU only."""
import numpy as np
import pytest
from scipy.special import gammaln

from gadget3_b200.configs import CONFIGS, parameter_batch
from gadget3_b200.run_cuda import unpack_report
from gadget3_b200.spec import K
from oracle.oracle_lib import Oracle


def _lsa(a, b):
    m = np.maximum(a, b)
    return m + np.log1p(np.exp(-np.abs(a - b)))


def _avoid_zero(a):
    return _lsa(a * 1000.0, 0.0) / 1000.0


def _comp(ints, reals, c):
    o = ints[K["G3H_COMP_OFF"]] + c * K["G3C_LEN"]
    nd, nt, nb = (int(ints[o + K[k]]) for k in ("G3C_N_DEST", "G3C_N_TIMES", "G3C_N_BINS"))
    obs = np.array(reals[ints[o + K["G3C_OBS_ROFF"]]:ints[o + K["G3C_OBS_ROFF"]] + nt * nd]).reshape(nt, nd)
    NT = int(ints[K["G3H_NT"]])
    tidx = np.array(ints[ints[o + K["G3C_TIDX_OFF"]]:ints[o + K["G3C_TIDX_OFF"]] + NT])
    return obs, tidx, nb


@pytest.fixture(scope="module")
def c1m(oracle_lib):
    model, p = CONFIGS["C1_MULTI"]()
    ints, reals = model.pack(p)
    return model, p, ints, reals, Oracle(ints, reals)


@pytest.mark.parametrize("seed", [0, 3])
def test_multinomial_rederived_from_the_reported_model_array(c1m, seed):
    model, p, ints, reals, orc = c1m
    full = np.tile(p.values(), (1, 1))
    if seed:
        full[:, p.optimised_indices()] = parameter_batch(p, 1, seed=seed)
    rep = unpack_report(model, orc.report(full[0]))
    checked = 0
    for c, name in enumerate(model.comp_names):
        if "multinomial" not in name:
            continue
        obs, tidx, nb = _comp(ints, reals, c)
        marr = np.asarray(rep[name + "_model__num"], dtype=float)
        marr = marr.reshape(marr.shape[0], -1, nb)                  # [time][slice][length bin]
        per_step = np.asarray(rep["nll_" + name + "__num"], dtype=float)
        eps = 10.0                                                  # g3l_distribution_multinomial(epsilon = 10)
        for t, ti in enumerate(tidx):
            if ti < 0:
                assert per_step[t] == 0.0
                continue
            want = 0.0
            for sl in range(marr.shape[1]):
                dist, data = marr[ti, sl], obs[ti].reshape(-1, nb)[sl]
                floor = 1.0 / (len(data) * eps)
                prop = _lsa(dist / _avoid_zero(dist.sum()) * 10000.0, floor * 10000.0) / 10000.0    # dif_pmax(., ., 10000)
                want += 2.0 * (-(data * np.log(prop)).sum() + gammaln(1.0 + data).sum() - gammaln(1.0 + data.sum()))
            assert per_step[t] == pytest.approx(want, rel=1e-11), (name, t)
            checked += 1
    assert checked == 2 * 34


def test_surveyindices_linear_regression_rederived(c1m):
    model, p, ints, reals, orc = c1m
    rep = unpack_report(model, orc.report(p.values()))
    name = [n for n in model.comp_names if "surveyindices_linear" in n][0]
    c = model.comp_names.index(name)
    obs, tidx, _ = _comp(ints, reals, c)
    N = np.asarray(rep[name + "_model__wgt"], dtype=float).reshape(-1)
    Iv = obs.reshape(-1)
    beta = ((Iv - Iv.mean()) * (N - N.mean())).sum() / _avoid_zero(((N - N.mean()) ** 2).sum())
    alpha = Iv.mean() - beta * N.mean()
    assert np.asarray(rep[name + "_model__params"]) == pytest.approx([alpha, beta], rel=1e-10)
    per_step = np.asarray(rep["nll_" + name + "__wgt"], dtype=float)
    last = int(np.max(np.nonzero(tidx >= 0)[0]))
    assert per_step[last] == pytest.approx(((alpha + beta * N - Iv) ** 2).sum(), rel=1e-10)
    assert (np.delete(per_step, last) == 0.0).all()                # scored once, when the last index has been collected
    # the weighted sum of the per-step rows is the objective
    total = sum(float(p.values()[p.index(n + "_weight")]) * np.asarray(rep["nll_" + n + ("__wgt" if "survey" in n else "__num")]).sum()
                for n in model.comp_names)
    total += 1e8 * np.asarray(rep["nll_understocking__wgt"]).sum()
    assert rep["nll"] == pytest.approx(total, rel=1e-9)


def test_stratified_sumofsquares_rederived_from_the_reported_model_array(oracle_lib):
    """g3l_distribution_sumofsquares(over = c('area', 'length')), R/likelihood_distribution.R:8-16: proportions within each
    length group, x / avoid_zero(rowSums(x))[length], for the model and the observations alike (here: proportion mature at
    length, two stock slices). Re-derived from the oracle's REPORTed model array; and it must differ from the plain form."""
    model, p = CONFIGS["C2_STRAT"]()
    ints, reals = model.pack(p)
    rep = unpack_report(model, Oracle(ints, reals).report(p.values()))
    name = [n for n in model.comp_names if "matp" in n][0]
    c = model.comp_names.index(name)
    obs, tidx, nb = _comp(ints, reals, c)
    marr = np.asarray(rep[name + "_model__num"], dtype=float)
    marr = marr.reshape(marr.shape[0], -1, nb)                      # [time][slice][length bin]
    per_step = np.asarray(rep["nll_" + name + "__num"], dtype=float)
    checked = 0
    for t, ti in enumerate(tidx):
        if ti < 0:
            continue
        m, o = marr[ti], obs[ti].reshape(-1, nb)
        want = ((m / _avoid_zero(m.sum(axis=0)) - o / _avoid_zero(o.sum(axis=0))) ** 2).sum()
        plain = ((m / _avoid_zero(m.sum()) - o / _avoid_zero(o.sum())) ** 2).sum()
        assert per_step[t] == pytest.approx(want, rel=1e-11), t
        assert abs(want - plain) > 1e-3 * abs(want)
        checked += 1
    assert checked == 34


def test_transform_fs_age_reading_error_and_length_matrix(oracle_lib):
    """tests/test-likelihood_distribution.R:143-232 (g3l_distribution:transform_fs): `wt` collects through an age-reading error
    matrix, `nt` does not; the test applies the matrix by hand to nt -- wt[, dest age, ] = sum_k nt[, k, ] * m[k, dest age] --
    for the identity and for 'age 1 smeared over three groups, age 2 over two', checks that a matrix given per stock is the
    same thing, and that the length transform diag(10 * midlen) scales every length group of nt."""
    from gadget3_b200 import configs
    from gadget3_b200.run_cuda import unpack_report
    from oracle.oracle_lib import Oracle
    for m, msg in ((np.eye(5), "identity matrix"), (configs.READER_MATRIX, "age1, age2 smeared")):
        model, p = configs.build_mini_transform(reader=m)
        rep = unpack_report(model, Oracle(*model.pack(p)).report(p.values()))
        nt = np.asarray(rep["adist_sumofsquares_nt_model__num"])                 # [time][area][stock][age][length]
        wt = np.asarray(rep["adist_sumofsquares_wt_model__num"])
        assert nt.shape == (2, 1, 1, 5, 5) and nt.min() > 0
        np.testing.assert_array_equal(wt, np.asarray(rep["adist_sumofsquares_wtperstock_model__num"]), err_msg=msg)
        for dest in range(5):
            want = sum(nt[:, :, :, k, :] * m[k, dest] for k in range(5))
            np.testing.assert_allclose(wt[:, :, :, dest, :], want, rtol=1e-13, err_msg=f"age{dest + 1}: {msg}")
        if m is not configs.READER_MATRIX:
            np.testing.assert_array_equal(wt, nt)
        else:
            assert np.abs(wt - nt).max() > 1e-3 * nt.max()
        midlen = np.array([1.5, 2.5, 3.5, 4.5, 5.5])
        np.testing.assert_allclose(np.asarray(rep["adist_sumofsquares_len_model__num"]), nt * (10 * midlen), rtol=1e-13,
                                   err_msg="length vector applied")


def test_transform_fs_refusals():
    """A matrix of parameters (g3_param_array) cannot be folded into the gather at lowering time; other dimensions are the
    reference's own error (R/likelihood_distribution.R:326)."""
    from gadget3_b200 import configs
    from gadget3_b200.formula import G3Unsupported, g3_parameterized
    with pytest.raises(G3Unsupported, match="constant matrices"):
        configs.build_mini_transform(reader=[[g3_parameterized("reader")] * 5] * 5)
    with pytest.raises(ValueError, match="5 x 5"):
        configs.build_mini_transform(reader=np.eye(4))
