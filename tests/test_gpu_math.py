"""The kernels' own logspace_add / avoid_zero / dif_pmin against extended-precision references (GPU).

The device code does not call libm's log1p: it evaluates log1p(exp(-|d|)) through one branch-free
series (gadget3_b200/csrc/g3_math.cuh). These tests pin that against numpy's long double (80-bit)
evaluation of the reference's formula (R/aab_env.R:16-31, R/env_dif.R:37-47) over every regime the model
visits: exact ties, tiny and huge arguments, negative numbers under avoid_zero, the early-out edges.
"""
import numpy as np
import pytest

from gadget3_b200._lib import math_probe

pytestmark = pytest.mark.gpu
LD = np.longdouble


def _ref_lsa(x, y):
    x, y = x.astype(LD), y.astype(LD)
    m = np.maximum(x, y)
    return m + np.log1p(np.exp(-np.abs(x - y)))


def _ulps(got, want, scale=None):
    """Error in units of the last place of `scale` (default: the result itself)."""
    ref = np.abs(want.astype(np.float64)) if scale is None else np.asarray(scale, dtype=np.float64)
    return np.abs(got.astype(LD) - want) / np.maximum(np.spacing(ref).astype(LD), LD(5e-324))


def test_logspace_add_matches_long_double():
    rng = np.random.default_rng(7)
    d = np.concatenate([np.linspace(-60, 60, 4001), rng.uniform(-1, 1, 2000), [0.0, 0.69, -0.69, 34.0, -34.0, 745.9, 746.1, -800.0],
                        10.0 ** rng.uniform(-12, 2.8, 3000) * rng.choice([-1, 1], 3000)])
    base = np.concatenate([rng.uniform(-20, 20, d.size // 2), 10.0 ** rng.uniform(-6, 6, d.size - d.size // 2)])
    x, y = base + d, base
    got = math_probe(0, x, y)
    want = _ref_lsa(x, y)
    # max + log1p(.) cancels when max < 0 < log1p(.) (as it does in the reference's own formula): the error is
    # bounded in ulps of the larger of the two terms, and in ulps of the result wherever nothing cancels
    terms = np.maximum(np.abs(np.maximum(x, y)), np.log1p(np.exp(-np.abs(x - y))))
    # (libm's exp + log1p would stay within 2; exp, one reciprocal and the series stay within 6: 1e-15 relative)
    assert _ulps(got, want, terms).max() <= 6.0
    nocancel = np.maximum(x, y) >= 0
    assert _ulps(got[nocancel], want[nocancel]).max() <= 4.0


def test_avoid_zero_keeps_relative_accuracy_everywhere():
    rng = np.random.default_rng(8)
    a = np.concatenate([[0.0, 1e-300, -1e-300, 5.0, -5.0, 0.034, 0.0007, -0.7], 10.0 ** rng.uniform(-12, 6, 6000),
                        -(10.0 ** rng.uniform(-12, -0.2, 3000)), rng.uniform(-0.05, 0.05, 4000)])
    got = math_probe(1, a)
    want = _ref_lsa(a * 1000.0, np.zeros_like(a)) / LD(1000)
    assert (got >= 0).all() and (got[a > -0.7] > 0).all()    # 0 only where the reference's own formula underflows (a < -0.745)
    assert _ulps(got, want).max() <= 4.0                     # incl. negative arguments, where the result is exp(1000 a) / 1000
    assert math_probe(1, np.array([0.0]))[0] == pytest.approx(np.log(2.0) / 1000.0, rel=1e-15)    # R/aab_env.R:24-31


def test_dif_pmin_matches_long_double():
    rng = np.random.default_rng(9)
    cr = np.concatenate([rng.uniform(0, 2, 5000), rng.uniform(0.9, 1.0, 3000), [0.0, 0.95, 0.9499, 0.9501, 50.0]])
    lim = np.full_like(cr, 0.95)
    got = math_probe(2, cr, lim)
    want = _ref_lsa(cr * -1e3, lim * -1e3) / LD(-1e3)
    assert _ulps(got, want, np.maximum(np.abs(want.astype(np.float64)), 0.7e-3)).max() <= 4.0
    assert (got <= 0.95).all() and (got <= cr * (1 + 1e-15)).all()      # a smooth minimum stays below both arguments
