"""Time-varying maturity parameters: g3a_mature_continuous with alpha by year and l50 a
g3_timevariable (R/action_mature.R:1-42, R/params.R:122-129, R/timevariable.R:1-31). The maturing share is part of the
growth bands, so the maturity slots vary with the growth parameter sets (include/g3cuda_spec.h, G3S_MAT_OFF). Pinned by
equivalence with the constant-parameter model, whose maturation the reference-derived tests pin."""
import fnmatch

import numpy as np
import pytest

from gadget3_b200 import G3Unsupported
from gadget3_b200.configs import build_mini_spawn
from gadget3_b200.run_cuda import unpack_report
from oracle.oracle_lib import Oracle


def _set(p, d):
    v = p.values().copy()
    for k, x in d.items():
        hit = [i for i, n in enumerate(p.switch) if fnmatch.fnmatch(n, k)]
        assert hit, k
        v[hit] = x
    return v


def test_moving_ogive_against_the_constant_model(oracle_lib):
    m0, p0 = build_mini_spawn()
    m1, p1 = build_mini_spawn(matvar=True)
    o0, o1 = Oracle(*m0.pack(p0)), Oracle(*m1.pack(p1))
    early = unpack_report(m0, o0.report(_set(p0, {"fish_imm.mat.l50": 28.0, "fish_imm.mat.alpha": 0.1})))
    late = unpack_report(m0, o0.report(_set(p0, {"fish_imm.mat.l50": 34.0, "fish_imm.mat.alpha": 0.1})))
    same = unpack_report(m1, o1.report(_set(p1, {"fish_imm.mat.l50.*": 28.0, "fish_imm.mat.alpha.*": 0.1})))
    moved = unpack_report(m1, o1.report(_set(p1, {"fish_imm.mat.l50.early": 28.0, "fish_imm.mat.l50.late": 34.0,
                                                  "fish_imm.mat.alpha.*": 0.1})))
    # the same numbers in every set: the constant model to the bit, every reported array
    for k in early:
        assert np.array_equal(early[k], same[k], equal_nan=True), k
    # l50 switches in 2002 step 2 (step index 9): the state at the start of steps 0..9 is the early model's to the bit,
    # from step 10 on more fish stay immature than in the early model
    for k in ("dstart_fish_imm__num", "dstart_fish_mat__num", "dstart_fish_imm__wgt"):
        assert np.array_equal(moved[k][..., :10], early[k][..., :10]), k
    assert (moved["dstart_fish_imm__num"][..., 10].sum() > early["dstart_fish_imm__num"][..., 10].sum())
    assert (moved["dstart_fish_mat__num"][..., 10].sum() < early["dstart_fish_mat__num"][..., 10].sum())
    # one step of the late ogive applied to the early state is not the late model's state (its history differs) ...
    assert not np.array_equal(moved["dstart_fish_imm__num"][..., 10], late["dstart_fish_imm__num"][..., 10])
    # ... but alpha by year does what the constant alpha does, year by year: alpha of 2003 alone changes nothing before 2003
    a03 = unpack_report(m1, o1.report(_set(p1, {"fish_imm.mat.l50.*": 28.0, "fish_imm.mat.alpha.*": 0.1,
                                                "fish_imm.mat.alpha.2003": 0.3})))
    assert np.array_equal(a03["dstart_fish_mat__num"][..., :13], same["dstart_fish_mat__num"][..., :13])
    assert not np.array_equal(a03["dstart_fish_mat__num"][..., 13], same["dstart_fish_mat__num"][..., 13])


def test_gradient_reaches_both_sets(oracle_lib):
    m1, p1 = build_mini_spawn(matvar=True)
    o1 = Oracle(*m1.pack(p1))
    v = p1.values()
    tix = np.array([p1.index("fish_imm.mat.l50.early"), p1.index("fish_imm.mat.l50.late"), p1.index("fish_imm.mat.alpha.2004")], dtype=np.int32)
    _, _, g = o1.eval_batch(v[None, :], tangent_idx=tix)
    for j, i in enumerate(tix):
        h = 1e-5 * max(1.0, abs(v[i]))
        vp, vm = v.copy(), v.copy()
        vp[i] += h; vm[i] -= h
        fd = (o1.eval_batch(vp[None, :])[0][0] - o1.eval_batch(vm[None, :])[0][0]) / (2 * h)
        assert g[0, j] == pytest.approx(fd, rel=1e-5, abs=1e-9), p1.switch[i]
        assert g[0, j] != 0.0


def test_constant_maturity_with_moving_parameters_is_refused():
    from gadget3_b200 import (g3_areas, g3_parameterized, g3_stock, g3_to_cuda, g3a_age, g3a_grow_impl_bbinom, g3a_growmature,
                              g3a_initialconditions_normalcv, g3a_mature_constant, g3a_time, g3s_age, g3s_livesonareas)
    areas = g3_areas(["a"])
    imm = g3s_livesonareas(g3s_age(g3_stock("imm", range(10, 50, 10)), 1, 3), areas)
    mat = g3s_livesonareas(g3s_age(g3_stock("mat", range(10, 50, 10)), 1, 3), areas)
    acts = [g3a_time(2000, 2001, [6, 6]), g3a_initialconditions_normalcv(imm), g3a_initialconditions_normalcv(mat),
            g3a_growmature(imm, g3a_grow_impl_bbinom(maxlengthgroupgrowth=2),
                           maturity_f=g3a_mature_constant(alpha=0.1, l50=g3_parameterized("l50", by_stock=True, by_year=True)),
                           output_stocks=[mat], transition_f="cur_step_final"),
            g3a_growmature(mat, g3a_grow_impl_bbinom(maxlengthgroupgrowth=2)), g3a_age(imm), g3a_age(mat)]
    with pytest.raises(G3Unsupported, match="g3a_mature_constant with time-varying parameters"):
        g3_to_cuda(acts)
