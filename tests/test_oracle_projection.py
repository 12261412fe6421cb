"""project_years (R/action_time.R:21-78): the loop runs that many years past end_year. The reference grows its loop at run
time; g3_to_cuda(max_project_years = K) lays every per-step table out K years further. In those years there is no
observation and no landing, by-year tables answer with their ifmissing (rec.proj) -- pinned by what that implies."""
import numpy as np
import pytest

from gadget3_b200 import G3Unsupported
from gadget3_b200.configs import CONFIGS, parameter_batch
from gadget3_b200.run_cuda import unpack_report
from oracle.oracle_lib import Oracle


def _batch(p, n, seed=3):
    full = np.tile(p.values(), (n, 1))
    opt = p.optimised_indices()
    full[:, opt] = parameter_batch(p, n, seed=seed)
    return full


def test_projection_years_on_the_vignette_model(oracle_lib):
    m0, p0 = CONFIGS["C1"]()
    m1, p1 = CONFIGS["C1_PROJ"]()          # the same actions, lowered with max_project_years = 3
    assert p0.switch == p1.switch and m1.n_steps == m0.n_steps + 3 * 4
    assert m1.time_labels[-1] == "2026-04"
    o0, o1 = Oracle(*m0.pack(p0)), Oracle(*m1.pack(p1))
    full = _batch(p0, 4)
    a, b = o0.eval_batch(full, want_comp=True), o1.eval_batch(full, want_comp=True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])        # room for a projection changes nothing by itself
    pj = p1.index("project_years")
    f2 = full.copy(); f2[:, pj] = 2
    c = o1.eval_batch(f2, want_comp=True)
    # no observation, no landing in the projected years: the objective is that of the historical run, to the bit
    assert np.array_equal(c[0], a[0]) and np.array_equal(c[1], a[1])
    f4 = full.copy(); f4[:, pj] = 4
    assert np.isnan(o1.eval_batch(f4)[0]).all()                              # past the lowered room (the library refuses it: EUNSUPP)
    rep = unpack_report(m1, o1.report(f2[0]))
    num = rep["dstart_fish__num"]                                            # [length, area, age, time]
    v = dict(zip(p1.switch, f2[0]))
    hist = unpack_report(m0, o0.report(full[0]))["dstart_fish__num"]
    assert np.array_equal(num[..., :136], hist)                              # the historical part is the historical run
    assert (num[..., 136 + 8:] == 0).all() and (num[..., 136:136 + 8].sum(axis=(0, 1, 2)) > 0).all()      # two more years, not three
    t = 136 + 4 + 2                                                          # 2025 step 3 -> step 4: only natural mortality moves numbers
    for age in range(5):
        r = num[:, 0, age, t + 1].sum() / num[:, 0, age, t].sum()
        assert r == pytest.approx(np.exp(-v[f"fish.M.{age + 1}"] * 0.25), rel=1e-13)
    # recruitment in a projected year comes from the table's ifmissing, fish.rec.proj (R/params.R:122-129)
    rn = rep["detail_fish__renewalnum"]
    t2 = 136 + 1                                                             # 2024 step 2: the renewal step
    last = 136 - 3                                                           # 2023 step 2
    assert rn[:, 0, 0, t2].sum() / rn[:, 0, 0, last].sum() == pytest.approx(v["fish.rec.proj"] / v["fish.rec.2023"], rel=1e-12)


def test_renewal_outside_projections_is_refused_when_projecting():
    from gadget3_b200 import (g3_areas, g3_stock, g3_to_cuda, g3a_age, g3a_initialconditions_normalcv, g3a_naturalmortality,
                              g3a_renewal_normalcv, g3a_time, g3s_age, g3s_livesonareas)
    st = g3s_livesonareas(g3s_age(g3_stock("st", range(10, 50, 10)), 1, 3), g3_areas(["a"]))
    acts = [g3a_time(2000, 2001, [6, 6]), g3a_initialconditions_normalcv(st), g3a_naturalmortality(st),
            g3a_renewal_normalcv(st, run_projection=False), g3a_age(st)]
    g3_to_cuda(acts)
    with pytest.raises(G3Unsupported, match="run_projection"):
        g3_to_cuda(acts, max_project_years=2)
    with pytest.raises(ValueError):
        g3_to_cuda(acts, max_project_years=-1)
