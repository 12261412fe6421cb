"""Shared helpers of the parity tests."""
import numpy as np

from oracle.parity import US_WEIGHT, consumed_biomass_per_step, us_tolerance  # noqa: E402,F401


def golden(name):
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"reference_{name.lower()}.npz")
    return np.load(path)


def report_to_reference_layout(model, rep, key):
    """Our unpack_report() arrays -> the dimension order of the reference's REPORT()ed arrays."""
    return rep[key]


def check_detail_reports(model, rep, g, vi):
    """detail_<stock>__renewalnum, detail_<prey>_<fleet>__suit/__cons, detail_<prey>__predby_<fleet>,
    suit_<prey>_<fleet>__report (baseline/multiple-substocks.cpp:598-601,763,2045-2065)."""
    times = g["report_times"]
    n = 0
    for sn in model.stock_names:
        key = f"detail_{sn}__renewalnum"
        if f"v{vi}__{key}" in g.files:
            np.testing.assert_allclose(rep[key][..., times], g[f"v{vi}__{key}"], rtol=1e-10, atol=1e-300)
            n += 1
    for pred, prey, _ in model.pp_info:
        for key in (f"detail_{prey}_{pred}__suit", f"detail_{prey}_{pred}__cons", f"detail_{prey}__predby_{pred}"):
            if f"v{vi}__{key}" in g.files:
                ref = g[f"v{vi}__{key}"]
                np.testing.assert_allclose(rep[key][..., times], ref, rtol=1e-10, atol=1e-13 * np.abs(ref).max())
                n += 1
        key = f"suit_{prey}_{pred}__report"
        np.testing.assert_allclose(rep[key], g[f"v{vi}__{key}"], rtol=1e-12)
        n += 1
    return n
