"""Shared helpers of the parity tests."""
import numpy as np

US_WEIGHT = 1e8      # g3l_understocking default weight (R/likelihood_understocking.R:7)


_US_CACHE = {}


def consumed_biomass_per_step(model, params):
    """sum over predators and prey cells of the biomass consumed at each step, at the template values
    (from the oracle's detail_<prey>_<pred>__cons report)."""
    key = id(model)
    if key not in _US_CACHE:
        from gadget3_b200.run_cuda import unpack_report
        from oracle.oracle_lib import Oracle
        ints, reals = model.pack(params)
        rep = unpack_report(model, Oracle(ints, reals).report(params.values()))
        tot = np.zeros(model.n_steps)
        for pred, prey, _ in model.pp_info:
            c = np.asarray(rep[f"detail_{prey}_{pred}__cons"], dtype=float)
            tot += np.nan_to_num(c.reshape(-1, c.shape[-1])).sum(axis=0)
        _US_CACHE[key] = tot
    return _US_CACHE[key]


def us_tolerance(model, params, comp_us, ulps=16, rtol=1e-10):
    """Allowed absolute difference of the (unweighted) g3l_understocking component: the contract's relative 1e-10
    of the term itself, plus its conditioning. Per step it adds weight * (sum(tp) - sum(tp'))^2 where both sums are the
    biomass S_t consumed at that step (R/action_predate.R:393-398): the two sums are formed in floating point, so
    their DIFFERENCE carries an absolute error of a few ulps of S_t whatever its size -- the reference's own
    backends differ by as much (R sums in long double, TMB in Eigen's packet order), and so do two summation
    orders on the GPU. With delta_t = ulps * eps * S_t the squared term moves by 2 sqrt(u_t) delta_t + delta_t^2;
    over the steps, by Cauchy-Schwarz, sum_t 2 sqrt(u_t) delta_t <= 2 sqrt(sum_t u_t * sum_t delta_t^2).
    `comp_us` is sum_t u_t (the unweighted understocking component). Measured on a B200 (tools/gpu_parity_diag.py): the
    conditioning part alone covers C2/C4/C5 with a margin of 30-100x (allowed slack on nll 1e-12 .. 3e-11 relative);
    where overconsumption takes away half the landings (C1 has such vectors, u ~ 3e5) the term follows the state's own
    relative error (observed 2e-13), which the relative part covers."""
    delta2 = float(((ulps * np.finfo(float).eps * consumed_biomass_per_step(model, params)) ** 2).sum())
    return rtol * np.abs(comp_us) + 2 * np.sqrt(np.abs(comp_us) * delta2) + delta2


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
