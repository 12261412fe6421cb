"""Shared helpers of the parity tests."""
import numpy as np

US_WEIGHT = 1e8      # g3l_understocking default weight (R/likelihood_understocking.R:7)


def us_tolerance(model, params, comp_us):
    """Conditioning of g3l_understocking. Per step it adds weight * (sum(tp) - sum(tp'))^2 where both
    sums are ~ the landings E of that step (R/action_predate.R:393-398): a one-ulp difference in any
    exp/log1p underneath moves the DIFFERENCE by delta ~ 64 eps E, i.e. the squared term by
    2 sqrt(u_t) delta + delta^2. The reference's own backends (R sums in long double, TMB in Eigen's
    packet order) differ by this much, so this term is compared against that bound, and everything
    else at 1e-10 relative. Over T steps: sum_t 2 sqrt(u_t) delta <= 2 delta sqrt(T * sum_t u_t)."""
    from gadget3_b200.spec import K
    ints, reals = model.pack(params)
    emax = 0.0
    for pr in range(ints[K["G3H_NPRED"]]):
        rec = ints[K["G3H_PRED_OFF"]] + pr * K["G3P_LEN"]
        n = ints[K["G3H_NT"]] * ints[rec + K["G3P_NAREA"]]
        if ints[rec + K["G3P_E_ROFF"]] >= 0 and ints[rec + K["G3P_KIND"]] == K["G3PRED_TOTALFLEET"]:
            emax = max(emax, np.abs(reals[ints[rec + K["G3P_E_ROFF"]]:ints[rec + K["G3P_E_ROFF"]] + n]).max())
    # only a totalfleet's table is the consumed biomass itself (a stock predator's consumption is no data
    # table; numberfleet / linearfleet / effortfleet tables are individuals, rates, effort): bound the others by
    # the prey biomass they can take (0.95 x biomass, R/action_predate.R:383), from the oracle's own state
    # history at the template values
    if any(ints[ints[K["G3H_PRED_OFF"]] + pr * K["G3P_LEN"] + K["G3P_KIND"]] != K["G3PRED_TOTALFLEET"]
           for pr in range(ints[K["G3H_NPRED"]])):
        from gadget3_b200.run_cuda import unpack_report
        from oracle.oracle_lib import Oracle
        rep = unpack_report(model, Oracle(ints, reals).report(params.values()))
        for sn in model.stock_names:
            b = (rep[f"dstart_{sn}__num"] * rep[f"dstart_{sn}__wgt"])
            emax = max(emax, float(b.reshape(-1, b.shape[-1]).sum(axis=0).max()))
    delta = 64 * np.finfo(float).eps * emax
    T = model.n_steps
    return 2 * delta * np.sqrt(T * np.abs(comp_us)) + T * delta * delta


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
