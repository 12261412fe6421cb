"""Parity bookkeeping shared by tests/ and bench.py's untimed spot check. TEST INFRASTRUCTURE (like everything under
oracle/): never imported by the product package.

The contract (BASELINE.json north_star): nll and per-component reports agree with the reference to 1e-10 relative.
One term needs its conditioning stated: g3l_understocking adds weight * (sum(tp) - sum(tp'))^2 per step, the difference
of two sums that both equal the biomass S_t consumed at that step (R/action_predate.R:393-398). Formed in floating
point, the DIFFERENCE carries an absolute error of a few ulps of S_t whatever its own size -- the reference's backends
differ by as much (R sums in long double, TMB in Eigen's packet order)."""
import numpy as np

US_WEIGHT = 1e8      # g3l_understocking default weight (R/likelihood_understocking.R:7)


def consumed_biomass_per_step(model, params):
    """Biomass consumed at each step, summed over predators and prey cells, at the template values
    (from the oracle's detail_<prey>_<pred>__cons report)."""
    # cached ON the model: a cache keyed by id(model) hands a freed model's numbers to the next object that gets its address
    # (seen once as a spurious failure of a CPU test)
    _CACHE = model.__dict__.setdefault("_parity_cache", {})
    key = "consumed_biomass"
    if key not in _CACHE:
        from gadget3_b200.run_cuda import unpack_report
        from oracle.oracle_lib import Oracle
        ints, reals = model.pack(params)
        rep = unpack_report(model, Oracle(ints, reals).report(params.values()))
        tot = np.zeros(model.n_steps)
        for pred, prey, _ in model.pp_info:
            c = np.asarray(rep[f"detail_{prey}_{pred}__cons"], dtype=float)
            tot += np.nan_to_num(c.reshape(-1, c.shape[-1])).sum(axis=0)
        _CACHE[key] = tot
    return _CACHE[key]


def us_tolerance(model, params, comp_us, ulps=16, rtol=1e-10):
    """Allowed absolute difference of the (unweighted) understocking component `comp_us` = sum_t u_t: the contract's
    relative `rtol` of the term itself plus its conditioning. With delta_t = ulps * eps * S_t the squared term moves by
    2 sqrt(u_t) delta_t + delta_t^2; over the steps, by Cauchy-Schwarz, sum_t 2 sqrt(u_t) delta_t <=
    2 sqrt(sum_t u_t * sum_t delta_t^2). Measured on a B200 (tools/gpu_parity_diag.py): the conditioning part alone
    covers C2/C4/C5 with a margin of 30-100x (allowed slack on nll 1e-12 .. 3e-11 relative); where overconsumption takes
    away half the landings (C1 has such vectors, u ~ 3e5) the term follows the state's own relative error (observed
    2e-13), which the relative part covers."""
    delta2 = float(((ulps * np.finfo(float).eps * consumed_biomass_per_step(model, params)) ** 2).sum())
    return rtol * np.abs(comp_us) + 2 * np.sqrt(np.abs(comp_us) * delta2) + delta2


def spot_check(model, params, g_nll, g_comp, o_nll, o_comp, rtol=1e-10):
    """Numbers a bench line carries about its own results (CUDA path `g_*` against the checker `o_*`)."""
    rel = np.abs(g_nll - o_nll) / np.abs(o_nll)
    ex = np.abs((g_nll - US_WEIGHT * g_comp[:, -2]) - (o_nll - US_WEIGHT * o_comp[:, -2])) / np.abs(o_nll)
    allowed = rtol * np.abs(o_nll) + US_WEIGHT * us_tolerance(model, params, o_comp[:, -2], rtol=0.0)
    ncomp = len(model.comp_names)
    crel = np.abs(g_comp[:, :ncomp] - o_comp[:, :ncomp]) / np.maximum(np.abs(o_comp[:, :ncomp]), 1e-300)
    return {"vectors": int(len(o_nll)), "max_rel_err_nll": float(rel.max()),
            "max_rel_err_excl_understocking": float(ex.max()), "max_rel_err_components": float(crel.max()),
            "n_over_1e-10": int((rel > rtol).sum()),
            "understocking_share_of_nll_max": float(np.max(US_WEIGHT * o_comp[:, -2] / np.abs(o_nll))),
            "within_contract": bool((np.abs(g_nll - o_nll) <= allowed).all() and crel.max() < rtol)}
