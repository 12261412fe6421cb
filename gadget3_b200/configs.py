"""The BASELINE.json benchmark configurations, built through the g3a_*/g3l_* mirror API.

Model structure and parameter settings follow the reference vignettes line by line; the observation
tables are synthetic draws (numpy default_rng(123)) with the distributions the vignettes use, because
the vignettes' own data are unstored R ``rnorm``/``runif`` draws.
"""
from __future__ import annotations

from collections import OrderedDict
import numpy as np

from . import (g3_areas, g3_fleet, g3_init_val, g3_parameterized, g3_stock, g3_timeareadata, g3_to_cuda,
               g3a_age, g3a_migrate, g3a_naturalmortality_exp, g3a_predate, g3a_predate_catchability_predator, g3a_grow_impl_bbinom, g3a_grow_lengthvbsimple, g3a_grow_weightsimple,
               g3a_growmature, g3a_initialconditions_normalcv, g3a_initialconditions_normalparam,
               g3a_mature_constant, g3a_mature_continuous, g3a_naturalmortality,
               g3a_predate_catchability_totalfleet, g3a_predate_fleet, g3a_renewal_normalcv,
               g3a_renewal_normalparam, g3a_renewal_vonb_recl, g3a_report_detail, g3a_time,
               g3l_abundancedistribution, g3l_bounds_penalty, g3l_catchdistribution,
               g3l_distribution_multinomial, g3l_distribution_sumofsquares, g3l_distribution_surveyindices_linear,
               g3l_distribution_surveyindices_log, g3l_understocking,
               g3s_age, g3s_livesonareas, g3_suitability_exponentiall50)


def _cut(x, breaks):
    """cut(x, breaks = c(breaks, Inf), right = FALSE) labels."""
    edges = list(breaks) + [np.inf]
    idx = np.searchsorted(edges, x, side="right") - 1
    lab = ["[%g,%s)" % (edges[i], "Inf" if np.isinf(edges[i + 1]) else "%g" % edges[i + 1])
           for i in range(len(edges) - 1)]
    return [lab[i] if i >= 0 else None for i in idx]


def _count(cols: dict, keys):
    """group_by(keys) |> summarise(number = n())."""
    n = len(cols[keys[0]])
    tab = {}
    for i in range(n):
        k = tuple(cols[c][i] for c in keys)
        if any(v is None for v in k):
            continue
        tab[k] = tab.get(k, 0) + 1
    out = {c: [] for c in keys}
    out["number"] = []
    for k in sorted(tab, key=lambda t: tuple(str(v) if isinstance(v, str) else v for v in t)):
        for c, v in zip(keys, k):
            out[c].append(v)
        out["number"].append(tab[k])
    return out


def _fleet_data(rng, years, landings_mean, landings_sd, n_ldist, len_breaks, n_al, age_fn, al_mean_fn,
                maturity=False):
    ny = len(years)
    landings = dict(year=list(years), step=[2] * ny, area=["IXa"] * ny,
                    weight=list(rng.normal(landings_mean, landings_sd, ny)))
    yy = np.repeat(years, n_ldist)
    ln = rng.normal(50, 20, len(yy))
    ldist = _count(dict(year=[int(y) for y in yy], step=[2] * len(yy), length=_cut(ln, len_breaks)),
                   ["year", "step", "length"])
    yy = np.repeat(years, n_al)
    ages = age_fn(rng, len(yy))
    ln = rng.normal(al_mean_fn(ages), 20, len(yy))
    raw = dict(year=[int(y) for y in yy], step=[2] * len(yy), age=[int(a) for a in ages],
               length=_cut(ln, len_breaks))
    aldist = _count(raw, ["year", "step", "age", "length"])
    out = dict(landings=landings, ldist=ldist, aldist=aldist)
    if maturity:
        mat = np.asarray(aldist["age"]) > rng.uniform(3, 5, len(aldist["age"]))
        m = dict(year=aldist["year"], step=aldist["step"], length=aldist["length"],
                 stock=["mat" if x else "imm" for x in mat], number=aldist["number"])
        # duplicate (year, step, length, stock) rows are summed, as merge()/xtabs would
        agg = {}
        for i in range(len(m["year"])):
            k = (m["year"][i], m["step"][i], m["length"][i], m["stock"][i])
            agg[k] = agg.get(k, 0) + m["number"][i]
        ks = list(agg.keys())
        out["matp"] = dict(year=[k[0] for k in ks], step=[k[1] for k in ks], length=[k[2] for k in ks],
                           stock=[k[3] for k in ks], number=[agg[k] for k in ks])
    si = dict(year=list(years), step=[3] * ny, area=["IXa"] * ny,
              weight=list(rng.uniform(10000, 100000, ny)))
    out["si"] = si
    return out


def build_c1(seed=123, data=None, likelihood="sumofsquares", random=False, max_project_years=0):
    """introduction-single-stock (vignettes/introduction-single-stock.Rmd:116-596, 666-701).
    `data`: the fleet tables (landings, ldist, aldist, si as column dicts) to use instead of the synthetic
    draws -- tools/replay_tmb_bundle.py feeds the tables an R session dumped.
    `likelihood = "multinomial"`: the same model scored with g3l_distribution_multinomial for the two catch
    distributions and g3l_distribution_surveyindices_linear (slope and intercept estimated) for the index
    (R/likelihood_distribution.R:41-60,82-125).
    `random`: plus a g3l_random_dnorm and a g3l_random_walk on the yearly recruitment parameters."""
    rng = np.random.default_rng(seed)
    years = list(range(1990, 2024))
    area_names = g3_areas(["IXa", "IXb"])
    ixa = {"IXa": area_names["IXa"]}
    fish = g3s_age(g3s_livesonareas(g3_stock("fish", range(10, 101, 10)), ixa), 1, 5)
    d = data or _fleet_data(rng, years, 1000, 100, 100, range(0, 81, 20), 10000,
                            lambda r, n: np.floor(r.uniform(1, 5, n)), lambda a: 30 + a * 10)
    f_surv = g3s_livesonareas(g3_fleet("f_surv"), ixa)
    if likelihood == "multinomial":
        dist_f, si_f = g3l_distribution_multinomial, g3l_distribution_surveyindices_linear(alpha=None, beta=None)
    else:
        dist_f, si_f = g3l_distribution_sumofsquares, g3l_distribution_surveyindices_log(alpha=None, beta=1)
    actions = [
        g3a_time(1990, 2023, [3, 3, 3, 3]),
        g3a_growmature(fish, g3a_grow_impl_bbinom(maxlengthgroupgrowth=4)),
        g3a_naturalmortality(fish),
        g3a_initialconditions_normalcv(fish),
        g3a_renewal_normalcv(fish, run_step=2),
        g3a_age(fish),
        g3l_understocking([fish], nll_breakdown=True),
        g3a_predate_fleet(f_surv, [fish], suitabilities=g3_suitability_exponentiall50(),
                          catchability_f=g3a_predate_catchability_totalfleet(
                              g3_timeareadata("landings_f_surv", d["landings"], "weight", areas=area_names))),
        g3l_catchdistribution("ldist_f_surv", d["ldist"], fleets=[f_surv], stocks=[fish],
                              function_f=dist_f(), area_group=area_names,
                              report=True, nll_breakdown=True),
        g3l_catchdistribution("aldist_f_surv", d["aldist"], fleets=[f_surv], stocks=[fish],
                              function_f=dist_f(), area_group=area_names,
                              report=True, nll_breakdown=True),
        g3l_abundancedistribution("dist_si_cpue", d["si"], stocks=[fish],
                                  function_f=si_f, area_group=area_names, report=True, nll_breakdown=True),
    ]
    if random:      # recruitment deviations penalised as a random effect would be (R/likelihood_random.R:14-109)
        from . import g3l_random_dnorm, g3l_random_walk
        rec = g3_parameterized("rec", by_stock=[fish], by_year=True, scale=1e4)
        actions += [g3l_random_dnorm("rec_dnorm", rec, mean_f=g3_parameterized("rec_mean", value=1.0, lower=0.1, upper=10),
                                     sigma_f=g3_parameterized("rec_sigma", value=2.0, lower=0.5, upper=5), period="year"),
                    g3l_random_walk("rec_walk", rec, sigma_f=g3_parameterized("rec_sigma", value=2.0, lower=0.5, upper=5),
                                    period="year")]
    actions += [g3a_report_detail(actions), g3l_bounds_penalty(actions)]
    model = g3_to_cuda(actions, max_project_years=max_project_years)
    midlen = fish.midlen
    est_l50 = midlen[len(midlen) // 2 - 1]
    est_linf = max(midlen)
    est_t0 = fish.minage - 0.8
    p = model.parameter_template
    p = g3_init_val(p, "*.rec|init.scalar", 1, optimise=False)
    p = g3_init_val(p, "*.init.#", 10, lower=0.001, upper=30)
    p = g3_init_val(p, "*.rec.#", 1e-4, lower=1e-6, upper=1e-2)
    p = g3_init_val(p, "*.rec.proj", 0.002)
    p = g3_init_val(p, "init.F", 0.5, lower=0.1, upper=1)
    p = g3_init_val(p, "*.M.#", 0.15, lower=0.001, upper=10)
    p = g3_init_val(p, "*.Linf", est_linf, spread=0.2)
    p = g3_init_val(p, "*.K", 0.3, lower=0.04, upper=1.2)
    p = g3_init_val(p, "*.t0", est_t0, spread=2)
    p = g3_init_val(p, "*.walpha", 0.01, optimise=False)
    p = g3_init_val(p, "*.wbeta", 3, optimise=False)
    p = g3_init_val(p, "*.*.alpha", 0.07, lower=0.01, upper=0.2)
    p = g3_init_val(p, "*.*.l50", est_l50, spread=0.25)
    p = g3_init_val(p, "*.bbin", 1, lower=1e-05, upper=10)
    if likelihood == "multinomial":     # a linear index is scored in squared numbers: weigh it down to the size of the others
        p = g3_init_val(p, "adist_surveyindices_linear_dist_si_cpue_weight", 1e-8, optimise=False)
    return model, p


def build_c2(seed=123, data=None, mat_age=False, stratified=False):
    """multiple-substocks (vignettes/multiple-substocks.Rmd:59-439). `data`: as build_c1, plus matp.
    `mat_age`: g3a_mature_continuous with its age term, beta * cur_step_size and a50 (R/action_mature.R:18,44-74).
    `stratified`: the maturity data scored as proportions WITHIN each length group, g3l_distribution_sumofsquares(over =
    c('area', 'length')) (R/likelihood_distribution.R:8-16) -- proportion mature at length."""
    rng = np.random.default_rng(seed)
    years = list(range(1990, 2024))
    area_names = g3_areas(["IXa", "IXb"])
    ixa = {"IXa": area_names["IXa"]}
    st_imm = g3s_livesonareas(g3s_age(g3_stock({"species": "fish", "maturity": "imm"}, range(5, 101, 5)), 1, 5), ixa)
    st_mat = g3s_livesonareas(g3s_age(g3_stock({"species": "fish", "maturity": "mat"}, range(5, 101, 5)), 3, 10), ixa)
    stocks = [st_imm, st_mat]
    d = data or _fleet_data(rng, years, 1e5, 10, 1000, range(0, 111, 5), 10000,
                            lambda r, n: np.floor(np.minimum(r.normal(6, 1, n), 10)), lambda a: 30 + a * 2,
                            maturity=True)
    f_surv = g3s_livesonareas(g3_fleet("f_surv"), ixa)
    actions = [
        g3a_time(1990, 2023, [3, 3, 3, 3]),
        g3a_growmature(st_imm, g3a_grow_impl_bbinom(maxlengthgroupgrowth=4),
                       maturity_f=(g3a_mature_continuous(beta=g3_parameterized("mat.beta", by_stock=True),
                                                         a50=g3_parameterized("mat.a50", by_stock=True))
                                   if mat_age else g3a_mature_continuous()), output_stocks=[st_mat], transition_f=True),
        g3a_naturalmortality(st_imm),
        g3a_initialconditions_normalcv(st_imm),
        g3a_renewal_normalcv(st_imm),
        g3a_age(st_imm, output_stocks=[st_mat]),
        g3a_growmature(st_mat, g3a_grow_impl_bbinom(maxlengthgroupgrowth=4)),
        g3a_naturalmortality(st_mat),
        g3a_initialconditions_normalcv(st_mat),
        g3a_age(st_mat),
        g3l_understocking(stocks, nll_breakdown=True),
        g3a_predate_fleet(f_surv, stocks,
                          suitabilities=g3_suitability_exponentiall50(by_stock="species"),
                          catchability_f=g3a_predate_catchability_totalfleet(
                              g3_timeareadata("landings_f_surv", d["landings"], "weight", areas=area_names))),
        g3l_catchdistribution("ldist_f_surv", d["ldist"], fleets=[f_surv], stocks=stocks,
                              function_f=g3l_distribution_sumofsquares(), area_group=area_names,
                              report=True, nll_breakdown=True),
        g3l_catchdistribution("aldist_f_surv", d["aldist"], fleets=[f_surv], stocks=stocks,
                              function_f=g3l_distribution_sumofsquares(), area_group=area_names,
                              report=True, nll_breakdown=True),
        g3l_catchdistribution("matp_f_surv", d["matp"], fleets=[f_surv], stocks=stocks,
                              function_f=(g3l_distribution_sumofsquares(over=("area", "length")) if stratified
                                          else g3l_distribution_sumofsquares()), area_group=area_names,
                              report=True, nll_breakdown=True),
        g3l_abundancedistribution("dist_si_cpue", d["si"], stocks=stocks,
                                  function_f=g3l_distribution_surveyindices_log(alpha=None, beta=1),
                                  area_group=area_names, report=True, nll_breakdown=True),
    ]
    actions += [g3a_report_detail(actions), g3l_bounds_penalty(actions)]
    model = g3_to_cuda(actions)
    midlen = st_imm.midlen
    est_l50 = midlen[len(midlen) // 2 - 1]
    est_linf = midlen[-3]
    est_t0 = st_imm.minage - 0.8
    p = model.parameter_template
    p = g3_init_val(p, "*.rec|init.scalar", 1000, optimise=False)
    p = g3_init_val(p, "*.init.#", 10, lower=0.001, upper=10)
    p = g3_init_val(p, "*.rec.#", 100, lower=1e-6, upper=1000)
    p = g3_init_val(p, "*.rec.proj", 0.002)
    p = g3_init_val(p, "*.M.#", 0.15, lower=0.001, upper=10)
    p = g3_init_val(p, "init.F", 0.5, lower=0.1, upper=10)
    p = g3_init_val(p, "*.Linf", est_linf, spread=2)
    p = g3_init_val(p, "*.K", 0.3, lower=0.04, upper=1.2)
    p = g3_init_val(p, "*.t0", est_t0, optimise=False)
    p = g3_init_val(p, "*.walpha", 0.01, optimise=False)
    p = g3_init_val(p, "*.wbeta", 3, optimise=False)
    p = g3_init_val(p, "*.*.alpha", 0.07, lower=0.01, upper=1.8)
    p = g3_init_val(p, "*.*.l50", est_l50, spread=1.5)
    p = g3_init_val(p, "*.mat.alpha", 0.07, lower=0.0001, upper=1.8)
    p = g3_init_val(p, "*.mat.l50", est_l50, spread=3.5)
    p = g3_init_val(p, "*.bbin", 100, lower=1e-05, upper=1500)
    if mat_age:
        p = g3_init_val(p, "*.mat.beta", 0.4, lower=0.01, upper=2)
        p = g3_init_val(p, "*.mat.a50", 3, lower=1, upper=6)
    return model, p


def parameter_batch(params, B, seed=2026, width=0.10):
    """Synthetic parameter batch: each optimised parameter is the template
    value perturbed uniformly within +-width of (upper - lower), clipped to the bounds; fixed
    parameters stay at their template values. Returns B x n_optimised."""
    rng = np.random.default_rng(seed)
    idx = params.optimised_indices()
    v = params.values()[idx]
    lo = np.asarray(params.lower)[idx]
    up = np.asarray(params.upper)[idx]
    P = v[None, :] + rng.uniform(-width, width, (B, len(idx))) * (up - lo)[None, :]
    return np.clip(P, lo[None, :], up[None, :])


def build_c4(seed=123):
    """Ling-style model (structure of inttest/codegeneration/run.R:17-183 as BASELINE.json config 4
    describes it): 2 stocks on 35 length groups, maxlengthgroupgrowth 15, constant maturity with the
    hand-over on the last step of the year, 3 total-catch fleets with exponentiall50 suitability,
    one length distribution per fleet (the lln fleet's is the reference's catchdistribution_ldist_lln.txt fixture) and
    two log survey indices, 40 years of quarterly steps."""
    rng = np.random.default_rng(seed)
    years = list(range(1979, 2019))
    areas = g3_areas(["area1"])
    lg = range(20, 157, 4)
    ling_imm = g3s_livesonareas(g3s_age(g3_stock({"species": "ling", "maturity": "imm"}, lg), 3, 10), areas)
    ling_mat = g3s_livesonareas(g3s_age(g3_stock({"species": "ling", "maturity": "mat"}, lg), 5, 15), areas)
    stocks = [ling_imm, ling_mat]

    def grow():
        return g3a_grow_impl_bbinom(g3a_grow_lengthvbsimple(by_stock="species"), g3a_grow_weightsimple(),
                                    beta_f=g3_parameterized("bbin", by_stock="species"), maxlengthgroupgrowth=15)

    def init(st):
        return g3a_initialconditions_normalparam(
            st, mean_f=g3a_renewal_vonb_recl(by_stock="species"),
            stddev_f=g3_parameterized("init.sd", by_stock=True, by_age=True, value=10, optimise=False))

    actions = [
        g3a_time(1979, 2018, [3, 3, 3, 3]),
        init(ling_imm),
        g3a_renewal_normalparam(ling_imm, mean_f=g3a_renewal_vonb_recl(by_stock="species"),
                                stddev_f=g3_parameterized("rec.sd", by_stock=True, value=8, optimise=False),
                                run_step=1),
        g3a_growmature(ling_imm, grow(),
                       maturity_f=g3a_mature_constant(alpha=g3_parameterized("mat1", by_stock="species") * 0.001,
                                                      l50=g3_parameterized("mat2", by_stock="species")),
                       output_stocks=[ling_mat]),
        g3a_naturalmortality(ling_imm),
        g3a_age(ling_imm, output_stocks=[ling_mat]),
        init(ling_mat),
        g3a_growmature(ling_mat, grow()),
        g3a_naturalmortality(ling_mat),
        g3a_age(ling_mat),
        g3l_understocking(stocks, nll_breakdown=True),
    ]
    edges = list(range(20, 157, 4))
    for fi, (fname, lmean) in enumerate((("lln", 90), ("bmt", 70), ("gil", 110))):
        fleet = g3s_livesonareas(g3_fleet(fname), areas)
        yy = np.repeat(years, 4)
        st = np.tile([1, 2, 3, 4], len(years))
        landings = dict(year=[int(y) for y in yy], step=[int(x) for x in st], area=["area1"] * len(yy),
                        total_weight=list((fi + 1) * 1e2 * np.tile([1.0, 2.0, 3.0, 4.0], len(years))
                                          * rng.uniform(0.8, 1.2, len(yy))))
        actions.append(g3a_predate_fleet(
            fleet, stocks, suitabilities=g3_suitability_exponentiall50(by_stock="species"),
            catchability_f=g3a_predate_catchability_totalfleet(
                g3_timeareadata("landings_" + fname, landings, "total_weight", areas=areas))))
        ns = 300
        yy2 = np.repeat(yy, ns)
        st2 = np.repeat(st, ns)
        ln = np.clip(rng.normal(lmean, 25, len(yy2)), 20, 400)
        ld = _count(dict(year=[int(y) for y in yy2], step=[int(x) for x in st2], length=_cut(ln, edges)),
                    ["year", "step", "length"])
        lln = ling_lln_table() if fname == "lln" else None
        if lln is not None:
            # the reference's own fixture for this fleet (inttest/codegeneration/catchdistribution_ldist_lln.txt, 1994-2018),
            # its years wrapped onto the 40-year grid
            rows = {}
            for y, s_, l, n in zip(lln["year"], lln["step"], lln["length"], lln["number"]):
                rows.setdefault(y, []).append((s_, l, n))
            ld = dict(year=[], step=[], length=[], number=[])
            for Y in years:
                for s_, l, n in rows.get(1994 + (Y - years[0]) % 25, []):
                    ld["year"].append(Y); ld["step"].append(s_); ld["length"].append(l); ld["number"].append(n)
        actions.append(g3l_catchdistribution("ldist_" + fname, ld, fleets=[fleet], stocks=stocks,
                                             function_f=g3l_distribution_sumofsquares(), nll_breakdown=True))
    for nm, lab in (("si_20_52", "[20,52)"), ("si_52_inf", "[52,Inf)")):
        si = dict(year=list(years), step=[2] * len(years), area=["area1"] * len(years), length=[lab] * len(years),
                  weight=list(rng.uniform(1e5, 1e6, len(years))))
        actions.append(g3l_abundancedistribution(nm, si, stocks=stocks,
                                                 function_f=g3l_distribution_surveyindices_log(alpha=None, beta=1),
                                                 area_group=areas, nll_breakdown=True))
    actions += [g3a_report_detail(actions), g3l_bounds_penalty(actions)]
    model = g3_to_cuda(actions)
    p = model.parameter_template
    p = g3_init_val(p, "*.Linf", 160, spread=0.2)
    p = g3_init_val(p, "*.K", 0.09, lower=0.04, upper=0.2)
    p = g3_init_val(p, "*.recl", 12, lower=8, upper=20)
    p = g3_init_val(p, "recage", 3, optimise=False)
    p = g3_init_val(p, "*.walpha", 2.27567436711055e-06, optimise=False)
    p = g3_init_val(p, "*.wbeta", 3.20200445996187, optimise=False)
    p = g3_init_val(p, "*.bbin", 60, lower=1, upper=500)
    p = g3_init_val(p, "*.M.#", 0.15, lower=0.01, upper=1)
    p = g3_init_val(p, "*.init.scalar", 200, optimise=False)
    p = g3_init_val(p, "*.init.#", 1, lower=0.001, upper=10)
    p = g3_init_val(p, "init.F", 0.4, lower=0.1, upper=1)
    p = g3_init_val(p, "*.mat1", 70, lower=10, upper=200)
    p = g3_init_val(p, "*.mat2", 75, lower=40, upper=120)
    p = g3_init_val(p, "*.rec.scalar", 400, optimise=False)
    p = g3_init_val(p, "*.rec.#", 1, lower=0.001, upper=10)
    p = g3_init_val(p, "*.rec.proj", 1)
    p = g3_init_val(p, "*.*.alpha", 0.5, lower=0.01, upper=2)
    p = g3_init_val(p, "*.*.l50", 50, spread=0.5)
    return model, p


def build_c5(seed=123):
    """Four-area predator-prey model (BASELINE.json config 5): prey on 20 length
    groups x ages 1-5 and a stock predator on 8 length groups x ages 2-7, both on areas a-d, both
    migrating every quarter (12 off-diagonal migrate.<src>.<dst> parameters each), the predator eating
    the prey (g3a_predate + g3a_predate_catchability_predator + exponentiall50), one total-catch fleet on
    the predator, a length distribution for the fleet and a log survey index per area; 20 years quarterly."""
    rng = np.random.default_rng(seed)
    years = list(range(2000, 2020))
    areas = g3_areas(["a", "b", "c", "d"])
    prey = g3s_livesonareas(g3s_age(g3_stock("prey", range(5, 101, 5)), 1, 5), areas)
    pred = g3s_livesonareas(g3s_age(g3_stock("pred", range(20, 91, 10)), 2, 7), areas)

    def mig(src, dst):
        return g3_parameterized(f"migrate.{src}.{dst}", by_stock=True, value=0.3)

    fleet = g3s_livesonareas(g3_fleet("f_pred"), areas)
    yy = np.repeat(years, 4 * 4)
    st = np.tile(np.repeat([1, 2, 3, 4], 4), len(years))
    ar = np.tile(["a", "b", "c", "d"], len(years) * 4)
    landings = dict(year=[int(y) for y in yy], step=[int(x) for x in st], area=list(ar),
                    total_weight=list(rng.uniform(20, 60, len(yy))))
    ns = 200
    yl = np.repeat(np.repeat(years, 4), ns)
    sl = np.tile(np.repeat([1, 2, 3, 4], ns), len(years))
    ln = np.clip(rng.normal(55, 15, len(yl)), 20, 200)
    ldist = _count(dict(year=[int(y) for y in yl], step=[int(x) for x in sl], length=_cut(ln, range(20, 91, 10))),
                   ["year", "step", "length"])
    si = dict(year=[int(y) for y in np.repeat(years, 4)], step=[2] * (4 * len(years)),
              area=list(np.tile(["a", "b", "c", "d"], len(years))),
              weight=list(rng.uniform(1e4, 1e5, 4 * len(years))))
    actions = [
        g3a_time(2000, 2019, [3, 3, 3, 3]),
        g3a_initialconditions_normalcv(prey), g3a_initialconditions_normalcv(pred),
        g3a_renewal_normalcv(prey), g3a_renewal_normalcv(pred),
        g3a_naturalmortality(prey), g3a_naturalmortality(pred),
        g3a_growmature(prey, g3a_grow_impl_bbinom(maxlengthgroupgrowth=4)),
        g3a_growmature(pred, g3a_grow_impl_bbinom(maxlengthgroupgrowth=2)),
        g3a_age(prey), g3a_age(pred),
        g3a_migrate(prey, mig), g3a_migrate(pred, mig),
        g3a_predate(pred, [prey], suitabilities=g3_suitability_exponentiall50(),
                    catchability_f=g3a_predate_catchability_predator()),
        g3a_predate_fleet(fleet, [pred], suitabilities=g3_suitability_exponentiall50(),
                          catchability_f=g3a_predate_catchability_totalfleet(
                              g3_timeareadata("landings_f_pred", landings, "total_weight", areas=areas))),
        g3l_understocking([prey, pred], nll_breakdown=True),
        g3l_catchdistribution("ldist_f_pred", ldist, fleets=[fleet], stocks=[pred],
                              function_f=g3l_distribution_sumofsquares(), nll_breakdown=True),
        g3l_abundancedistribution("si_prey", si, stocks=[prey],
                                  function_f=g3l_distribution_surveyindices_log(alpha=None, beta=1),
                                  area_group=areas, nll_breakdown=True),
    ]
    actions += [g3a_report_detail(actions), g3l_bounds_penalty(actions)]
    model = g3_to_cuda(actions)
    p = model.parameter_template
    p = g3_init_val(p, "prey.Linf", 90, spread=0.2)
    p = g3_init_val(p, "pred.Linf", 100, spread=0.2)
    p = g3_init_val(p, "*.K", 0.25, lower=0.04, upper=1.2)
    p = g3_init_val(p, "*.t0", -0.5, optimise=False)
    p = g3_init_val(p, "*.lencv", 0.15, optimise=False)
    p = g3_init_val(p, "*.init.scalar", 100, optimise=False)
    p = g3_init_val(p, "*.init.#", 1, lower=0.001, upper=10)
    p = g3_init_val(p, "*.M.#", 0.2, lower=0.01, upper=1)
    p = g3_init_val(p, "init.F", 0.3, lower=0.1, upper=1)
    p = g3_init_val(p, "recage", 1, optimise=False)
    p = g3_init_val(p, "*.walpha", 1e-5, optimise=False)
    p = g3_init_val(p, "*.wbeta", 3, optimise=False)
    p = g3_init_val(p, "*.migrate.*", 0.2, lower=0, upper=0.5)
    p = g3_init_val(p, "*.consumption.m0", 1e-6, optimise=False)
    p = g3_init_val(p, "*.consumption.m3", 2.0, optimise=False)
    p = g3_init_val(p, "*.halffeeding", 1e4, lower=1e2, upper=1e6)
    p = g3_init_val(p, "*.energycontent", 1, optimise=False)
    p = g3_init_val(p, "*.bbin", 30, lower=1, upper=500)
    p = g3_init_val(p, "*.rec.#", 20, lower=0.1, upper=100)
    p = g3_init_val(p, "*.rec.scalar", 10, optimise=False)
    p = g3_init_val(p, "pred.rec.scalar", 1, optimise=False)
    p = g3_init_val(p, "pred.init.scalar", 10, optimise=False)
    p = g3_init_val(p, "*.rec.proj", 1)
    p = g3_init_val(p, "prey.pred.alpha", 0.15, lower=0.01, upper=1)
    p = g3_init_val(p, "prey.pred.l50", 25, lower=5, upper=60)
    p = g3_init_val(p, "pred.f_pred.alpha", 0.2, lower=0.01, upper=1)
    p = g3_init_val(p, "pred.f_pred.l50", 50, lower=20, upper=90)
    return model, p


def build_mini_andersen(seed=123, suits=False, otherfood=False, prefs=False, stomach=False):
    """Small 2-area test model: a stock predator with g3_suitability_andersen (predator-length dependent
    suitability), migration on one step only for the predator, renewal, ageing -- exercises the code
    paths config 5 does not (used by the parity tests, not by bench.py). `otherfood`: the predator also eats an
    otherfood stock (g3a_otherfood, R/action_renewal.R:276-308; one length group, abundance by year x step) with a
    constant suitability, as in the reference's tests/test-action_predate-predator.R:8,39-45."""
    from . import (g3_suitability_andersen, g3_suitability_andersenfleet, g3_suitability_gamma,
                   g3_suitability_richards, g3_suitability_straightline)
    areas = g3_areas(["a", "b"])
    prey = g3s_livesonareas(g3s_age(g3_stock("prey", range(5, 50, 5)), 1, 3), areas)
    pred = g3s_livesonareas(g3s_age(g3_stock("pred", range(20, 70, 10)), 2, 4), areas)

    def mig(src, dst):
        return g3_parameterized(f"migrate.{src}.{dst}", by_stock=True, value=0.3)

    su = g3_suitability_andersen(g3_parameterized("and.p0", value=0.0), g3_parameterized("and.p1", value=1.2),
                                 g3_parameterized("and.p2", value=1.0), g3_parameterized("and.p3", value=0.5),
                                 g3_parameterized("and.p4", value=0.8))
    if suits:       # the other suitability forms of R/suitability.R:32-94: richards for the predator, three fleets on the prey
        gp = g3_parameterized
        su = g3_suitability_richards(gp("ri.p0", value=-4.0), gp("ri.p1", value=0.15), gp("ri.p2", value=0.02),
                                     gp("ri.p3", value=0.9), gp("ri.p4", value=1.6))
    actions = [
        g3a_time(2000, 2003, [3, 3, 3, 3]),
        g3a_initialconditions_normalcv(prey), g3a_initialconditions_normalcv(pred),
        g3a_naturalmortality(prey), g3a_naturalmortality(pred),
        g3a_growmature(prey, g3a_grow_impl_bbinom(maxlengthgroupgrowth=3)),
        g3a_growmature(pred, g3a_grow_impl_bbinom(maxlengthgroupgrowth=2)),
        g3a_renewal_normalcv(prey), g3a_age(pred),
        g3l_understocking([prey]),
    ]
    # `prefs`: no ageing of the prey and no migration. Ageing leaves the youngest column exactly empty until the renewal, and
    # 0^p has no derivative (NaN in the reference's tape as well); migration passes every weight through ratio_add_pop's
    # avoid_zero, which moves the near-empty cells that biomass^p (p < 1) makes visible -- the re-derivation from the
    # state at the start of the step (tests/test_oracle_predator.py) would not see that.
    if not prefs:
        actions += [g3a_age(prey), g3a_migrate(prey, mig), g3a_migrate(pred, mig, run_steps=[2])]
    if prefs:       # a suitability that is never exactly zero: 0^p has no derivative (NaN in the reference's tape as well)
        su = g3_suitability_exponentiall50(g3_parameterized("pl.alpha", value=0.1), g3_parameterized("pl.l50", value=25))
    if otherfood:
        from . import g3a_otherfood, g3_suitability_constant
        of = g3s_livesonareas(g3s_age(g3_stock("otherfood", [0]), 0, 0), areas)
        actions += [g3a_otherfood(of),
                    g3a_predate(pred, [prey, of], suitabilities={"prey": su, "otherfood": g3_suitability_constant(0.02)},
                                catchability_f=g3a_predate_catchability_predator(
                                    # `prefs`: prey_preferences other than 1 (R/action_predate.R:208,220-225), one by prey name
                                    prey_preferences={"prey": g3_parameterized("pref.prey", value=0.9), "otherfood": 1.3} if prefs else 1))]
    else:
        actions.append(g3a_predate(pred, [prey], suitabilities=su, catchability_f=g3a_predate_catchability_predator()))
    rng = np.random.default_rng(seed)
    yrs = [2000, 2001, 2002, 2003]
    if suits:
        gp = g3_parameterized
        fsu = dict(f_gam=g3_suitability_gamma(gp("ga.alpha", value=3.0), gp("ga.beta", value=2.0), gp("ga.gamma", value=3.5)),
                   f_and=g3_suitability_andersenfleet(),
                   f_lin=g3_suitability_straightline(gp("sl.alpha", value=0.05), gp("sl.beta", value=0.01)))
        for fname, fs in fsu.items():
            fl = g3s_livesonareas(g3_fleet(fname), areas)
            land = dict(year=[y for y in yrs for _ in range(2)], step=[2] * 8, area=["a", "b"] * 4,
                        weight=list(rng.uniform(5, 20, 8)))
            actions.append(g3a_predate_fleet(fl, [prey], suitabilities=fs,
                                             catchability_f=g3a_predate_catchability_totalfleet(
                                                 g3_timeareadata("landings_" + fname, land, "weight", areas=areas))))
            if fname == "f_lin":
                continue        # (the kernels keep at most four collect vectors: two catches + two survey indices here)
            ld = dict(year=[y for y in yrs for _ in range(3)], step=[2] * 12,
                      length=["[5,20)", "[20,35)", "[35,Inf)"] * 4, number=list(rng.uniform(10, 100, 12)))
            actions.append(g3l_catchdistribution("ldist_" + fname, ld, fleets=[fl], stocks=[prey],
                                                 function_f=g3l_distribution_sumofsquares(), area_group=None))
    if stomach:     # stomach contents: what the predator ate, by prey length (g3l_catchdistribution(predators = ...))
        sd = dict(year=[y for y in yrs for _ in range(3)], step=[2] * 12,
                  length=["[5,20)", "[20,35)", "[35,Inf)"] * 4, weight=list(rng.uniform(1, 10, 12)))
        actions.append(g3l_catchdistribution("stomach_pred", sd, predators=[pred], stocks=[prey],
                                             function_f=g3l_distribution_sumofsquares(), area_group=None, report=True))
        sn = dict(year=[y for y in yrs for _ in range(2)], step=[4] * 8, area=["a", "b"] * 4, number=list(rng.uniform(100, 1000, 8)))
        actions.append(g3l_catchdistribution("eaten_num", sn, predators=[pred], stocks=[prey], area_group=areas,
                                             function_f=g3l_distribution_surveyindices_log(alpha=None, beta=1)))
    for nm, stk in (("si_prey", prey), ("si_pred", pred)):
        si = dict(year=[y for y in yrs for _ in range(2)], step=[3] * 8, area=["a", "b"] * 4,
                  number=list(rng.uniform(1e4, 1e5, 8)))
        actions.append(g3l_abundancedistribution(nm, si, stocks=[stk], area_group=areas,
                                                 function_f=g3l_distribution_surveyindices_log(alpha=None, beta=1)))
    actions += [g3l_bounds_penalty(actions)]
    model = g3_to_cuda(actions)
    p = model.parameter_template
    p = g3_init_val(p, "*.Linf", 60, spread=0.2)
    p = g3_init_val(p, "*.K", 0.3, lower=0.05, upper=1)
    p = g3_init_val(p, "*.t0", -0.5, optimise=False)
    p = g3_init_val(p, "*.lencv", 0.15, optimise=False)
    p = g3_init_val(p, "*.init.scalar", 10, optimise=False)
    p = g3_init_val(p, "*.init.#", 1, lower=0.1, upper=5)
    p = g3_init_val(p, "*.M.#", 0.2, lower=0.01, upper=1)
    p = g3_init_val(p, "init.F", 0.3, lower=0.1, upper=1)
    p = g3_init_val(p, "recage", 1, optimise=False)
    p = g3_init_val(p, "*.walpha", 1e-5, optimise=False)
    p = g3_init_val(p, "*.wbeta", 3, optimise=False)
    if not prefs:
        p = g3_init_val(p, "*.migrate.*", 0.3, lower=0, upper=0.7)
    p = g3_init_val(p, "*.consumption.m0", 1e-5, optimise=False)
    p = g3_init_val(p, "*.consumption.m3", 1.5, optimise=False)
    p = g3_init_val(p, "*.halffeeding", 1e3, lower=10, upper=1e5)
    p = g3_init_val(p, "*.bbin", 10, lower=1, upper=100)
    p = g3_init_val(p, "*.rec.#", 5, lower=0.1, upper=20)
    p = g3_init_val(p, "*.rec.scalar", 10, optimise=False)
    p = g3_init_val(p, "*.rec.proj", 1)
    if suits:
        p = g3_init_val(p, "ri.p0", -4.0, lower=-6, upper=-2)
        p = g3_init_val(p, "ri.p1", 0.15, lower=0.05, upper=0.3)
        p = g3_init_val(p, "ri.p4", 1.6, lower=0.8, upper=3)
        p = g3_init_val(p, "ga.alpha", 3.0, lower=2, upper=5)
        p = g3_init_val(p, "ga.beta", 2.0, lower=1, upper=4)
        p = g3_init_val(p, "sl.beta", 0.01, lower=0.001, upper=0.05)
        p = g3_init_val(p, "*.andersen.p1", np.log(2), lower=0.2, upper=1.5)
        p = g3_init_val(p, "*.andersen.p3_exp", 0.1, lower=-1, upper=1)
        p = g3_init_val(p, "*.andersen.p4_exp", 0.1, lower=-1, upper=1)
        return model, p
    if prefs:
        p = g3_init_val(p, "pl.alpha", 0.1, lower=0.05, upper=0.3)
        p = g3_init_val(p, "pl.l50", 25, lower=15, upper=35)
    else:
        p = g3_init_val(p, "and.p1", 1.2, lower=0.5, upper=2)
        p = g3_init_val(p, "and.p3", 0.5, lower=0.1, upper=2)
        p = g3_init_val(p, "and.p4", 0.8, lower=0.1, upper=2)
    if otherfood:
        p = g3_init_val(p, "otherfood.of_abund.#", 5e4, lower=1e3, upper=1e6)
        p = g3_init_val(p, "otherfood.of_abund.step.#", 1.0, lower=0.5, upper=2)
        p = g3_init_val(p, "otherfood.of_abund.proj", 5e4, optimise=False)
        p = g3_init_val(p, "otherfood.of_meanwgt", 0.05, lower=0.01, upper=0.2)
    if prefs:
        p = g3_init_val(p, "pref.prey", 0.9, lower=0.7, upper=1.2)
    return model, p


def build_mini_migrate(seed=123):
    """The set-up of the reference's migration test (tests/test-action_migrate.R:6-52): a stock on areas a, c, d of
    (a, b, c, d) -- it does not live in b --, ages 3-7, a spring migration d -> a on step 2 and a winter migration
    a -> c, c -> d on step 4 (two g3a_migrate actions), nothing else moving the fish. Initial numbers and weights differ
    by area (abundance factor and walpha by area) so that ratio_add_pop has something to mix."""
    allareas = g3_areas(["a", "b", "c", "d"])
    areas = OrderedDict((k, allareas[k]) for k in ("a", "c", "d"))
    st = g3s_livesonareas(g3s_age(g3_stock("stock_acd", range(10, 50, 10)), 3, 7), areas)
    gp = g3_parameterized

    def spring(src, dst):
        return gp("migrate_spring", value=0.6324555320336759) if (src, dst) == ("d", "a") else 0

    def winter(src, dst):
        return gp("migrate_winter", value=0.7745966692414834) if (src, dst) in (("a", "c"), ("c", "d")) else 0

    actions = [
        g3a_time(2000, 2004, [3, 3, 3, 3]),
        g3a_initialconditions_normalparam(st, factor_f=gp("init.fac", by_stock=True, by_area=True, value=100),
                                          mean_f=gp("init.mean", by_stock=True, value=25),
                                          stddev_f=gp("init.sd", by_stock=True, value=10),
                                          alpha_f=gp("walpha", by_stock=True, by_area=True, value=1e-5),
                                          beta_f=gp("wbeta", by_stock=True, value=3)),
        g3a_migrate(st, spring, run_steps=[2]),
        g3a_migrate(st, winter, run_steps=[4]),
    ]
    rng = np.random.default_rng(seed)
    yrs = [2000, 2001, 2002, 2003, 2004]
    si = dict(year=[y for y in yrs for _ in range(3)], step=[3] * 15, area=["a", "c", "d"] * 5,
              number=list(rng.uniform(1e2, 1e3, 15)))
    actions.append(g3l_abundancedistribution("si_acd", si, stocks=[st], area_group=areas,
                                             function_f=g3l_distribution_surveyindices_log(alpha=None, beta=1)))
    actions += [g3l_bounds_penalty(actions)]
    model = g3_to_cuda(actions)
    p = model.parameter_template
    for i, a in enumerate(("a", "c", "d")):
        p = g3_init_val(p, f"stock_acd.init.fac.{a}", 100.0 * (i + 1), lower=10, upper=1000)
        p = g3_init_val(p, f"stock_acd.walpha.{a}", 1e-5 * (i + 1), optimise=False)
    p = g3_init_val(p, "stock_acd.init.mean", 25, lower=15, upper=35)
    p = g3_init_val(p, "stock_acd.init.sd", 10, optimise=False)
    p = g3_init_val(p, "stock_acd.wbeta", 3, optimise=False)
    p = g3_init_val(p, "migrate_spring", 0.6324555320336759, lower=0.1, upper=0.8)
    p = g3_init_val(p, "migrate_winter", 0.7745966692414834, lower=0.1, upper=0.8)
    return model, p


READER_MATRIX = np.array([[0.5, 0.25, 0.25, 0, 0], [0, 0.75, 0.25, 0, 0], [0, 0, 1, 0, 0], [0, 0, 0, 1, 0], [0, 0, 0, 0, 1.0]])


def build_mini_transform(seed=123, reader=READER_MATRIX):
    """The set-up of the reference's transform_fs test (tests/test-likelihood_distribution.R:143-232): a stock with five
    length groups and ages 1-5 on two areas, and four abundance distributions by length x age over the same observations --
    `nt` plain, `wt` with an age-reading error matrix (age 1 smeared over three age groups, age 2 over two: the test's second
    parameter set), `wtperstock` the same matrix given by stock name, `len` with the length transform diag(10 * midlen).
    Here the fish also die and grow, so the four model arrays are collected from a state that moves."""
    areas = g3_areas(["1", "2"])
    st = g3s_livesonareas(g3s_age(g3_stock("prey_a", range(1, 6)), 1, 5), areas)
    rng = np.random.default_rng(seed)
    obs = dict(year=[y for y in (2000, 2001) for _ in range(25)], length=[l for _ in range(2) for l in range(1, 6) for _ in range(5)],
               age=[a for _ in range(10) for a in range(1, 6)], number=list(rng.uniform(0.0, 1.0, 50) * 1e4))
    ss = g3l_distribution_sumofsquares
    actions = [
        g3a_time(2000, 2001, [6, 6]),
        g3a_initialconditions_normalcv(st),
        g3a_naturalmortality(st),
        g3a_growmature(st, g3a_grow_impl_bbinom(maxlengthgroupgrowth=2)),
        g3l_abundancedistribution("nt", obs, stocks=[st], function_f=ss()),
        g3l_abundancedistribution("wt", obs, stocks=[st], function_f=ss(), transform_fs={"age": reader}),
        g3l_abundancedistribution("wtperstock", obs, stocks=[st], function_f=ss(), transform_fs={"age": {"prey_a": reader}}),
        g3l_abundancedistribution("len", obs, stocks=[st], function_f=ss(), transform_fs={"length": np.diag(10.0 * np.asarray(st.midlen))}),
    ]
    actions += [g3l_bounds_penalty(actions)]
    model = g3_to_cuda(actions)
    p = model.parameter_template
    p = g3_init_val(p, "*.Linf", 6, lower=4, upper=9)
    p = g3_init_val(p, "*.K", 0.3, lower=0.05, upper=1)
    p = g3_init_val(p, "*.t0", -0.5, optimise=False)
    p = g3_init_val(p, "*.lencv", 0.3, optimise=False)
    p = g3_init_val(p, "*.init.scalar", 10, optimise=False)
    p = g3_init_val(p, "*.init.#", 1, lower=0.1, upper=5)
    p = g3_init_val(p, "*.M.#", 0.2, lower=0.01, upper=1)
    p = g3_init_val(p, "init.F", 0.3, lower=0.1, upper=1)
    p = g3_init_val(p, "recage", 1, optimise=False)
    p = g3_init_val(p, "*.walpha", 1e-2, optimise=False)
    p = g3_init_val(p, "*.wbeta", 3, optimise=False)
    p = g3_init_val(p, "*.bbin", 10, lower=1, upper=100)
    return model, p


SPAWN_KINDS = ("fecundity", "simplessb", "ricker", "bevertonholt", "bevertonholt_ss3", "hockeystick")


def build_spawn_rig(seed=123):
    """The rig of the reference's spawning test (tests/test-action_spawn.R:41-88): for every recruitment function a
    mature stock whose number and mean weight are assigned every step (g3a_otherfood: abundance by year x step) spawns
    into an immature stock that nothing else touches, so what arrives in a step is that step's recruitment. One survey
    index on all the immature stocks together gives the model a likelihood."""
    from . import (g3a_otherfood, g3a_spawn, g3a_spawn_recruitment_bevertonholt, g3a_spawn_recruitment_bevertonholt_ss3,
                   g3a_spawn_recruitment_fecundity, g3a_spawn_recruitment_hockeystick, g3a_spawn_recruitment_ricker,
                   g3a_spawn_recruitment_simplessb)
    gp = g3_parameterized
    areas = g3_areas(["a"])
    rf = dict(fecundity=g3a_spawn_recruitment_fecundity(*(gp(f"fecundity_p{i}", value=0.5) for i in range(5))),
              simplessb=g3a_spawn_recruitment_simplessb(gp("simplessb_mu", value=0.5)),
              ricker=g3a_spawn_recruitment_ricker(gp("ricker_mu", value=1.0), gp("ricker_lambda", value=5e-6)),
              bevertonholt=g3a_spawn_recruitment_bevertonholt(gp("bevertonholt_mu", value=0.5), gp("bevertonholt_lambda", value=0.5)),
              bevertonholt_ss3=g3a_spawn_recruitment_bevertonholt_ss3(),
              hockeystick=g3a_spawn_recruitment_hockeystick(gp("hockeystick_r0", value=0.5), gp("hockeystick_blim", value=0.5)))
    actions = [g3a_time(1982, 1985, [3, 3, 3, 3])]
    imms = []
    for k in SPAWN_KINDS:
        mat = g3s_livesonareas(g3s_age(g3_stock("mat_" + k, range(10, 50, 10)), 5, 5), areas)
        imm = g3s_livesonareas(g3s_age(g3_stock("imm_" + k, range(10, 50, 10)), 0, 0), areas)
        imms.append(imm)
        actions += [g3a_otherfood(mat, by_stock=False), g3a_spawn(mat, rf[k], output_stocks=[imm])]
    rng = np.random.default_rng(seed)
    yrs = [1982, 1983, 1984, 1985]
    si = dict(year=yrs, step=[4] * 4, number=list(rng.uniform(1e3, 1e4, 4)))
    actions.append(g3l_abundancedistribution("si_imm", si, stocks=imms, function_f=g3l_distribution_surveyindices_log(alpha=None, beta=1)))
    actions += [g3l_bounds_penalty(actions)]
    model = g3_to_cuda(actions)
    p = model.parameter_template
    p = g3_init_val(p, "of_abund.#", 100.0, lower=50, upper=200)
    for i, y in enumerate(yrs):
        p = g3_init_val(p, f"of_abund.{y}", 100.0 + 40 * i, lower=50, upper=400)
    for st in (1, 2, 3, 4):
        p = g3_init_val(p, f"of_abund.step.{st}", 0.7 + 0.2 * st, lower=0.5, upper=2)
    p = g3_init_val(p, "of_abund.proj", 100.0, optimise=False)
    p = g3_init_val(p, "of_meanwgt", 10.0, lower=5, upper=20)
    for i in range(5):
        p = g3_init_val(p, f"fecundity_p{i}", 0.3 + 0.1 * i, lower=0.1, upper=0.9)
    p = g3_init_val(p, "simplessb_mu", 0.4, lower=0.1, upper=0.9)
    p = g3_init_val(p, "ricker_mu", 1.1, lower=0.9, upper=1.2)
    p = g3_init_val(p, "ricker_lambda", 5e-6, lower=1e-6, upper=1e-5)
    p = g3_init_val(p, "bevertonholt_mu", 0.45, lower=0.4, upper=0.6)
    p = g3_init_val(p, "bevertonholt_lambda", 0.55, lower=0.4, upper=0.6)
    p = g3_init_val(p, "hockeystick_r0", 0.5, lower=0.4, upper=0.6)
    p = g3_init_val(p, "hockeystick_blim", 5000.0, lower=1000, upper=10000)      # (inside the range of s: both arms of the stick)
    p = g3_init_val(p, "mat_*.spawn.h", 0.7, lower=0.3, upper=0.95)
    p = g3_init_val(p, "mat_*.spawn.R0", 1e4, optimise=False)
    p = g3_init_val(p, "mat_*.spawn.R_exp.#", 0.1, lower=-1, upper=1)
    p = g3_init_val(p, "mat_*.spawn.B0", 1e3, lower=500, upper=2000)
    p = g3_init_val(p, "*.Linf", 50, optimise=False)
    p = g3_init_val(p, "*.K", 0.3, optimise=False)
    p = g3_init_val(p, "*.t0", -2.0, optimise=False)
    p = g3_init_val(p, "*.rec.sd", 10, optimise=False)
    p = g3_init_val(p, "*.walpha", 1e-5, optimise=False)
    p = g3_init_val(p, "*.wbeta", 3, optimise=False)
    return model, p


def build_mini_spawn(seed=123, timevar=False, matvar=False):
    """A closed life cycle on two areas: the immature stock grows, matures into the mature stock (continuous maturity,
    every step) and ages into it; the mature stock spawns on step 2 -- proportion spawning by length
    (g3_suitability_exponentiall50), Ricker recruitment, spawning mortality -- into the immature stock's youngest age
    (g3a_spawn, R/action_spawn.R:86-233). A fleet, two likelihood components, understocking.
    `timevar`: time-varying natural mortality -- the immature stock's M by age AND year
    (g3_parameterized(by_year = TRUE), R/params.R:122-129), the mature stock's M a g3_timevariable that steps up in
    2002 step 3 and again in 2004 (R/timevariable.R:1-31).
    `matvar`: time-varying maturity parameters (g3a_mature_continuous's alpha by year, l50 a g3_timevariable)."""
    from . import g3a_spawn, g3a_spawn_recruitment_ricker, g3_timevariable, g3a_naturalmortality_exp
    areas = g3_areas(["a", "b"])
    imm = g3s_livesonareas(g3s_age(g3_stock({"species": "fish", "maturity": "imm"}, range(10, 60, 5)), 1, 3), areas)
    mat = g3s_livesonareas(g3s_age(g3_stock({"species": "fish", "maturity": "mat"}, range(10, 60, 5)), 2, 5), areas)
    fl = g3s_livesonareas(g3_fleet("f_comm"), areas)
    rng = np.random.default_rng(seed)
    yrs = [2000, 2001, 2002, 2003, 2004]
    land = dict(year=[y for y in yrs for _ in range(8)], step=[st for _ in yrs for st in (1, 2, 3, 4) for _ in range(2)],
                area=["a", "b"] * 20, weight=list(rng.uniform(20, 60, 40)))
    actions = [
        g3a_time(2000, 2004, [3, 3, 3, 3]),
        g3a_initialconditions_normalcv(imm), g3a_initialconditions_normalcv(mat),
        (g3a_naturalmortality(imm, g3a_naturalmortality_exp(g3_parameterized("Mt", by_stock=True, by_age=True, by_year=True)))
         if timevar else g3a_naturalmortality(imm)),
        (g3a_naturalmortality(mat, g3a_naturalmortality_exp(g3_timevariable("matM", OrderedDict([
            ("init", g3_parameterized("M.early", by_stock=True, value=0.2)),
            ("2002-03", g3_parameterized("M.mid", by_stock=True, value=0.3)),
            ("2004", g3_parameterized("M.late", by_stock=True, value=0.15))]))))
         if timevar else g3a_naturalmortality(mat)),
        g3a_growmature(imm, g3a_grow_impl_bbinom(maxlengthgroupgrowth=3),
                       # `matvar`: the maturation ogive moves -- l50 a g3_timevariable that switches in 2002 step 2, alpha by year
                       maturity_f=(g3a_mature_continuous(
                           alpha=g3_parameterized("mat.alpha", by_stock=True, by_year=True),
                           l50=g3_timevariable("matl50", OrderedDict([
                               ("init", g3_parameterized("mat.l50.early", by_stock=True, value=30)),
                               ("2002-02", g3_parameterized("mat.l50.late", by_stock=True, value=30))])))
                                   if matvar else g3a_mature_continuous()),
                       output_stocks=[mat], transition_f=True),
        g3a_growmature(mat, g3a_grow_impl_bbinom(maxlengthgroupgrowth=3)),
        g3a_age(imm, output_stocks=[mat]), g3a_age(mat),
        g3a_spawn(mat, g3a_spawn_recruitment_ricker(), proportion_f=g3_suitability_exponentiall50(),
                  mortality_f=g3_parameterized("spawn.mort", by_stock=True, value=0.1), output_stocks=[imm], run_step=2,
                  # (lighter than the fish already there: the mean weight of the cells they join moves)
                  alpha_f=g3_parameterized("spawn.walpha", by_stock=True, value=0.8e-5, optimise=False)),
        g3a_predate_fleet(fl, [imm, mat],
                          # `timevar`: the fleet's l50 by year (a selectivity that changes over time)
                          suitabilities=(g3_suitability_exponentiall50(l50=g3_parameterized("l50", by_stock="species", by_predator=True, by_year=True),
                                                                       by_stock="species") if timevar
                                         else g3_suitability_exponentiall50(by_stock="species")),
                          catchability_f=g3a_predate_catchability_totalfleet(g3_timeareadata("landings", land, "weight", areas=areas))),
        g3l_understocking([imm, mat]),
    ]
    ld = dict(year=[y for y in yrs for _ in range(5)], step=[2] * 25, length=["[10,20)", "[20,30)", "[30,40)", "[40,50)", "[50,Inf)"] * 5,
              number=list(rng.uniform(10, 100, 25)))
    actions.append(g3l_catchdistribution("ldist", ld, fleets=[fl], stocks=[imm, mat], function_f=g3l_distribution_sumofsquares(),
                                         area_group=None))
    si = dict(year=[y for y in yrs for _ in range(2)], step=[3] * 10, area=["a", "b"] * 5, number=list(rng.uniform(1e4, 1e5, 10)))
    actions.append(g3l_abundancedistribution("si_imm", si, stocks=[imm], area_group=areas,
                                             function_f=g3l_distribution_surveyindices_log(alpha=None, beta=1)))
    actions += [g3l_bounds_penalty(actions)]
    model = g3_to_cuda(actions)
    p = model.parameter_template
    p = g3_init_val(p, "*.Linf", 60, spread=0.2)
    p = g3_init_val(p, "*.K", 0.3, lower=0.05, upper=1)
    p = g3_init_val(p, "*.t0", -0.5, optimise=False)
    p = g3_init_val(p, "*.lencv", 0.15, optimise=False)
    p = g3_init_val(p, "*.init.scalar", 10, optimise=False)
    p = g3_init_val(p, "*.init.#", 1, lower=0.1, upper=5)
    if timevar:
        for y in yrs:
            for a in (1, 2, 3):
                p = g3_init_val(p, f"fish_imm.Mt.{y}.{a}", 0.15 + 0.02 * a + 0.03 * (y - 2000), lower=0.01, upper=1)
        p = g3_init_val(p, "fish_mat.M.early", 0.2, lower=0.05, upper=0.5)
        p = g3_init_val(p, "fish_mat.M.mid", 0.3, lower=0.05, upper=0.6)
        p = g3_init_val(p, "fish_mat.M.late", 0.15, lower=0.05, upper=0.5)
        p = g3_init_val(p, "*.M.#", 0.2, optimise=False)                # (the initial conditions' M-based abundance)
    else:
        p = g3_init_val(p, "*.M.#", 0.2, lower=0.01, upper=1)
    p = g3_init_val(p, "init.F", 0.3, lower=0.1, upper=1)
    p = g3_init_val(p, "recage", 1, optimise=False)
    p = g3_init_val(p, "*.walpha", 1e-5, optimise=False)
    p = g3_init_val(p, "*.wbeta", 3, optimise=False)
    p = g3_init_val(p, "*.bbin", 10, lower=1, upper=100)
    if matvar:
        p = g3_init_val(p, "*.mat.alpha.#", 0.1, lower=0.01, upper=1)
        p = g3_init_val(p, "*.mat.l50.early", 28, lower=20, upper=40)
        p = g3_init_val(p, "*.mat.l50.late", 34, lower=20, upper=40)
    else:
        p = g3_init_val(p, "*.mat.alpha", 0.1, lower=0.01, upper=1)
        p = g3_init_val(p, "*.mat.l50", 30, lower=20, upper=40)
    p = g3_init_val(p, "*.spawn.alpha", 0.2, lower=0.05, upper=1)
    p = g3_init_val(p, "*.spawn.l50", 35, lower=25, upper=45)
    p = g3_init_val(p, "*.spawn.mort", 0.1, lower=0.01, upper=0.5)
    p = g3_init_val(p, "*.spawn.walpha", 0.8e-5, optimise=False)
    p = g3_init_val(p, "*.spawn.mu", 0.5, lower=0.05, upper=2)
    p = g3_init_val(p, "*.spawn.lambda", 1e-6, lower=1e-7, upper=5e-6)
    p = g3_init_val(p, "*.rec.sd", 3, lower=1, upper=6)
    p = g3_init_val(p, "fish.f_comm.alpha", 0.1, lower=0.01, upper=1)
    if timevar:
        for y in yrs:
            p = g3_init_val(p, f"fish.f_comm.l50.{y}", 26 + 2 * (y - 2000), lower=15, upper=45)
    else:
        p = g3_init_val(p, "fish.f_comm.l50", 30, lower=15, upper=45)
    return model, p


def build_timevar_rig(seed=123):
    """Two stocks that only die, with time-varying natural mortality: `st_year` by age, year and
    step (g3_parameterized(by_year = TRUE, by_step = TRUE), R/params.R:122-129), `st_tv` through a g3_timevariable that
    switches in 2001 step 2 and in 2002 (R/timevariable.R:1-31). The numbers at the start of a step over those of the step
    before must be exp(-M dt) of the M in force (tests/test_oracle_kat.py)."""
    from . import g3_timevariable, g3a_naturalmortality_exp
    gp = g3_parameterized
    areas = g3_areas(["a"])
    sy = g3s_livesonareas(g3s_age(g3_stock("st_year", range(10, 50, 10)), 1, 2), areas)
    sv = g3s_livesonareas(g3s_age(g3_stock("st_tv", range(10, 50, 10)), 1, 1), areas)
    tv = g3_timevariable("M", OrderedDict([("init", gp("tvM.init", value=0.1)), ("2001-02", gp("tvM.a", value=0.4)),
                                           ("2002", gp("tvM.b", value=0.2) * 1.5)]))
    # a third stock that only grows: Linf by year, K a g3_timevariable (the growth matrices follow the parameters in force)
    sg = g3s_livesonareas(g3s_age(g3_stock("st_grow", range(10, 60, 5)), 1, 2), areas)
    tk = g3_timevariable("K", OrderedDict([("init", gp("tvK.init", value=0.2)), ("2001", gp("tvK.a", value=0.5)), ("2002-03", gp("tvK.b", value=0.1))]))
    actions = [
        g3a_time(2000, 2002, [3, 3, 6]),
        g3a_initialconditions_normalcv(sy), g3a_initialconditions_normalcv(sv), g3a_initialconditions_normalcv(sg),
        g3a_naturalmortality(sy, g3a_naturalmortality_exp(gp("Mt", by_stock=True, by_age=True, by_year=True, by_step=True))),
        g3a_naturalmortality(sv, g3a_naturalmortality_exp(tv)),
        g3a_growmature(sg, g3a_grow_impl_bbinom(g3a_grow_lengthvbsimple(gp("Linf_t", by_stock=True, by_year=True), tk),
                                                g3a_grow_weightsimple(), maxlengthgroupgrowth=3)),
    ]
    rng = np.random.default_rng(seed)
    yrs = [2000, 2001, 2002]
    si = dict(year=yrs, step=[3] * 3, number=list(rng.uniform(1e3, 1e4, 3)))
    actions.append(g3l_abundancedistribution("si", si, stocks=[sy, sv], function_f=g3l_distribution_surveyindices_log(alpha=None, beta=1)))
    ld = dict(year=[y for y in yrs for _ in range(5)], step=[3] * 15, length=["[10,20)", "[20,30)", "[30,40)", "[40,50)", "[50,Inf)"] * 3,
              number=list(rng.uniform(10, 100, 15)))
    actions.append(g3l_abundancedistribution("ldist_grow", ld, stocks=[sg], function_f=g3l_distribution_sumofsquares()))
    actions += [g3l_bounds_penalty(actions)]
    model = g3_to_cuda(actions)
    p = model.parameter_template
    for y in yrs:
        for st in (1, 2, 3):
            for a in (1, 2):
                p = g3_init_val(p, f"st_year.Mt.{y}.{st}.{a}", 0.1 + 0.05 * (y - 2000) + 0.02 * st + 0.01 * a, lower=0.01, upper=1)
    for i, y in enumerate(yrs):
        p = g3_init_val(p, f"st_grow.Linf_t.{y}", 55 + 10 * i, lower=40, upper=90)
    p = g3_init_val(p, "tvK.init", 0.2, lower=0.05, upper=1)
    p = g3_init_val(p, "tvK.a", 0.5, lower=0.05, upper=1)
    p = g3_init_val(p, "tvK.b", 0.1, lower=0.05, upper=1)
    p = g3_init_val(p, "*.bbin", 10, lower=1, upper=100)
    p = g3_init_val(p, "tvM.init", 0.1, lower=0.01, upper=1)
    p = g3_init_val(p, "tvM.a", 0.4, lower=0.01, upper=1)
    p = g3_init_val(p, "tvM.b", 0.2, lower=0.01, upper=1)
    p = g3_init_val(p, "*.Linf", 50, optimise=False)
    p = g3_init_val(p, "*.K", 0.3, lower=0.1, upper=0.6)
    p = g3_init_val(p, "*.t0", -1.0, optimise=False)
    p = g3_init_val(p, "*.lencv", 0.3, optimise=False)
    p = g3_init_val(p, "*.init.scalar", 10, optimise=False)
    p = g3_init_val(p, "*.init.#", 1, lower=0.1, upper=5)
    p = g3_init_val(p, "*.M.#", 0.2, optimise=False)
    p = g3_init_val(p, "init.F", 0.3, optimise=False)
    p = g3_init_val(p, "recage", 1, optimise=False)
    p = g3_init_val(p, "*.walpha", 1e-5, optimise=False)
    p = g3_init_val(p, "*.wbeta", 3, optimise=False)
    return model, p


def build_random_rig(seed=123):
    """The rig of the reference's random-effect test (tests/test-likelihood_random.R:6-34): 1990-1994 in quarters, a
    g3l_random_dnorm on the log scale and one on the linear scale (single parameters: scored once, at the last step), a
    g3l_random_walk over a by-year table and one over a by-year-by-step table, each switched on by its own weight
    parameter. The reference's rig has no stock at all; here one stock that only dies keeps the planner busy (it adds
    nothing to the objective)."""
    from . import g3a_naturalmortality_exp, g3l_random_dnorm, g3l_random_walk
    gp = g3_parameterized
    areas = g3_areas(["a"])
    st = g3s_livesonareas(g3s_age(g3_stock("st", range(10, 50, 10)), 1, 2), areas)
    actions = [
        g3a_time(1990, 1994, [3, 3, 3, 3]),
        g3a_initialconditions_normalcv(st),
        g3a_naturalmortality(st, g3a_naturalmortality_exp(gp("M", by_stock=True))),
        g3l_random_dnorm("dnorm_log", gp("dnorm_log", value=0.0), mean_f=gp("dnorm_log_mean", value=0.0, optimise=False),
                         sigma_f=gp("dnorm_log_sigma", value=1.0, optimise=False),
                         weight=gp("dnorm_log_enabled", value=0.0, optimise=False)),
        g3l_random_dnorm("dnorm_lin", gp("dnorm_lin", value=0.0), mean_f=gp("dnorm_lin_mean", value=0.0, optimise=False),
                         sigma_f=gp("dnorm_lin_sigma", value=1.0, optimise=False), log_f=False,
                         weight=gp("dnorm_lin_enabled", value=0.0, optimise=False)),
        g3l_random_walk("walk_year", gp("walk_year", by_year=True, value=0.0),
                        sigma_f=gp("walk_year_sigma", value=1.0, optimise=False),
                        weight=gp("walk_year_enabled", value=0.0, optimise=False)),
        g3l_random_walk("walk_step", gp("walk_step", by_year=True, by_step=True, value=0.0),
                        sigma_f=gp("walk_step_sigma", value=1.0, optimise=False),
                        weight=gp("walk_step_enabled", value=0.0, optimise=False)),
    ]
    model = g3_to_cuda(actions)
    rng = np.random.default_rng(seed)
    p = model.parameter_template
    for n in ("dnorm_log", "dnorm_lin"):
        p = g3_init_val(p, n, float(rng.uniform(0, 10)), lower=0, upper=10)
        p = g3_init_val(p, n + "_mean", float(rng.uniform(0, 10)), optimise=False)
        p = g3_init_val(p, n + "_sigma", float(rng.uniform(0.5, 10)), optimise=False)
        p = g3_init_val(p, n + "_enabled", 1, optimise=False)
    for y in range(1990, 1995):
        p = g3_init_val(p, f"walk_year.{y}", float(rng.uniform(1, 10)), lower=1, upper=10)
        for stp in range(1, 5):
            p = g3_init_val(p, f"walk_step.{y}.{stp}", float(rng.uniform(1, 10)), lower=1, upper=10)
    p = g3_init_val(p, "walk_year_sigma", 1.7, optimise=False)
    p = g3_init_val(p, "walk_step_sigma", 0.6, optimise=False)
    p = g3_init_val(p, "walk_year_enabled", 1, optimise=False)
    p = g3_init_val(p, "walk_step_enabled", 1, optimise=False)
    p = g3_init_val(p, "st.M", 0.2, lower=0.05, upper=0.5)
    p = g3_init_val(p, "*.Linf", 50, optimise=False)
    p = g3_init_val(p, "*.K", 0.3, optimise=False)
    p = g3_init_val(p, "*.t0", -1.0, optimise=False)
    p = g3_init_val(p, "*.lencv", 0.3, optimise=False)
    p = g3_init_val(p, "*.init.scalar", 10, optimise=False)
    p = g3_init_val(p, "*.init.#", 1, lower=0.1, upper=5)
    p = g3_init_val(p, "*.M.#", 0.2, optimise=False)
    p = g3_init_val(p, "init.F", 0.3, optimise=False)
    p = g3_init_val(p, "recage", 1, optimise=False)
    p = g3_init_val(p, "*.walpha", 1e-5, optimise=False)
    p = g3_init_val(p, "*.wbeta", 3, optimise=False)
    return model, p


def build_six_fleets(seed=123):
    """Three stocks and six fleets (the sizes real models reach -- gadget-models' ling and bling
    run five to seven fleets): the imm / mat pair of multiple-substocks plus a second species, four commercial fleets on
    the pair in different quarters, one on the second species, a survey on everything; a length and an age-length
    distribution per fleet, maturity proportions, two survey indices: 15 likelihood components, 9 collect vectors, 5
    predators on one stock. Ten years in quarters."""
    rng = np.random.default_rng(seed)
    years = list(range(1990, 2000))
    area_names = g3_areas(["IXa", "IXb"])
    ixa = {"IXa": area_names["IXa"]}
    st_imm = g3s_livesonareas(g3s_age(g3_stock({"species": "fish", "maturity": "imm"}, range(5, 101, 5)), 1, 5), ixa)
    st_mat = g3s_livesonareas(g3s_age(g3_stock({"species": "fish", "maturity": "mat"}, range(5, 101, 5)), 3, 10), ixa)
    st_cod = g3s_livesonareas(g3s_age(g3_stock({"species": "cod", "maturity": "all"}, range(10, 131, 8)), 1, 8), ixa)
    pair = [st_imm, st_mat]
    actions = [
        g3a_time(1990, 1999, [3, 3, 3, 3]),
        g3a_growmature(st_imm, g3a_grow_impl_bbinom(maxlengthgroupgrowth=4), maturity_f=g3a_mature_continuous(),
                       output_stocks=[st_mat], transition_f=True),
        g3a_naturalmortality(st_imm), g3a_initialconditions_normalcv(st_imm), g3a_renewal_normalcv(st_imm),
        g3a_age(st_imm, output_stocks=[st_mat]),
        g3a_growmature(st_mat, g3a_grow_impl_bbinom(maxlengthgroupgrowth=4)),
        g3a_naturalmortality(st_mat), g3a_initialconditions_normalcv(st_mat), g3a_age(st_mat),
        g3a_growmature(st_cod, g3a_grow_impl_bbinom(maxlengthgroupgrowth=3)),
        g3a_naturalmortality(st_cod), g3a_initialconditions_normalcv(st_cod), g3a_renewal_normalcv(st_cod), g3a_age(st_cod),
        g3l_understocking(pair + [st_cod], nll_breakdown=True),
    ]
    fleets = [("f_comm1", pair, 1, 2e4), ("f_comm2", pair, 2, 3e4), ("f_comm3", pair, 3, 1e4), ("f_comm4", pair, 4, 2e4),
              ("f_cod", [st_cod], 2, 1e4), ("f_surv", pair + [st_cod], 3, 500)]
    for name, stocks, step, land in fleets:
        d = _fleet_data(rng, years, land, land / 100, 300, range(0, 121, 10), 2000,
                        lambda r, n: np.floor(np.minimum(r.normal(5, 1.5, n).clip(1), 10)), lambda a: 30 + a * 3,
                        maturity=(name == "f_surv"))
        for tab in d.values():
            if "step" in tab:
                tab["step"] = [step] * len(tab["step"])
        fl = g3s_livesonareas(g3_fleet(name), ixa)
        actions += [
            g3a_predate_fleet(fl, stocks, suitabilities=g3_suitability_exponentiall50(by_stock="species"),
                              catchability_f=g3a_predate_catchability_totalfleet(
                                  g3_timeareadata("landings_" + name, d["landings"], "weight", areas=area_names))),
            g3l_catchdistribution("ldist_" + name, d["ldist"], fleets=[fl], stocks=stocks,
                                  function_f=g3l_distribution_sumofsquares(), area_group=area_names, nll_breakdown=True),
            g3l_catchdistribution("aldist_" + name, d["aldist"], fleets=[fl], stocks=stocks,
                                  function_f=g3l_distribution_sumofsquares(), area_group=area_names, nll_breakdown=True),
        ]
        if name == "f_surv":
            matp = {k: [v for v, st_ in zip(vals, d["matp"]["stock"])] for k, vals in d["matp"].items()}
            actions.append(g3l_catchdistribution("matp_f_surv", matp, fleets=[fl], stocks=pair,
                                                 function_f=g3l_distribution_sumofsquares(), area_group=area_names, nll_breakdown=True))
            actions.append(g3l_abundancedistribution("si_fish", d["si"], stocks=pair,
                                                     function_f=g3l_distribution_surveyindices_log(alpha=None, beta=1),
                                                     area_group=area_names, nll_breakdown=True))
            si2 = dict(d["si"]); si2["weight"] = list(rng.uniform(1e4, 1e5, len(years)))
            actions.append(g3l_abundancedistribution("si_cod", si2, stocks=[st_cod],
                                                     function_f=g3l_distribution_surveyindices_log(alpha=None, beta=1),
                                                     area_group=area_names, nll_breakdown=True))
    actions += [g3a_report_detail(actions), g3l_bounds_penalty(actions)]
    model = g3_to_cuda(actions)
    p = model.parameter_template
    p = g3_init_val(p, "*.rec|init.scalar", 1000, optimise=False)
    p = g3_init_val(p, "*.init.#", 10, lower=0.001, upper=10)
    p = g3_init_val(p, "*.rec.#", 100, lower=1e-6, upper=1000)
    p = g3_init_val(p, "*.rec.proj", 0.002)
    p = g3_init_val(p, "*.M.#", 0.15, lower=0.001, upper=10)
    p = g3_init_val(p, "init.F", 0.5, lower=0.1, upper=10)
    p = g3_init_val(p, "fish*.Linf", 85, spread=0.2)
    p = g3_init_val(p, "cod*.Linf", 120, spread=0.2)
    p = g3_init_val(p, "*.K", 0.3, lower=0.04, upper=1.2)
    p = g3_init_val(p, "fish*.t0", 0.2, optimise=False)
    p = g3_init_val(p, "cod*.t0", 0.2, optimise=False)
    p = g3_init_val(p, "*.walpha", 0.01, optimise=False)
    p = g3_init_val(p, "*.wbeta", 3, optimise=False)
    p = g3_init_val(p, "*.*.alpha", 0.07, lower=0.01, upper=1.8)
    p = g3_init_val(p, "*.*.l50", 45, spread=0.5)
    p = g3_init_val(p, "*.mat.alpha", 0.07, lower=0.0001, upper=1.8)
    p = g3_init_val(p, "*.mat.l50", 50, spread=0.5)
    p = g3_init_val(p, "*.bbin", 100, lower=1e-05, upper=1500)
    return model, p


def build_renewal_rig(seed=123):
    """A stock whose recruits change shape over the years: mean length a by-year table
    (g3_parameterized(by_year = TRUE), R/params.R:122-129), standard deviation a g3_timevariable that switches in 2002
    (R/timevariable.R:1-31), numbers by year as usual. The reported renewal (detail_st__renewalnum) of each year must be
    the normal density of that year's parameters on the length grid, normalised, times 1e4 * rec (R/action_renewal.R:56-80)
    -- tests/test_oracle_kat.py."""
    from . import g3_timevariable, g3a_renewal_normalparam
    gp = g3_parameterized
    areas = g3_areas(["a"])
    st = g3s_livesonareas(g3s_age(g3_stock("st", range(10, 55, 5)), 1, 3), areas)
    sd = g3_timevariable("recsd", OrderedDict([("init", gp("recsd.early", value=4.0)), ("2002", gp("recsd.late", value=8.0))]))
    actions = [
        g3a_time(2000, 2003, [6, 6]),
        g3a_initialconditions_normalcv(st),
        g3a_naturalmortality(st),
        g3a_renewal_normalparam(st, factor_f=gp("rec", by_stock=True, by_year=True), mean_f=gp("recl", by_stock=True, by_year=True),
                                stddev_f=sd, run_step=1),
        g3a_age(st),
    ]
    rng = np.random.default_rng(seed)
    yrs = [2000, 2001, 2002, 2003]
    si = dict(year=yrs, step=[2] * 4, number=list(rng.uniform(1e3, 1e4, 4)))
    actions.append(g3l_abundancedistribution("si", si, stocks=[st], function_f=g3l_distribution_surveyindices_log(alpha=None, beta=1)))
    ld = dict(year=[y for y in yrs for _ in range(4)], step=[2] * 16, length=["[10,20)", "[20,30)", "[30,40)", "[40,Inf)"] * 4,
              number=list(rng.uniform(10, 100, 16)))
    actions.append(g3l_abundancedistribution("ldist", ld, stocks=[st], function_f=g3l_distribution_sumofsquares()))
    actions += [g3a_report_detail(actions), g3l_bounds_penalty(actions)]
    model = g3_to_cuda(actions)
    p = model.parameter_template
    for i, y in enumerate(yrs):
        p = g3_init_val(p, f"st.rec.{y}", 1.0 + 0.5 * i, lower=0.1, upper=5)
        p = g3_init_val(p, f"st.recl.{y}", 18.0 + 3.0 * i, lower=12, upper=35)
    p = g3_init_val(p, "recsd.early", 4.0, lower=2, upper=10)
    p = g3_init_val(p, "recsd.late", 8.0, lower=2, upper=12)
    p = g3_init_val(p, "*.Linf", 50, optimise=False)
    p = g3_init_val(p, "*.K", 0.3, optimise=False)
    p = g3_init_val(p, "*.t0", -1.0, optimise=False)
    p = g3_init_val(p, "*.lencv", 0.3, optimise=False)
    p = g3_init_val(p, "*.init.scalar", 10, optimise=False)
    p = g3_init_val(p, "*.init.#", 1, lower=0.1, upper=5)
    p = g3_init_val(p, "*.M.#", 0.2, lower=0.05, upper=0.5)
    p = g3_init_val(p, "init.F", 0.3, optimise=False)
    p = g3_init_val(p, "recage", 1, optimise=False)
    p = g3_init_val(p, "*.walpha", 1e-5, optimise=False)
    p = g3_init_val(p, "*.wbeta", 3, optimise=False)
    return model, p


def build_wloss_rig(seed=123):
    """Spawning weight loss (R/action_spawn.R:166-181 -> g3a_weightloss, R/action_weightloss.R:19-44) where nothing else
    touches the parents: two mature stocks that neither grow nor die spawn in step 2 with a length-dependent proportion
    and a spawning mortality -- `mat_rel` loses a relative share of its weight (weightloss_f), `mat_min` the same down to
    a floor (weightloss_args = list(rel_loss, min_weight)). Between the start of step 2 and the start of step 3 the
    parents' number and mean weight change by the spawning formulas alone (tests/test_oracle_spawn.py)."""
    from . import g3a_spawn, g3a_spawn_recruitment_simplessb
    gp = g3_parameterized
    areas = g3_areas(["a"])
    actions = [g3a_time(2000, 2001, [3, 3, 3, 3])]
    imms = []
    for k in ("rel", "min"):
        mat = g3s_livesonareas(g3s_age(g3_stock("mat_" + k, range(10, 60, 10)), 3, 4), areas)
        imm = g3s_livesonareas(g3s_age(g3_stock("imm_" + k, range(10, 60, 10)), 0, 0), areas)
        imms.append(imm)
        wl = (dict(weightloss_f=gp("spawn.wloss", by_stock=True, value=0.15)) if k == "rel" else
              dict(weightloss_args=dict(rel_loss=gp("spawn.wloss", by_stock=True, value=0.3), min_weight=gp("spawn.minwgt", by_stock=True, value=0.02))))
        actions += [g3a_initialconditions_normalcv(mat),
                    g3a_spawn(mat, g3a_spawn_recruitment_simplessb(gp("spawn.mu", by_stock=True, value=0.5)),
                              proportion_f=g3_suitability_exponentiall50(), mortality_f=gp("spawn.mort", by_stock=True, value=0.2),
                              output_stocks=[imm], run_step=2, **wl)]
    rng = np.random.default_rng(seed)
    si = dict(year=[2000, 2001], step=[4] * 2, number=list(rng.uniform(1e3, 1e4, 2)))
    actions.append(g3l_abundancedistribution("si_imm", si, stocks=imms, function_f=g3l_distribution_surveyindices_log(alpha=None, beta=1)))
    ld = dict(year=[y for y in (2000, 2001) for _ in range(5)], step=[3] * 10, length=["[10,20)", "[20,30)", "[30,40)", "[40,50)", "[50,Inf)"] * 2,
              weight=list(rng.uniform(1, 10, 10)))
    actions.append(g3l_abundancedistribution("wdist_mat", ld, stocks=[actions[1].stock, actions[3].stock], function_f=g3l_distribution_sumofsquares()))
    actions += [g3l_bounds_penalty(actions)]
    model = g3_to_cuda(actions)
    p = model.parameter_template
    p = g3_init_val(p, "*.Linf", 70, optimise=False)
    p = g3_init_val(p, "*.K", 0.3, optimise=False)
    p = g3_init_val(p, "*.t0", -1.0, optimise=False)
    p = g3_init_val(p, "*.lencv", 0.3, optimise=False)
    p = g3_init_val(p, "*.init.scalar", 10, optimise=False)
    p = g3_init_val(p, "*.init.#", 1, lower=0.1, upper=5)
    p = g3_init_val(p, "*.M.#", 0.2, optimise=False)
    p = g3_init_val(p, "init.F", 0.3, optimise=False)
    p = g3_init_val(p, "recage", 0, optimise=False)
    p = g3_init_val(p, "*.walpha", 1e-5, optimise=False)
    p = g3_init_val(p, "*.wbeta", 3, optimise=False)
    p = g3_init_val(p, "*.rec.sd", 5, optimise=False)
    p = g3_init_val(p, "*.spawn.alpha", 0.15, lower=0.05, upper=1)
    p = g3_init_val(p, "*.spawn.l50", 35, lower=25, upper=45)
    p = g3_init_val(p, "*.spawn.mort", 0.2, lower=0.01, upper=0.5)
    p = g3_init_val(p, "*.spawn.mu", 0.5, lower=0.05, upper=2)
    p = g3_init_val(p, "mat_rel.spawn.wloss", 0.15, lower=0.01, upper=0.5)
    p = g3_init_val(p, "mat_min.spawn.wloss", 0.3, lower=0.01, upper=0.5)
    p = g3_init_val(p, "mat_min.spawn.minwgt", 0.02, lower=0.001, upper=0.1)
    return model, p


def build_mini_fleets(seed=123):
    """One stock on two areas fished by a numberfleet, a linearfleet and an effortfleet
    (R/action_predate.R:7-18,117-126) -- parity-test model for the catchabilities bench configs do not use."""
    from . import (g3a_predate_catchability_effortfleet, g3a_predate_catchability_linearfleet,
                   g3a_predate_catchability_numberfleet)
    areas = g3_areas(["a", "b"])
    fish = g3s_livesonareas(g3s_age(g3_stock("fish", range(5, 50, 5)), 1, 3), areas)
    rng = np.random.default_rng(seed)
    yrs = [2000, 2001, 2002, 2003]

    def table(name, lo, hi, steps):
        n = len(yrs) * len(steps) * 2
        d = dict(year=[y for y in yrs for _ in range(2 * len(steps))], step=[st for st in steps for _ in range(2)] * len(yrs),
                 area=["a", "b"] * (len(yrs) * len(steps)), amount=list(rng.uniform(lo, hi, n)))
        return g3_timeareadata(name, d, "amount", areas=areas)

    fleets = dict(f_num=g3a_predate_catchability_numberfleet(table("n_f_num", 50, 200, [2])),
                  f_lin=g3a_predate_catchability_linearfleet(table("r_f_lin", 0.1, 0.3, [1, 2, 3, 4])),
                  f_eff=g3a_predate_catchability_effortfleet(
                      g3_parameterized("catchability", by_stock=True, by_predator=True, value=0.02),
                      table("e_f_eff", 5, 10, [3, 4])))
    actions = [
        g3a_time(2000, 2003, [3, 3, 3, 3]),
        g3a_initialconditions_normalcv(fish), g3a_naturalmortality(fish),
        g3a_growmature(fish, g3a_grow_impl_bbinom(maxlengthgroupgrowth=3)),
        g3a_renewal_normalcv(fish), g3a_age(fish),
        g3l_understocking([fish]),
    ]
    for fname, cat in fleets.items():
        fl = g3s_livesonareas(g3_fleet(fname), areas)
        actions.append(g3a_predate_fleet(fl, [fish], suitabilities=g3_suitability_exponentiall50(), catchability_f=cat))
        ld = dict(year=[y for y in yrs for _ in range(3)], step=[2 if fname == "f_num" else 3] * 12,
                  length=["[5,20)", "[20,35)", "[35,Inf)"] * 4, number=list(rng.uniform(10, 100, 12)))
        if fname == "f_lin":        # catch in kilos per year and area: sum of squared log ratios (R/likelihood_distribution.R:127-136)
            from . import g3l_distribution_sumofsquaredlogratios
            kg = dict(year=[y for y in yrs for _ in range(2)], step=[3] * 8, area=["a", "b"] * 4,
                      weight=list(rng.uniform(500, 5000, 8)))
            actions.append(g3l_catchdistribution("kilos_" + fname, kg, fleets=[fl], stocks=[fish],
                                                 function_f=g3l_distribution_sumofsquaredlogratios(), area_group=areas))
            continue
        actions.append(g3l_catchdistribution("ldist_" + fname, ld, fleets=[fl], stocks=[fish],
                                             function_f=g3l_distribution_sumofsquares(), area_group=None))
    actions += [g3a_report_detail(actions), g3l_bounds_penalty(actions)]
    model = g3_to_cuda(actions)
    p = model.parameter_template
    p = g3_init_val(p, "*.Linf", 60, spread=0.2)
    p = g3_init_val(p, "*.K", 0.3, lower=0.05, upper=1)
    p = g3_init_val(p, "*.t0", -0.5, optimise=False)
    p = g3_init_val(p, "*.lencv", 0.15, optimise=False)
    p = g3_init_val(p, "*.init.scalar", 10, optimise=False)
    p = g3_init_val(p, "*.init.#", 1, lower=0.1, upper=5)
    p = g3_init_val(p, "*.M.#", 0.2, lower=0.01, upper=1)
    p = g3_init_val(p, "init.F", 0.3, lower=0.1, upper=1)
    p = g3_init_val(p, "recage", 1, optimise=False)
    p = g3_init_val(p, "*.walpha", 1e-5, optimise=False)
    p = g3_init_val(p, "*.wbeta", 3, optimise=False)
    p = g3_init_val(p, "*.bbin", 10, lower=1, upper=100)
    p = g3_init_val(p, "*.rec.#", 5, lower=0.1, upper=20)
    p = g3_init_val(p, "*.rec.scalar", 10, optimise=False)
    p = g3_init_val(p, "*.rec.proj", 1)
    p = g3_init_val(p, "*.*.alpha", 0.2, lower=0.05, upper=0.5)
    p = g3_init_val(p, "*.*.l50", 25, lower=15, upper=35)
    p = g3_init_val(p, "*.*.catchability", 0.02, lower=0.005, upper=0.05)
    return model, p


LING_IMM_STDDEV = [8.25, 10.5644599516659, 12.4081745588022, 11.5741565728647, 11.0523508874244,
                   11.3447991170274, 11.7721342759715, 13.6152275606449]
LING_MAT_STDDEV = [12.4081745588022, 11.5741565728647, 11.0523508874244, 11.3447991170274, 11.7721342759715,
                   13.6152275606449, 14.8004893270652, 16.2753802766344, 17.9426701121357, 19.1787817582897,
                   15.9776436358384]


def ling_lln_table():
    """The reference's length-distribution fixture as column dicts (igfs_obs_data of inttest/codegeneration/run.R:120-122:
    read.table(...) plus area = 1), or None where the file is not shipped."""
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "catchdistribution_ldist_lln.txt")
    if not os.path.exists(path):
        return None
    rows = [ln.split() for ln in open(path).read().strip().split("\n")[1:]]
    return dict(year=[int(r[0]) for r in rows], step=[int(r[1]) for r in rows], area=["area1"] * len(rows),
                length=[float(r[2]) for r in rows], number=[float(r[3]) for r in rows])


def build_ling(seed=123):
    """The ling model of the reference's code-generation snapshot (inttest/codegeneration/run.R:17-183,
    generated C++ checked in as inttest/codegeneration/ling.cpp): same stocks, actions, formulas and
    parameter names, so that the oracle can be compared with that generated code (oracle/ref_lib.py).
    The observation table is the reference's own fixture, inttest/codegeneration/catchdistribution_ldist_lln.txt
    (2,270 rows of year, step, length, number; copied as data to tests/golden/, read as run.R:120-122 does); the numbers
    in ling_*_stddev and the parameter values are model DATA quoted from run.R:32-40,92-103,161-183."""
    from .formula import Lookup, ParamExpr, age as AGE, g3_as_integer, g3_exp
    rng = np.random.default_rng(seed)
    areas = g3_areas(["area1"])
    lg = range(20, 157, 4)
    ling_imm = g3s_livesonareas(g3s_age(g3_stock({"species": "ling", "maturity": "imm"}, lg), 3, 10), areas)
    ling_mat = g3s_livesonareas(g3s_age(g3_stock({"species": "ling", "maturity": "mat"}, lg), 5, 15), areas)
    P = lambda n, **kw: g3_parameterized(n, **kw)         # plain g3_param("name")

    def vec(name):      # g3_param_vector(name)[[as_integer(age) - minage + 1]]
        return ParamExpr(name, False, False, False, False, True, False, None, 0.0, float("nan"), float("nan"), False, ())

    def growth(wa, wb):
        return g3a_grow_impl_bbinom(
            g3a_grow_lengthvbsimple(linf_f=P("ling.Linf"), kappa_f=P("ling.K") * 0.001),
            g3a_grow_weightsimple(alpha_f=P(wa), beta_f=P(wb)),
            beta_f=P("ling.bbin") * 10, maxlengthgroupgrowth=15)

    vonb = lambda: g3a_renewal_vonb_recl(by_stock="species")
    rec_factor = P("ling.rec.scalar") * g3_parameterized("ling.rec", by_year=True)
    igfs = g3s_livesonareas(g3_fleet("igfs"), areas)
    years = list(range(1994, 2019))
    yy = np.repeat(years, 4)
    st = np.tile([1, 2, 3, 4], len(years))
    totaldata = dict(year=[int(y) for y in yy], step=[int(x) for x in st], area=["area1"] * len(yy),
                     total_weight=[float(x) for x in st])
    obs = ling_lln_table()
    if obs is None:         # (the fixture is missing: a table of the same shape)
        ns = 120
        yl, sl = np.repeat(yy, ns), np.repeat(st, ns)
        ln = np.clip(rng.normal(80, 25, len(yl)), 20, 400)
        obs = _count(dict(year=[int(y) for y in yl], step=[int(x) for x in sl], area=["area1"] * len(yl),
                          length=_cut(ln, list(range(20, 157, 4)))), ["year", "step", "area", "length"])
    actions = [
        g3a_time(1994, 2018, [3, 3, 3, 3]),
        # ling_mat_actions
        g3a_initialconditions_normalparam(
            ling_mat,
            factor_f=P("lingmat.init.scalar") * g3_exp(-1 * (P("lingmat.M") + P("ling.init.F")) * AGE) * vec("lingmat.init"),
            mean_f=vonb(), stddev_f=Lookup(LING_MAT_STDDEV, g3_as_integer(AGE), 5 - 1),
            alpha_f=P("lingmat.walpha"), beta_f=P("lingmat.wbeta")),
        g3a_growmature(ling_mat, growth("lingmat.walpha", "lingmat.wbeta")),
        g3a_naturalmortality(ling_mat, g3a_naturalmortality_exp(P("lingmat.M"))),
        g3a_age(ling_mat),
        # ling_imm_actions
        g3a_initialconditions_normalparam(
            ling_imm,
            factor_f=P("lingimm.init.scalar") * g3_exp(-1 * (P("lingimm.M") + P("ling.init.F")) * AGE) * vec("lingimm.init"),
            mean_f=vonb(), stddev_f=Lookup(LING_IMM_STDDEV, g3_as_integer(AGE), 3 - 1),
            alpha_f=P("lingimm.walpha"), beta_f=P("lingimm.wbeta")),
        g3a_renewal_normalparam(ling_imm, factor_f=rec_factor, mean_f=vonb(),
                                stddev_f=Lookup(LING_IMM_STDDEV, g3_as_integer(AGE), 3),
                                alpha_f=P("lingimm.walpha"), beta_f=P("lingimm.wbeta"), run_age=3, run_step=1),
        g3a_renewal_normalparam(ling_imm, factor_f=rec_factor, mean_f=vonb(),
                                stddev_f=Lookup(LING_IMM_STDDEV, g3_as_integer(AGE), 3),
                                alpha_f=P("lingimm.walpha"), beta_f=P("lingimm.wbeta"), run_age=5, run_step=1),
        g3a_growmature(ling_imm, growth("lingimm.walpha", "lingimm.wbeta"),
                       maturity_f=g3a_mature_constant(alpha=0.001 * P("ling.mat1"), l50=P("ling.mat2")),
                       output_stocks=[ling_mat]),
        g3a_naturalmortality(ling_imm, g3a_naturalmortality_exp(P("lingimm.M"))),
        g3a_age(ling_imm, output_stocks=[ling_mat]),
        # igfs_actions
        g3a_predate_fleet(igfs, [ling_imm, ling_mat],
                          suitabilities=g3_suitability_exponentiall50(P("ling.igfs.alpha"), P("ling.igfs.l50")),
                          catchability_f=g3a_predate_catchability_totalfleet(
                              g3_timeareadata("igfs_totaldata", totaldata, "total_weight", areas=areas))),
        # likelihood_actions
        g3l_understocking([ling_imm, ling_mat], nll_breakdown=True),
        g3l_catchdistribution("ldist_lln", obs, fleets=[igfs], stocks=[ling_imm, ling_mat],
                              function_f=g3l_distribution_sumofsquares(), area_group=areas, nll_breakdown=True),
    ]
    model = g3_to_cuda(actions)
    p = model.parameter_template
    for n, v in (("ling.Linf", 160), ("ling.K", 90), ("lingimm.walpha", 2.27567436711055e-06),
                 ("lingimm.wbeta", 3.20200445996187), ("ling.bbin", 6), ("lingimm.M", 0.15),
                 ("lingimm.init.scalar", 200), ("ling.init.F", 0.4), ("ling.recl", 12), ("ling.mat1", 70),
                 ("ling.mat2", 75), ("ling.rec.scalar", 400), ("lingmat.M", 0.15), ("lingmat.init.scalar", 200),
                 ("lingmat.walpha", 2.27567436711055e-06), ("lingmat.wbeta", 3.20200445996187),
                 ("ling.igfs.alpha", 0.5), ("ling.igfs.l50", 50)):
        p = g3_init_val(p, n, v, lower=0.5 * v, upper=1.5 * v)
    p = g3_init_val(p, "lingimm.init.#", 1, lower=0.5, upper=1.5)
    p = g3_init_val(p, "lingmat.init.#", 1, lower=0.5, upper=1.5)
    p = g3_init_val(p, "ling.rec.#", 1, lower=0.5, upper=1.5)
    p = g3_init_val(p, "recage", 0, optimise=False)
    return model, p


def profile_sweep(params, names, n_per_axis):
    """Likelihood-profile grid (BASELINE.json config 5): the template everywhere except `names`, which
    sweep their full [lower, upper] range on an n x n (x ...) grid. Returns B x n_optimised."""
    idx = params.optimised_indices()
    v = params.values()[idx]
    pos = [list(idx).index(params.index(n)) for n in names]
    axes = [np.linspace(params.lower[params.index(n)], params.upper[params.index(n)], n_per_axis) for n in names]
    grid = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, len(names))
    P = np.tile(v, (grid.shape[0], 1))
    P[:, pos] = grid
    return P


CONFIGS = {"C1": build_c1, "C1_MULTI": lambda: build_c1(likelihood="multinomial"), "C2": build_c2, "C2_MATAGE": lambda: build_c2(mat_age=True), "C2_STRAT": lambda: build_c2(stratified=True), "C4": build_c4, "C5": build_c5, "MINI_ANDERSEN": build_mini_andersen, "MINI_OTHERFOOD": lambda: build_mini_andersen(otherfood=True), "MINI_PREFS": lambda: build_mini_andersen(otherfood=True, prefs=True), "MINI_SUITS": lambda: build_mini_andersen(suits=True), "MINI_STOMACH": lambda: build_mini_andersen(stomach=True), "MINI_FLEETS": build_mini_fleets, "MINI_MIGRATE": build_mini_migrate, "MINI_TRANSFORM": build_mini_transform, "SPAWN_RIG": build_spawn_rig, "MINI_SPAWN": build_mini_spawn, "MINI_TIMEVAR": lambda: build_mini_spawn(timevar=True), "MINI_MATVAR": lambda: build_mini_spawn(matvar=True), "TIMEVAR_RIG": build_timevar_rig, "RANDOM_RIG": build_random_rig, "WLOSS_RIG": build_wloss_rig, "RENEWAL_RIG": build_renewal_rig, "SIX_FLEETS": build_six_fleets, "C1_RANDOM": lambda: build_c1(random=True), "C1_PROJ": lambda: build_c1(max_project_years=3),
           "LING": build_ling}
