"""g3_to_cuda(): lower an action list into the packed spec; g3_cuda_adfun(): bind it to the GPU.

Sits beside the reference's ``g3_to_r()`` / ``g3_to_tmb()`` (R/run_r.R:11, R/run_tmb.R:709) and
``g3_tmb_adfun()`` (R/run_tmb.R:1118): same inputs (action list, parameter template), same outputs
(``$par``, ``$fn``, ``$gr``, ``$report``) plus the batched entry points. Models outside the supported
action subset raise :class:`G3Unsupported`; there is no CPU fallback.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np

from . import actions as A
from . import likelihood as LK
from .formula import Ctx, G3Unsupported, Lowerer, ParameterTemplate, g3_parameterized, wrap
from .spec import K, SpecBuilder
from .stock import Stock


def _flatten(actions):
    out = []
    for a in actions:
        if a is None:
            continue
        if isinstance(a, (list, tuple)):
            out.extend(_flatten(a))
        else:
            out.append(a)
    return out


def _lgmatrix(src: Stock, dst: Stock):
    """stock_reshape() length-regrouping matrix (R/step.R:249-282); None when grids are equal."""
    s_lg, d_lg = list(src.minlen), list(dst.minlen)
    if len(s_lg) == len(d_lg) and np.allclose(s_lg, d_lg):
        return None
    s_up = src.upperlen if math.isfinite(src.upperlen) else s_lg[-1] + 1
    d_up = dst.upperlen if math.isfinite(dst.upperlen) else max(s_lg[-1], d_lg[-1]) + 1
    M = np.zeros((len(d_lg), len(s_lg)))
    for di in range(len(d_lg)):
        lo_d = d_lg[di]
        up_d = d_up if di >= len(d_lg) - 1 else d_lg[di + 1]
        for si in range(len(s_lg)):
            lo_s = s_lg[si]
            up_s = s_up if si >= len(s_lg) - 1 else s_lg[si + 1]
            inter = min(up_s, up_d) - max(lo_s, lo_d)
            M[di, si] = inter / (up_s - lo_s) if inter > 0 else 0.0
    return M


class G3CudaModel:
    """Result of g3_to_cuda(): carries the same attributes g3_to_tmb() attaches
    (``actions``, ``parameter_template``) plus the packed spec."""

    def __init__(self):
        self.actions = None
        self.parameter_template = None
        self._builder = None
        self.stock_names, self.stock_shapes, self.stock_dims = [], [], []
        self.comp_names, self.comp_shapes, self.comp_units = [], [], []
        self.random_names = []      # report names of the g3l_random_dnorm / g3l_random_walk terms, packed order
        self.pp_info = []           # (predator name, prey name, predator length groups or 1), packed pair order
        self.nll_names = []
        self.n_steps = 0
        self.time_labels = []
        self.has_bounds = None

    def pack(self, parameters: ParameterTemplate = None):
        """Final (ints, reals): adds the g3l_bounds_penalty section from the template's lower/upper,
        as g3_tmb_adfun fills the __lower/__upper DATA_SCALARs (R/run_tmb.R:1234-1235)."""
        pt = self.parameter_template if parameters is None else parameters
        if list(pt.switch) != list(self.parameter_template.switch):
            raise ValueError("Parameters not in expected order")   # R/run_tmb.R:1156-1158
        b = SpecBuilder()
        b.ints = list(self._builder.ints)
        b.reals = list(self._builder.reals)
        idx, lu = [], []
        if self.has_bounds is not None:
            for name in sorted(pt.switch):   # radix / C-locale order of step ids (R/run.R:67-69)
                i = pt.index(name)
                lo, up = pt.lower[i], pt.upper[i]
                if math.isfinite(lo) and math.isfinite(up):
                    idx.append(i)
                    lu += [lo, up]
            ws = [self.has_bounds.weight, self.has_bounds.scale]
        else:
            ws = [0.0, 0.0]
        b.ints[K["G3H_NBOUND"]] = len(idx)
        b.ints[K["G3H_BOUND_OFF"]] = b.alloc_ints(idx)
        b.ints[K["G3H_BOUND_ROFF"]] = b.alloc_reals(lu)
        b.ints[K["G3H_BOUND_WS_ROFF"]] = b.alloc_reals(ws)
        return b.finish()


def save_spec(model: "G3CudaModel", path, parameters: ParameterTemplate = None, report_len=None):
    """Write the packed spec for a consumer without this package -- the R glue (R/run_cuda.R: g3_cuda_load_spec()):
    <path>.ints (int32), <path>.reals (float64), <path>.json (sizes, the parameter template, the report layout) and,
    when ``report_len`` (g3cuda_report_len) is given, <path>.idx (int32): for every report array, in the model's own
    dimension order (first index fastest, as R stores arrays), the position of each element in g3cuda_report()'s vector."""
    import json
    pt = model.parameter_template if parameters is None else parameters
    ints, reals = model.pack(pt)
    np.asarray(ints, dtype=np.int32).tofile(str(path) + ".ints")
    np.asarray(reals, dtype=np.float64).tofile(str(path) + ".reals")
    layout = []
    if report_len:
        pieces = unpack_report(model, np.arange(int(report_len), dtype=np.float64))
        idx, pos = [], 0
        for name, arr in pieces.items():
            a = np.asarray(arr, dtype=np.float64)
            flat = a.reshape(-1, order="F") if a.ndim > 1 else a.reshape(-1)
            flat = np.nan_to_num(flat, nan=-1.0)
            layout.append({"name": name, "dim": list(a.shape) if a.ndim else [1], "idx_offset": pos})
            idx.append(flat.astype(np.int32)); pos += flat.size
        np.concatenate(idx).tofile(str(path) + ".idx")
    meta = {"n_ints": int(len(ints)), "n_reals": int(len(reals)), "report_layout": layout,
            "parameter_template": {"switch": list(pt.switch), "value": [float(v) for v in pt.values()],
                                   "optimise": [bool(i in set(pt.optimised_indices().tolist())) for i in range(len(pt.switch))],
                                   "lower": [float(v) for v in pt.lower], "upper": [float(v) for v in pt.upper]}}
    with open(str(path) + ".json", "w") as f:
        json.dump(meta, f)
    return meta


def g3_to_cuda(actions, strict=False, max_project_years=0) -> G3CudaModel:
    """``max_project_years``: how many years past end_year a parameter vector may run (its project_years, R/action_time.R:21-78).
    The reference grows its loop at run time; here every per-step table is laid out at lowering time, so the lowering
    needs the bound. In those years no observation has a time index and no landing is recorded (g3_timeareadata: missing =
    0), by-year parameter tables answer with their `ifmissing` (rec.proj) or NaN (R/run_tmb.R:1047-1054) -- as in the
    reference. Projection fleets / quotas / g3_param_project are outside the supported subset."""
    acts = _flatten(actions)
    by_kind = {}
    for a in acts:
        if not isinstance(a, A.Action):
            raise G3Unsupported(f"unsupported action object {a!r}")
        by_kind.setdefault(a.kind, []).append(a)

    times = by_kind.get("time", [])
    if len(times) != 1:
        raise ValueError("exactly one g3a_time() action is required")
    tm = times[0]
    S = len(tm.step_lengths)
    if S > 31:
        raise G3Unsupported("more than 31 steps per year")
    max_project_years = int(max_project_years)
    if max_project_years < 0:
        raise ValueError("max_project_years must be >= 0")
    last_year = tm.end_year + max_project_years        # every per-year / per-step table reaches this far
    n_years = last_year - tm.start_year + 1
    NT = S * n_years
    dt0 = tm.step_lengths[0] / 12.0

    # ---- stocks (alphabetical = step_id radix order, R/step.R:523-543) and fleets
    stocks = OrderedDict()
    preds = OrderedDict()

    def note_stock(s):
        if s.lengthgroups is None:
            raise G3Unsupported(f"{s.name} has no length dimension")
        if s.minage is None or s.areas is None:
            raise G3Unsupported(f"stock {s.name} must have age and area dimensions")
        stocks.setdefault(s.name, s)

    for k in ("init", "otherfood", "renewal", "naturalmortality", "growmature", "age", "migrate", "spawn"):
        for a in by_kind.get(k, []):
            note_stock(a.stock)
    for a in by_kind.get("growmature", []) + by_kind.get("age", []) + by_kind.get("spawn", []):
        for o in a.output_stocks:
            note_stock(o)
    for a in by_kind.get("predate", []):
        for p in a.prey_stocks:
            note_stock(p)
        preds.setdefault(a.predstock.name, a)
        if a.predstock.lengthgroups is not None:
            note_stock(a.predstock)
    for a in by_kind.get("distribution", []):
        for p in a.stocks:
            note_stock(p)
    for a in by_kind.get("understocking", []):
        for p in a.prey_stocks:
            note_stock(p)
    stocks = OrderedDict(sorted(stocks.items()))
    sidx = {n: i for i, n in enumerate(stocks)}
    preds = OrderedDict(sorted(preds.items()))

    model_areas = sorted({v for s in stocks.values() for v in s.areas.values()})

    b = SpecBuilder()
    H = b.ints
    tmpl = ParameterTemplate()
    low = Lowerer(b, tmpl, tm.start_year, tm.end_year, S)

    H[K["G3H_START_YEAR"]], H[K["G3H_END_YEAR"]] = tm.start_year, tm.end_year
    H[K["G3H_NSTEPS"]] = S
    H[K["G3H_STEPLEN_OFF"]] = b.alloc_ints(tm.step_lengths)
    H[K["G3H_NT"]] = NT
    H[K["G3H_NAREA"]] = len(model_areas)

    # parameter order follows first appearance in the collated code (R/run_tmb.R:715):
    H[K["G3H_PAR_RETRO"]] = tmpl.add("retro_years", 0.0, optimise=False)

    # ---- stock records
    cell_off = 0
    srec = {}
    for name, s in stocks.items():
        r = [0] * K["G3S_LEN"]
        r[K["G3S_L"]], r[K["G3S_MINAGE"]], r[K["G3S_NAGE"]] = s.L, s.minage, s.nage
        r[K["G3S_NAREA"]] = len(s.areas)
        r[K["G3S_AREA_OFF"]] = b.alloc_ints(s.areas.values())
        r[K["G3S_CELL_OFF"]] = cell_off
        r[K["G3S_MIDLEN_ROFF"]] = b.alloc_reals(s.midlen)
        r[K["G3S_PLUSDL_ROFF"]] = b.alloc_reals([s.plusdl])
        for f in ("G3S_INIT_OFF", "G3S_MORT_OFF", "G3S_GROW_N", "G3S_MIGRATE_OFF", "G3S_OTHERFOOD_OFF", "G3S_SPAWN_OFF", "G3S_MORT_TMAP_OFF",
                  "G3S_GROW_TMAP_OFF"):
            r[K[f]] = -1
        dims = [d for d in s.dims if d != "length"]
        r[K["G3S_DIMORDER"]] = 0 if dims == ["age", "area"] else 1
        s._cell_off = cell_off
        cell_off += s.L * s.nage * len(s.areas)
        srec[name] = r
    ncell = cell_off
    H[K["G3H_NCELL"]] = ncell

    def columns(s):
        for ri, (an, aid) in enumerate(s.areas.items()):
            for ai in range(s.nage):
                yield ri, an, aid, ai, s.minage + ai

    def one_per_stock(kind, label):
        out = {}
        for a in by_kind.get(kind, []):
            if a.stock.name in out:
                raise G3Unsupported(f"more than one {label} action for {a.stock.name}")
            out[a.stock.name] = a
        return out

    # ---- initial conditions (order -1)
    inits = one_per_stock("init", "g3a_initialconditions")
    for name, s in stocks.items():
        a = inits.get(name)
        if a is None:
            continue
        tab = []
        for ri, an, aid, ai, ag in columns(s):
            c = Ctx(stock=s, age=ag, area=aid, area_name=an, cur_year=tm.start_year, cur_step=1,
                    cur_step_size=dt0)
            tab += [low.slot(a.mean_f, c), low.slot(a.stddev_f, c), low.slot(a.factor_f, c),
                    low.slot(a.alpha_f, c), low.slot(a.beta_f, c)]
        srec[name][K["G3S_INIT_OFF"]] = b.alloc_ints(tab)

    # ---- g3a_otherfood (order -1, after the initial conditions in step-id order): (num, wgt) slots per step and area
    for name, a in one_per_stock("otherfood", "g3a_otherfood").items():
        s = stocks[name]
        if inits.get(name) is not None or any(x.stock.name == name for k in ("renewal", "growmature", "age", "migrate", "naturalmortality")
                                               for x in by_kind.get(k, [])):
            raise G3Unsupported(f"g3a_otherfood({name}): an otherfood stock has no other actions of its own")
        tab = []
        for t in range(NT):
            for an, aid in s.areas.items():
                c = Ctx(stock=s, age=s.minage, area=aid, area_name=an, cur_year=tm.start_year + t // S, cur_step=t % S + 1,
                        cur_step_size=tm.step_lengths[t % S] / 12.0)
                tab += [low.slot(a.num_f, c), low.slot(a.wgt_f, c)]
        srec[name][K["G3S_OTHERFOOD_OFF"]] = b.alloc_ints(tab)

    if by_kind.get("report_detail"):
        tmpl.add("report_detail", 1.0, optimise=False)

    # ---- migration (order 2). Several g3a_migrate actions on one stock are allowed when they run on different steps
    # (tests/test-action_migrate.R:41-52: a spring and a winter migration): the table is per step.
    mig_tabs = {}
    for a in by_kind.get("migrate", []):
        s = stocks[a.stock.name]
        anames = list(s.areas.keys())
        NR = len(anames)
        tab, taken = mig_tabs.setdefault(s.name, ([-1] * (S * NR * NR), set()))
        run_steps = list(range(1, S + 1)) if a.run_steps is None else [int(x) for x in a.run_steps]
        if taken & set(run_steps):
            raise G3Unsupported(f"more than one g3a_migrate action for {s.name} on step(s) {sorted(taken & set(run_steps))}")
        taken |= set(run_steps)
        for st in run_steps:
            for di, dn in enumerate(anames):
                for si, sn in enumerate(anames):
                    if di == si:
                        continue        # the stationary share balances the row (R/action_migrate.R:2-15)
                    e = a.migrate_f(sn, dn)
                    if isinstance(e, (int, float)) and e == 0:
                        continue
                    c = Ctx(stock=s, area=s.areas[sn], area_name=sn, cur_step=st)
                    tab[((st - 1) * NR + di) * NR + si] = low.slot(wrap(e), c)
    for name, (tab, _) in mig_tabs.items():
        srec[name][K["G3S_MIGRATE_OFF"]] = b.alloc_ints(tab)

    def time_sets(slots_at):
        """Time-varying formulas (by_year / by_step tables, g3_timevariable -- R/params.R:122-129, R/timevariable.R:1-31): the
        slots of every step, deduplicated into sets. Returns (flat list of the sets' slots, per-step set index or None when one
        set serves every step)."""
        sets, index, tmap = [], {}, []
        for t in range(NT):
            key = tuple(slots_at(dict(cur_year=tm.start_year + t // S, cur_step=t % S + 1,
                                      cur_step_size=tm.step_lengths[t % S] / 12.0)))
            if key not in index:
                index[key] = len(sets)
                sets.append(key)
            tmap.append(index[key])
        return [x for k in sets for x in k], (tmap if len(sets) > 1 else None)

    # ---- predation (order 3): predators by name, prey by name
    pred_recs, pp_recs, pp_index, pp_info = [], [], {}, []
    for pname, a in preds.items():
        pr = [0] * K["G3P_LEN"]
        ps = a.predstock
        if ps.areas is None:
            raise G3Unsupported(f"predator {pname} must live on areas")
        cat = a.catchability_f
        if cat.kind in ("totalfleet", "numberfleet", "linearfleet", "effortfleet"):
            if ps.lengthgroups is not None:
                raise G3Unsupported(f"{cat.kind} catchability on a stock predator")
            pr[K["G3P_KIND"]] = K["G3PRED_" + cat.kind.upper()]
            E = []
            for t in range(NT):
                y, st = tm.start_year + t // S, t % S + 1
                for aid in ps.areas.values():
                    E.append(cat.E.lookup(y, st, aid))
            pr[K["G3P_E_ROFF"]] = b.alloc_reals(E)
            pr[K["G3P_STOCK"]] = -1
            pr[K["G3P_SLOT_OFF"]] = -1
        elif cat.kind == "predator":
            # g3a_predate_catchability_predator (R/action_predate.R:207-238) for a stock predator
            if ps.lengthgroups is None or ps.name not in sidx:
                raise G3Unsupported("g3a_predate_catchability_predator needs a stock predator with length groups")
            pr[K["G3P_KIND"]] = K["G3PRED_STOCK"]
            pr[K["G3P_E_ROFF"]] = -1
            pr[K["G3P_STOCK"]] = sidx[ps.name]
            mc = cat.kw["max_consumption"]
            cp = Ctx(predstock=ps)
            pr[K["G3P_SLOT_OFF"]] = b.alloc_ints([low.slot(mc["m0"], cp), low.slot(mc["m1"], cp), low.slot(mc["m2"], cp),
                                                  low.slot(mc["m3"], cp), low.slot(cat.kw["half_feeding_f"], cp),
                                                  low.slot(mc["temperature"], cp)])
        else:
            raise G3Unsupported(f"catchability {cat.kind}")
        pr[K["G3P_NAREA"]] = len(ps.areas)
        pr[K["G3P_AREA_OFF"]] = b.alloc_ints(ps.areas.values())
        pr[K["G3P_PP_FIRST"]] = len(pp_recs)
        for prey in sorted(a.prey_stocks, key=lambda x: x.name):
            su = a.suit_for(prey)
            c = Ctx(stock=prey, predstock=ps)
            q = [0] * K["G3PP_LEN"]
            q[K["G3PP_PRED"]], q[K["G3PP_PREY"]] = len(pred_recs), sidx[prey.name]
            q[K["G3PP_SUIT_KIND"]] = K["G3SUIT_" + {"expl50": "EXPL50"}.get(su.kind, su.kind.upper())]
            slots, stmap = time_sets(lambda tc: (lambda sl: sl + [-1] * (6 - len(sl)))([low.slot(x, c.replace(**tc)) for x in su.args]))
            q[K["G3PP_SUIT_OFF"]] = b.alloc_ints(slots)
            q[K["G3PP_SUIT_TMAP"]] = -1 if stmap is None else b.alloc_ints(stmap)
            q[K["G3PP_ENERGY"]] = q[K["G3PP_PREF"]] = -1
            if cat.kind == "predator":
                q[K["G3PP_ENERGY"]] = low.slot(cat.kw["energycontent"], c)
                # prey_preferences: one value / formula, or a list by prey stock name (list_to_stock_switch,
                # R/action_predate.R:208,225); the exponent 1 is not lowered at all
                pref = cat.kw["prey_preferences"]
                if isinstance(pref, dict):
                    if prey.name not in pref:
                        raise G3Unsupported(f"g3a_predate_catchability_predator: no prey_preferences entry for {prey.name}")
                    pref = pref[prey.name]
                if not (isinstance(pref, (int, float)) and pref == 1):
                    q[K["G3PP_PREF"]] = low.slot(wrap(pref), c)
            elif cat.kind == "effortfleet":
                cf = cat.kw["catchability_fs"]
                q[K["G3PP_ENERGY"]] = low.slot(wrap(cf[prey.name] if isinstance(cf, dict) else cf), c)
            if cat.kind != "predator" and su.kind in ("andersen", "exponential", "richards"):
                raise G3Unsupported(f"g3_suitability_{su.kind} needs a stock predator (predator_length)")
            pp_index[(pname, prey.name)] = len(pp_recs)
            pp_recs.append(q)
            pp_info.append((pname, prey.name, ps.L if cat.kind == "predator" else 1))
        pr[K["G3P_N_PP"]] = len(pp_recs) - pr[K["G3P_PP_FIRST"]]
        pred_recs.append(pr)

    # ---- natural mortality (order 4)
    for name, a in one_per_stock("naturalmortality", "g3a_naturalmortality").items():
        s = stocks[name]
        tab, tmap = time_sets(lambda tc: [low.slot(a.param_f, Ctx(stock=s, age=ag, area=aid, area_name=an, **tc))
                                          for ri, an, aid, ai, ag in columns(s)])
        srec[name][K["G3S_MORT_OFF"]] = b.alloc_ints(tab)
        if tmap is not None:
            srec[name][K["G3S_MORT_TMAP_OFF"]] = b.alloc_ints(tmap)

    # ---- growth + maturity (order 5 / 7)
    maxn = 0

    def out_table(s, outs, ratios):
        tab = []
        for o, rt in zip(outs, ratios):
            if not np.allclose(o.minlen, s.minlen) or o.L != s.L:
                raise G3Unsupported(f"moving {s.name} into {o.name}: length groups differ")
            tab += [sidx[o.name], low.slot(wrap(rt), Ctx(stock=s))]
        return b.alloc_ints(tab)

    for name, a in one_per_stock("growmature", "g3a_growmature").items():
        s, r, im = stocks[name], srec[name], a.impl_f
        c = Ctx(stock=s)
        r[K["G3S_GROW_N"]] = im.maxlengthgroupgrowth
        maxn = max(maxn, im.maxlengthgroupgrowth)
        m = a.maturity_f
        mat_fs = () if m is None else (m.alpha, m.l50, m.beta, m.a50)
        # the growth parameters and, with them, the maturity parameters of every step, deduplicated into sets: the maturing
        # share is part of the growth bands (R/action_grow.R:420-437), so the two vary together
        gmtab, gtmap = time_sets(lambda tc: [low.slot(f, c.replace(**tc)) for f in (
            im.delta_len_f.linf_f, im.delta_len_f.kappa_f, im.beta_f, im.delta_wgt_f.alpha_f, im.delta_wgt_f.beta_f)]
            + [(-1 if x is None else low.slot(x, c.replace(**tc))) for x in mat_fs])
        w = 5 + len(mat_fs)
        r[K["G3S_GROW_OFF"]] = b.alloc_ints([x for i in range(0, len(gmtab), w) for x in gmtab[i:i + 5]])
        if gtmap is not None:
            r[K["G3S_GROW_TMAP_OFF"]] = b.alloc_ints(gtmap)
        if m is not None:
            r[K["G3S_MAT_KIND"]] = K["G3MAT_CONTINUOUS"] if m.kind == "continuous" else K["G3MAT_CONSTANT"]
            mtab = [x for i in range(0, len(gmtab), w) for x in gmtab[i + 5:i + 9]]
            if m.kind != "continuous" and any(mtab[i:i + 4] != mtab[:4] for i in range(0, len(mtab), 4)):
                raise G3Unsupported("g3a_mature_constant with time-varying parameters")
            r[K["G3S_MAT_OFF"]] = b.alloc_ints(mtab)
            if m.kind == "continuous" and (m.alpha is None or m.l50 is None):
                raise G3Unsupported("g3a_mature_continuous needs alpha and l50")
            if (m.alpha is None) != (m.l50 is None) or (m.beta is None) != (m.a50 is None):
                raise ValueError("l50 must be supplied if alpha is supplied (and a50 with beta)")
            tf = a.transition_f
            if tf == "cur_step_final":
                mask = 1 << (S - 1)
            elif tf is True:
                mask = (1 << S) - 1
            else:
                mask = 0
                for st in tf:
                    mask |= 1 << (int(st) - 1)
            r[K["G3S_TRANS_MASK"]] = mask
            r[K["G3S_N_OUT"]] = len(a.output_stocks)
            r[K["G3S_OUT_OFF"]] = out_table(s, a.output_stocks, a.output_ratios)
            omap = []
            for o in a.output_stocks:
                for ri, an, aid, ai, ag in columns(s):
                    ok = (o.minage <= ag <= o.maxage) and (aid in o.areas.values())
                    if ok:
                        ro = list(o.areas.values()).index(aid)
                        base = o._cell_off + (ro * o.nage + (ag - o.minage)) * o.L
                    for l in range(s.L):
                        omap.append(base + l if ok else -1)
            r[K["G3S_OUT_MAP_OFF"]] = b.alloc_ints(omap)

    # ---- spawning (order 6: proportion, recruitment, parent mortality)
    def suit_slots(f, c, what):
        """A formula (-> G3SUIT_CONSTANT) or a g3_suitability_*() of the stock's own length."""
        if hasattr(f, "kind") and hasattr(f, "args"):
            if f.kind in ("andersen", "exponential", "richards"):
                raise G3Unsupported(f"{what}: g3_suitability_{f.kind} reads predator_length")
            kind = K["G3SUIT_" + {"expl50": "EXPL50"}.get(f.kind, f.kind.upper())]
            sl = [low.slot(x, c) for x in f.args]
        else:
            kind, sl = K["G3SUIT_CONSTANT"], [low.slot(wrap(f), c)]
        return [kind] + sl + [-1] * (5 - len(sl))

    spawn_recs = {}
    for name, a in one_per_stock("spawn", "g3a_spawn").items():
        s = stocks[name]
        # by_predator names read "spawn" (proportion_f is substituted with predstock = 'spawn', R/action_spawn.R:165)
        c = Ctx(stock=s, predstock=type("_SpawnName", (), {"name": "spawn", "name_parts": {}})())
        rf = a.recruitment_f
        q = [0] * K["G3SP_LEN"]
        q[K["G3SP_RUN_STEP"]] = 0 if a.run_step is None else int(a.run_step)
        q[K["G3SP_RECR_KIND"]] = K["G3SPAWN_" + rf.kind.upper()]
        q[K["G3SP_PROP_KIND"]:K["G3SP_PROP_KIND"] + 6] = suit_slots(a.proportion_f, c, "g3a_spawn(proportion_f)")
        rs = [low.slot(x, c) for x in rf.s_args] + [low.slot(x, c) for x in rf.r_args]
        if rf.kind == "fecundity":
            rs = rs[4:] + rs[:4]        # stored p0..p4; declared in the order the reference's code reads them (s, then r)
        if rf.R_by_year is not None:
            rs.append(b.alloc_ints([low.slot(rf.R_by_year, Ctx(stock=s, cur_year=y)) for y in range(tm.start_year, last_year + 1)]))
        q[K["G3SP_RECR_SLOTS"]:K["G3SP_RECR_SLOTS"] + 5] = rs + [-1] * (5 - len(rs))
        if isinstance(a.mortality_f, (int, float)) and a.mortality_f == 0:
            q[K["G3SP_MORT_KIND"]:K["G3SP_MORT_KIND"] + 6] = [-1] * 6
        else:
            q[K["G3SP_MORT_KIND"]:K["G3SP_MORT_KIND"] + 6] = suit_slots(a.mortality_f, c, "g3a_spawn(mortality_f)")
        wl = getattr(a, "weightloss", None)
        q[K["G3SP_WLOSS"]:K["G3SP_WLOSS"] + 2] = [-1, -1] if wl is None else [low.slot(wl[0], c), low.slot(wl[1], c)]
        spawn_recs[name] = q

    # ---- renewal (order 8)
    for name, s in stocks.items():
        recs = []
        for a in by_kind.get("renewal", []):
            if a.stock.name != name:
                continue
            if a.run_age is None:
                raise G3Unsupported("g3a_renewal: run_age = NULL")
            if max_project_years > 0 and not a.run_projection:
                raise G3Unsupported("g3a_renewal(run_projection = FALSE) in a model lowered with max_project_years > 0")
            if not (s.minage <= a.run_age <= s.maxage):
                continue
            an0, aid0 = next(iter(s.areas.items()))
            c = Ctx(stock=s, age=a.run_age, area=aid0, area_name=an0)
            # Shape parameters that vary by year (mean length, sd, walpha, wbeta as by_year tables or g3_timevariable): one
            # record per set of years that share their values, each with the factor slots of its own years only -- in any
            # year one record of the family runs, the others skip (FACTOR_OFF = -1), so the kernels need nothing new.
            run_steps = [int(a.run_step)] if a.run_step is not None else list(range(1, S + 1))
            families = OrderedDict()
            for y in range(tm.start_year, last_year + 1):
                keys = {tuple(low.slot(f, c.replace(cur_year=y, cur_step=st, cur_step_size=tm.step_lengths[st - 1] / 12.0))
                              for f in (a.mean_f, a.stddev_f, a.alpha_f, a.beta_f)) for st in run_steps}
                if len(keys) > 1:
                    raise G3Unsupported(f"g3a_renewal of {name}: shape parameters that change between the steps of a year")
                families.setdefault(keys.pop(), set()).add(y)
            for shape, fam_years in families.items():
                q = [0] * K["G3R_LEN"]
                q[K["G3R_RUN_STEP"]] = 0 if a.run_step is None else int(a.run_step)
                q[K["G3R_AGE_IDX"]] = a.run_age - s.minage
                q[K["G3R_MEAN"]], q[K["G3R_SD"]], q[K["G3R_WALPHA"]], q[K["G3R_WBETA"]] = shape
                ftab = []
                for y in range(tm.start_year, last_year + 1):
                    for an, aid in s.areas.items():
                        cy = Ctx(stock=s, age=a.run_age, area=aid, area_name=an, cur_year=y,
                                 cur_step=(a.run_step or 1))
                        fslot = low.slot(a.factor_f, cy)         # (lowered for every year: the parameter order does not depend on the split)
                        ftab.append(fslot if y in fam_years else -1)
                q[K["G3R_FACTOR_OFF"]] = b.alloc_ints(ftab)
                recs.append(q)
        srec[name][K["G3S_N_RENEW"]] = len(recs)
        srec[name][K["G3S_RENEW_OFF"]] = b.alloc_ints([x for q in recs for x in q])

    # ---- spawning, the offspring (order 8, after the renewals in step-id order): formulas in the OUTPUT stock's context
    for name, q in spawn_recs.items():
        a = [x for x in by_kind["spawn"] if x.stock.name == name][0]
        tab = []
        for o, rt in zip(a.output_stocks, a.output_ratios):
            co = Ctx(stock=o, age=o.minage)
            tab += [sidx[o.name], low.slot(a.mean_f, co), low.slot(a.stddev_f, co)]
            ratio = low.slot(wrap(rt), co)
            tab += [low.slot(a.alpha_f, co), low.slot(a.beta_f, co), ratio]
        q[K["G3SP_N_OUT"]], q[K["G3SP_OUT_OFF"]] = len(a.output_stocks), b.alloc_ints(tab)
        srec[name][K["G3S_SPAWN_OFF"]] = b.alloc_ints(q)

    # ---- ageing (order 12)
    for name, a in one_per_stock("age", "g3a_age").items():
        s, r = stocks[name], srec[name]
        r[K["G3S_AGE_ENABLE"]] = 1
        r[K["G3S_AGE_N_OUT"]] = len(a.output_stocks)
        r[K["G3S_AGE_OUT_OFF"]] = out_table(s, a.output_stocks, a.output_ratios)
        amap = []
        for o in a.output_stocks:
            ag = s.maxage + 1
            for an, aid in s.areas.items():
                ok = (o.minage <= ag <= o.maxage) and (aid in o.areas.values())
                if ok:
                    ro = list(o.areas.values()).index(aid)
                    base = o._cell_off + (ro * o.nage + (ag - o.minage)) * o.L
                for l in range(s.L):
                    amap.append(base + l if ok else -1)
        r[K["G3S_AGE_MAP_OFF"]] = b.alloc_ints(amap)

    # ---- likelihood components (order 10), by generated name
    comps = sorted(by_kind.get("distribution", []), key=lambda a: a.nll_name)
    xbufs, xbuf_index, comp_recs = [], {}, []
    model = G3CudaModel()
    for a in comps:
        ld = a.ld
        wslot_ctx = Ctx()
        wcode = low._lower(a.weight, wslot_ctx)
        if wcode[0] != "x" or len(wcode[1]) != 2 or wcode[1][0] != K["G3OP_PARAM"]:
            raise G3Unsupported("likelihood weight must be a plain parameter")
        q = [0] * K["G3C_LEN"]
        fn = a.function_f
        q[K["G3C_FUNC"]] = {"sumofsquares": K["G3F_SUMOFSQUARES"], "multinomial": K["G3F_MULTINOMIAL"],
                            "surveyindices_log": K["G3F_SI_LOG"],
                            "surveyindices_linear": K["G3F_SI_LINEAR"],
                            "sumofsquaredlogratios": K["G3F_SUMSQLOGRATIOS"]}[fn.name]
        q[K["G3C_WEIGHT_PAR"]] = wcode[1][1]
        st_sorted = sorted(a.stocks, key=lambda x: x.name)
        if a.fleets:
            pps = []
            for f in sorted(a.fleets, key=lambda x: x.name):
                for p in st_sorted:
                    if (f.name, p.name) not in pp_index:
                        raise G3Unsupported(f"{a.nll_name}: no g3a_predate for {f.name} on {p.name}")
                    pps.append(pp_index[(f.name, p.name)])
            key = ("catch", ld.unit, tuple(pps))
        else:
            pps = []
            key = ("abund", ld.unit, tuple(sidx[p.name] for p in st_sorted))
        # destination of every cell
        entries = [[] for _ in range(ld.n_dest)]
        multi_dest = False          # an age transform sends one cell to several destinations: CSR mode whatever n_bins
        cell_dest = [-1] * ncell
        cell_coef = [0.0] * ncell
        for p in st_sorted:
            sg = 0
            if ld.stock_map is not None:
                sg = ld.stock_map[p.name]
                if sg < 0:
                    continue
            M = _lgmatrix(p, ld.lenstock)
            # transform_fs (R/likelihood_distribution.R:298-347), constant matrices folded into the gather: the length
            # matrix acts on the collected vector before the reshape (lgmatrix %*% (T %*% x)); what is collected at age
            # `preage` lands in every age `a` with the factor Tage[preage_idx, a_idx]
            Tl, Ta = a.transform_fs[p.name]["length"], a.transform_fs[p.name]["age"]
            if Tl is not None:
                M = (np.eye(p.L) if M is None else np.asarray(M, dtype=float)) @ Tl
            for ri, an, aid, ai, ag in columns(p):
                if ld.area_group is not None:
                    hit = [i for i, v in enumerate(ld.area_group.values()) if int(v) == aid]
                    if not hit:
                        continue
                    areag = hit[0]
                else:
                    areag = 0
                # (destination age group, factor) of what this column holds
                if Ta is None:
                    dests = [(ag, 1.0)]
                else:
                    multi_dest = True
                    dests = [(p.minage + k, float(Ta[ai, k])) for k in range(p.nage) if Ta[ai, k] != 0.0]
                for dag, af in dests:
                    if ld.minage is not None:
                        if not (ld.minage <= dag <= ld.maxage):
                            continue
                        ageg = dag - ld.minage
                    else:
                        ageg = 0
                    for l in range(p.L):
                        cell = p._cell_off + (ri * p.nage + ai) * p.L + l
                        for bi in range(ld.n_bins):
                            coef = ((1.0 if bi == l else 0.0) if M is None else M[bi, l]) * af
                            if coef != 0.0:
                                e = ld.dest(areag, sg, ageg, bi)
                                entries[e].append((cell, coef))
                                cell_dest[cell], cell_coef[cell] = e, coef
        mode = 1 if (ld.n_bins == 1 and not multi_dest) else 0
        q[K["G3C_MODE"]] = mode
        q[K["G3C_N_DEST"]], q[K["G3C_N_BINS"]] = ld.n_dest, ld.n_bins
        q[K["G3C_N_SLICES"]], q[K["G3C_N_AREAG"]] = ld.n_slices, ld.n_areag
        if mode == 0:
            ptr, idx, coef = [0], [], []
            for e in entries:
                for cell, cf in e:
                    idx.append(cell)
                    coef.append(cf)
                ptr.append(len(idx))
            q[K["G3C_CSR_PTR_OFF"]] = b.alloc_ints(ptr)
            q[K["G3C_CSR_IDX_OFF"]] = b.alloc_ints(idx)
            q[K["G3C_CSR_COEF_ROFF"]] = b.alloc_reals(coef)
            q[K["G3C_CELL_DEST_OFF"]] = q[K["G3C_CELL_COEF_ROFF"]] = -1
        else:
            q[K["G3C_CELL_DEST_OFF"]] = b.alloc_ints(cell_dest)
            q[K["G3C_CELL_COEF_ROFF"]] = b.alloc_reals(cell_coef)
            q[K["G3C_CSR_PTR_OFF"]] = q[K["G3C_CSR_IDX_OFF"]] = q[K["G3C_CSR_COEF_ROFF"]] = -1
        tkey = {t: i for i, t in enumerate(ld.times)}
        tidx, cmp_ = [], []
        for t in range(NT):
            y, st = tm.start_year + t // S, t % S + 1
            tidx.append(tkey.get(y * 100 + (st if ld.has_step else 0), -1))
            cmp_.append(1 if (ld.has_step or st == S) else 0)
        q[K["G3C_TIDX_OFF"]], q[K["G3C_CMP_OFF"]] = b.alloc_ints(tidx), b.alloc_ints(cmp_)
        q[K["G3C_N_TIMES"]] = len(ld.times)
        q[K["G3C_OBS_ROFF"]] = b.alloc_reals(ld.obs)
        if fn.name.startswith("surveyindices"):
            al, be = fn.kw["alpha"], fn.kw["beta"]
            q[K["G3C_PARAM_ROFF"]] = b.alloc_reals([float("nan") if al is None else float(al),
                                                    float("nan") if be is None else float(be)])
        elif fn.name in ("multinomial", "sumofsquaredlogratios"):
            q[K["G3C_PARAM_ROFF"]] = b.alloc_reals([fn.kw["epsilon"], 0.0])
        else:       # sumofsquares: PARAM[0] = 1 for the stratified form (proportions within each length group)
            q[K["G3C_PARAM_ROFF"]] = b.alloc_reals([1.0 if fn.kw.get("stratified") else 0.0, 0.0])
        if key not in xbuf_index:
            xbuf_index[key] = len(xbufs)
            xbufs.append(dict(kind=0 if key[0] == "catch" else 1, unit=0 if ld.unit == "num" else 1,
                              pps=list(pps), active=[0] * NT))
        xb = xbufs[xbuf_index[key]]
        q[K["G3C_XBUF"]] = xbuf_index[key]
        for t in range(NT):
            if tidx[t] >= 0:
                xb["active"][t] = 1
        comp_recs.append(q)
        model.comp_names.append(a.nll_name)
        model.comp_units.append(ld.unit)
        model.comp_shapes.append(dict(n_times=len(ld.times), n_areag=ld.n_areag, n_stockg=ld.n_stockg,
                                      n_age=ld.n_age, n_bins=ld.n_bins, times=list(ld.times)))

    # ---- random-effect penalties (order 10, after the distributions by step id): g3l_random_dnorm / g3l_random_walk
    # (R/likelihood_random.R:14-109). Everything they read is a formula of parameters: per step, the slots of param_f,
    # mean_f (walk: param_f of the period's previous step) and sigma_f.
    rand_recs = []
    model.random_names = []
    rands = sorted(by_kind.get("random", []), key=lambda a: ("g3l_random_walk" if a.walk else "g3l_random_dnorm", a.nll_name))
    for a in rands:
        run, ps, ms, ss = [0] * NT, [0] * NT, [0] * NT, [0] * NT
        prev = -1
        for t in range(NT):
            c = Ctx(cur_year=tm.start_year + t // S, cur_step=t % S + 1, cur_step_size=tm.step_lengths[t % S] / 12.0)
            ps[t], ss[t] = low.slot(a.param_f, c), low.slot(a.sigma_f, c)
            # 'single' scores once, at the last step of the RUN (cur_time == total_steps, R/likelihood_random.R:46,97): that
            # step moves with retro_years, so the flag says "the run's last step" (2) and the kernels compare with total_steps
            on = {"step": True, "year": t % S == S - 1, "single": True}[a.period]
            if a.walk:
                ms[t] = prev if prev >= 0 else ps[t]
                run[t] = 1 if (on and prev >= 0 and a.period != "single") else 0       # (a 'single' walk has no previous value: never scores)
                if on:
                    prev = ps[t]
            else:
                ms[t] = low.slot(a.mean_f, c)
                run[t] = (2 if a.period == "single" else 1) if on else 0
        q = [0] * K["G3RN_LEN"]
        q[K["G3RN_RUN_OFF"]], q[K["G3RN_PARAM_OFF"]] = b.alloc_ints(run), b.alloc_ints(ps)
        q[K["G3RN_MEAN_OFF"]], q[K["G3RN_SIGMA_OFF"]] = b.alloc_ints(ms), b.alloc_ints(ss)
        q[K["G3RN_WEIGHT"]] = low.slot(a.weight, Ctx())
        q[K["G3RN_LOG"]] = 1 if a.log_f else 0
        rand_recs.append(q)
        model.random_names.append(a.report_name)

    xrecs = []
    for xb in xbufs:
        x = [0] * K["G3X_LEN"]
        x[K["G3X_KIND"]], x[K["G3X_UNIT"]] = xb["kind"], xb["unit"]
        x[K["G3X_N_PP"]], x[K["G3X_PP_OFF"]] = len(xb["pps"]), b.alloc_ints(xb["pps"])
        x[K["G3X_ACTIVE_OFF"]] = b.alloc_ints(xb["active"])
        xrecs.append(x)

    # ---- understocking
    us = by_kind.get("understocking", [])
    if len(us) > 1:
        raise G3Unsupported("more than one g3l_understocking")
    if us:
        H[K["G3H_US_N"]] = len(us[0].prey_stocks)
        H[K["G3H_US_OFF"]] = b.alloc_ints(sorted(sidx[p.name] for p in us[0].prey_stocks))
        H[K["G3H_US_ROFF"]] = b.alloc_reals([us[0].power, us[0].weight])
    else:
        H[K["G3H_US_N"]], H[K["G3H_US_OFF"]] = 0, 0
        H[K["G3H_US_ROFF"]] = b.alloc_reals([2.0, 0.0])

    H[K["G3H_PAR_PROJECT"]] = tmpl.add("project_years", 0.0, optimise=False)

    # ---- write record tables
    H[K["G3H_NSTOCK"]] = len(stocks)
    H[K["G3H_STOCK_OFF"]] = b.alloc_ints([x for n in stocks for x in srec[n]])
    H[K["G3H_NPRED"]] = len(pred_recs)
    H[K["G3H_PRED_OFF"]] = b.alloc_ints([x for r in pred_recs for x in r])
    H[K["G3H_NPP"]] = len(pp_recs)
    H[K["G3H_PP_OFF"]] = b.alloc_ints([x for r in pp_recs for x in r])
    H[K["G3H_NCOMP"]] = len(comp_recs)
    H[K["G3H_COMP_OFF"]] = b.alloc_ints([x for r in comp_recs for x in r])
    H[K["G3H_NXBUF"]] = len(xrecs)
    H[K["G3H_XBUF_OFF"]] = b.alloc_ints([x for r in xrecs for x in r])
    H[K["G3H_NSLOT"]], H[K["G3H_SLOT_OFF"]] = low.finish()
    H[K["G3H_NPAR"]] = len(tmpl)
    H[K["G3H_MAXL"]] = max(s.L for s in stocks.values())
    H[K["G3H_MAXN"]] = maxn
    H[K["G3H_NRAND"]] = len(rand_recs)
    H[K["G3H_RAND_OFF"]] = b.alloc_ints([x for r in rand_recs for x in r])
    H[K["G3H_NNLL"]] = len(comp_recs) + len(rand_recs) + 2

    model.actions = acts
    model.parameter_template = tmpl
    model._builder = b
    bp = by_kind.get("bounds_penalty", [])
    model.has_bounds = bp[0] if bp else None
    for n, s in stocks.items():
        model.stock_names.append(n)
        model.stock_shapes.append((s.L, s.nage, len(s.areas)))
        model.stock_dims.append(list(s.dims))
    model.pp_info = pp_info
    model.nll_names = list(model.comp_names) + list(model.random_names) + ["understocking", "bounds"]
    model.n_steps = NT
    model.time_labels = ["%d-%02d" % (tm.start_year + t // S, t % S + 1) for t in range(NT)]
    return model


# --------------------------------------------------------------------------- binding
class G3CudaADFun:
    """Mirror of the list g3_tmb_adfun() returns (R/run_tmb.R:1254-1325): ``par``, ``fn``, ``gr``,
    ``report`` -- plus batched ``fn_batch`` / ``gr_batch`` / ``nll_components_batch``."""

    def __init__(self, model: G3CudaModel, parameters: ParameterTemplate = None, device=0, devices=None):
        from ._lib import Handle, MultiHandle
        self.model = model
        self.parameters = (model.parameter_template if parameters is None else parameters).copy()
        ints, reals = model.pack(self.parameters)
        self.opt_idx = self.parameters.optimised_indices()
        if len(self.opt_idx) == 0:
            raise ValueError("No parameters with optimise=TRUE set")    # R/run_tmb.R:1166-1168
        self.par_names = [self.parameters.switch[i] for i in self.opt_idx]
        self.par = self.parameters.values()[self.opt_idx].copy()
        self._full = self.parameters.values()
        # devices = [0, 1, ...]: one handle over several GPUs of the box, the rows of every batch cut into one block per
        # device (g3cuda_multi_*); `handle` is then the first device's (report, single-vector calls, tuning hooks)
        self.multi = MultiHandle(ints, reals, devices) if devices is not None and len(devices) > 1 else None
        self.handle = self.multi.handles[0] if self.multi else Handle(ints, reals, devices[0] if devices else device)
        self._batch = self.multi if self.multi else self.handle
        self.last_par = self.par.copy()

    # full PARAMETER matrix from optimiser-facing vectors
    def full_par(self, P) -> np.ndarray:
        P = np.atleast_2d(np.asarray(P, dtype=np.float64))
        if P.shape[1] != len(self.opt_idx):
            raise ValueError(f"expected {len(self.opt_idx)} parameters per vector, got {P.shape[1]}")
        out = np.tile(self._full, (P.shape[0], 1))
        out[:, self.opt_idx] = P
        return out

    def fn(self, par=None) -> float:
        par = self.par if par is None else np.asarray(par, dtype=np.float64)
        self.last_par = par.copy()
        return float(self.fn_batch(par[None, :])[0])

    def gr(self, par=None) -> np.ndarray:
        par = self.par if par is None else np.asarray(par, dtype=np.float64)
        self.last_par = par.copy()
        return self.gr_batch(par[None, :])[1]          # 1 x n matrix, like TMB

    def fn_batch(self, P) -> np.ndarray:
        return self._batch.eval_batch(self.full_par(P))[0]

    def nll_components_batch(self, P):
        nll, comp, _ = self._batch.eval_batch(self.full_par(P), want_comp=True)
        return nll, comp

    def gr_batch(self, P):
        nll, _, grad = self._batch.eval_batch(self.full_par(P), tangent_idx=self.opt_idx)
        return nll, grad

    def report(self, par=None) -> dict:
        par = self.par if par is None else np.asarray(par, dtype=np.float64)
        raw = self.handle.report(self.full_par(par[None, :])[0])
        return unpack_report(self.model, raw)


def unpack_report(model: G3CudaModel, raw: np.ndarray) -> dict:
    """Split the flat g3cuda_report() buffer into named arrays in the model's own dimension order
    (names as in the reference's REPORT()s, baseline/multiple-substocks.cpp:763-862)."""
    NT = model.n_steps
    out = OrderedDict()
    pos = 0
    out["nll"] = float(raw[pos]); pos += 1
    nn = len(model.nll_names)
    br = raw[pos:pos + nn * NT].reshape(nn, NT); pos += nn * NT
    for i, n in enumerate(model.comp_names):
        out[f"nll_{n}__{model.comp_units[i]}"] = br[i].copy()
    for i, n in enumerate(getattr(model, "random_names", [])):       # nll_random_dnorm_<name>__dnorm (R/likelihood_random.R:56-60)
        out[f"{n}__dnorm"] = br[len(model.comp_names) + i].copy()
    out["nll_understocking__wgt"] = br[nn - 2].copy()
    out["nll_bounds"] = br[nn - 1].copy()
    for n, (L, A_, R), dims in zip(model.stock_names, model.stock_shapes, model.stock_dims):
        nc = L * A_ * R
        for what in ("num", "wgt"):
            a = raw[pos:pos + NT * nc].reshape(NT, R, A_, L); pos += NT * nc
            order = [d for d in dims if d != "length"]
            # canonical (time, area, age, length) -> model order with length fastest, time last
            if order == ["age", "area"]:
                arr = a.transpose(3, 2, 1, 0)     # length, age, area, time
            else:
                arr = a.transpose(3, 1, 2, 0)     # length, area, age, time
            out[f"dstart_{n}__{what}"] = np.ascontiguousarray(arr)
    for n, u, sh in zip(model.comp_names, model.comp_units, model.comp_shapes):
        nd = sh["n_areag"] * sh["n_stockg"] * sh["n_age"] * sh["n_bins"]
        a = raw[pos:pos + sh["n_times"] * nd]; pos += sh["n_times"] * nd
        out[f"{n}_model__{u}"] = a.reshape(sh["n_times"], sh["n_areag"], sh["n_stockg"], sh["n_age"],
                                            sh["n_bins"]).copy()
    for n in model.comp_names:
        out[f"{n}_model__params"] = raw[pos:pos + 2].copy(); pos += 2

    def detail(n_stock):            # canonical [time][area][age][length] -> the stock's own order, time last
        L, A_, R = model.stock_shapes[n_stock]
        nonlocal pos
        a = raw[pos:pos + NT * L * A_ * R].reshape(NT, R, A_, L); pos += NT * L * A_ * R
        order = [d for d in model.stock_dims[n_stock] if d != "length"]
        return np.ascontiguousarray(a.transpose(3, 2, 1, 0) if order == ["age", "area"] else a.transpose(3, 1, 2, 0))

    # g3a_report_detail's remaining arrays (R/action_report.R:177-213; baseline/multiple-substocks.cpp:598-601,2045-2065).
    # Written by the thread-per-evaluation kernel; NaN when the report ran on the CTA kernel.
    for i, n in enumerate(model.stock_names):
        out[f"detail_{n}__renewalnum"] = detail(i)
    for pred, prey, PL in model.pp_info:
        si = model.stock_names.index(prey)
        L = model.stock_shapes[si][0]
        sr = raw[pos:pos + PL * L].copy(); pos += PL * L
        out[f"suit_{prey}_{pred}__report"] = sr if PL == 1 else np.ascontiguousarray(sr.reshape(PL, L).T)
        out[f"detail_{prey}_{pred}__suit"] = detail(si)
        out[f"detail_{prey}_{pred}__cons"] = detail(si)
        out[f"detail_{prey}__predby_{pred}"] = out[f"detail_{prey}_{pred}__cons"]
    if pos != len(raw):
        raise ValueError(f"report buffer has {len(raw)} doubles, layout needs {pos}")
    return out


def g3_cuda_adfun(model: G3CudaModel, parameters: ParameterTemplate = None, device=0, devices=None) -> G3CudaADFun:
    """``device``: one CUDA ordinal; ``devices``: several (one process drives them all)."""
    return G3CudaADFun(model, parameters, device, devices)
