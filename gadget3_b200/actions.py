"""Stock and fleet actions: host-side mirror of gadget3's g3a_* API for the supported subset.

Each function returns an action object; ``g3_to_cuda()`` (run_cuda.py) collates them in
``g3_action_order`` (R/action_order.R:1-17) and lowers them to the packed spec. Arguments keep the
reference's names and meaning; anything outside the supported subset raises G3Unsupported.
"""
from __future__ import annotations

import numpy as np

from .formula import (Expr, G3Unsupported, Var, age, cur_step_size, g3_exp, g3_log,
                      g3_parameterized, wrap)
from .stock import Stock

# R/action_order.R
g3_action_order = dict(initial=-1, time=1, predate=3, naturalmortality=4, grow=5, mature=7,
                       renewal=8, likelihood=10, report=11, age=12)


class Action:
    kind = ""


# ------------------------------------------------------------------ time
class TimeAction(Action):
    kind = "time"

    def __init__(self, start_year, end_year, step_lengths):
        self.start_year, self.end_year = int(start_year), int(end_year)
        self.step_lengths = [int(x) for x in step_lengths]
        if sum(self.step_lengths) != 12:
            raise ValueError("step_lengths should sum to 12 (i.e. represent a whole year)")


def g3a_time(start_year, end_year, step_lengths=(12,)):
    """R/action_time.R:21-106 (retro_years / project_years stay PARAMETERs of the model)."""
    return TimeAction(start_year, end_year, step_lengths)


# ------------------------------------------------------------------ renewal / initial conditions
def g3a_renewal_vonb_t0(Linf=None, K=None, t0=None, by_stock=True) -> Expr:
    """R/action_renewal.R:22-33."""
    Linf = g3_parameterized("Linf", value=1, by_stock=by_stock) if Linf is None else wrap(Linf)
    K = g3_parameterized("K", value=1, by_stock=by_stock) if K is None else wrap(K)
    t0 = g3_parameterized("t0", by_stock=by_stock) if t0 is None else wrap(t0)
    return Linf * (1 - g3_exp(-1 * K * (age - t0)))


def g3a_renewal_vonb_recl(Linf=None, K=None, recl=None, recage=None, by_stock=True) -> Expr:
    """R/action_renewal.R:8-21."""
    Linf = g3_parameterized("Linf", value=1, by_stock=by_stock) if Linf is None else wrap(Linf)
    K = g3_parameterized("K", value=1, by_stock=by_stock) if K is None else wrap(K)
    recl = g3_parameterized("recl", by_stock=by_stock) if recl is None else wrap(recl)
    recage = g3_parameterized("recage", by_stock=False, optimise=False) if recage is None else wrap(recage)
    return Linf * (1 - g3_exp(-1 * K * (age - (recage + g3_log(1 - recl / Linf) / K))))


g3a_renewal_vonb = g3a_renewal_vonb_recl


def g3a_renewal_initabund(scalar=None, init=None, M=None, init_F=None, recage=None,
                          proportion_f=1, by_stock=True, by_stock_f=False) -> Expr:
    """R/action_renewal.R:36-54."""
    scalar = g3_parameterized("init.scalar", value=1, by_stock=by_stock) if scalar is None else wrap(scalar)
    init = g3_parameterized("init", value=1, by_stock=by_stock, by_age=True) if init is None else wrap(init)
    M = g3_parameterized("M", by_stock=by_stock, by_age=True) if M is None else wrap(M)
    init_F = g3_parameterized("init.F", by_stock=by_stock_f) if init_F is None else wrap(init_F)
    recage = g3_parameterized("recage", by_stock=False, optimise=False) if recage is None else wrap(recage)
    out = scalar * init * g3_exp(-1 * (M + init_F) * (age - recage))
    if not (isinstance(proportion_f, (int, float)) and proportion_f == 1):
        out = out * wrap(proportion_f)
    return out


def _subst_age(e: Expr, repl: Expr) -> Expr:
    """f_substitute(e, list(age = repl))."""
    from .formula import BinOp, Lookup, ParamExpr, UnOp
    if isinstance(e, Var) and e.name == "age":
        return repl
    if isinstance(e, BinOp):
        return BinOp(e.op, _subst_age(e.a, repl), _subst_age(e.b, repl))
    if isinstance(e, UnOp):
        return UnOp(e.op, _subst_age(e.a, repl))
    if isinstance(e, Lookup):
        return Lookup(e.values, _subst_age(e.key, repl), e.base)
    return e


class RenewalAction(Action):
    """g3a_initialconditions_* / g3a_renewal_*: num = normalize_vec(dnorm(midlen, mean, avoid_zero(sd)))
    * 10000 * factor ; wgt = alpha * midlen^beta (R/action_renewal.R:56-80)."""

    def __init__(self, kind, stock, factor_f, mean_f, stddev_f, alpha_f, beta_f,
                 run_age=None, run_step=None, run_projection=True):
        self.kind = kind  # 'init' | 'renewal'
        self.stock = stock
        self.factor_f, self.mean_f, self.stddev_f = wrap(factor_f), wrap(mean_f), wrap(stddev_f)
        self.alpha_f, self.beta_f = wrap(alpha_f), wrap(beta_f)
        self.run_age, self.run_step, self.run_projection = run_age, run_step, run_projection


def g3a_initialconditions_normalparam(stock, factor_f=None, mean_f=None, stddev_f=None,
                                      alpha_f=None, beta_f=None, age_offset=cur_step_size,
                                      by_stock=True, by_age=False, wgt_by_stock=True):
    """R/action_renewal.R:105-137."""
    factor_f = g3a_renewal_initabund(by_stock=by_stock) if factor_f is None else wrap(factor_f)
    mean_f = g3a_renewal_vonb_t0(by_stock=by_stock) if mean_f is None else wrap(mean_f)
    stddev_f = (g3_parameterized("init.sd", value=10, by_stock=by_stock, by_age=by_age)
                if stddev_f is None else wrap(stddev_f))
    alpha_f = g3_parameterized("walpha", by_stock=wgt_by_stock) if alpha_f is None else wrap(alpha_f)
    beta_f = g3_parameterized("wbeta", by_stock=wgt_by_stock) if beta_f is None else wrap(beta_f)
    if age_offset is not None:
        # "pretending this is happening at time -1" (R/action_renewal.R:119-124)
        shifted = age - wrap(age_offset)
        mean_f = _subst_age(mean_f, shifted)
        stddev_f = _subst_age(stddev_f, shifted)
    return RenewalAction("init", stock, factor_f, mean_f, stddev_f, alpha_f, beta_f)


def g3a_initialconditions_normalcv(stock, factor_f=None, mean_f=None, cv_f=None, alpha_f=None,
                                   beta_f=None, age_offset=cur_step_size, by_stock=True,
                                   by_age=False, wgt_by_stock=True):
    """R/action_renewal.R:139-163."""
    mean_f = g3a_renewal_vonb_t0(by_stock=by_stock) if mean_f is None else wrap(mean_f)
    cv_f = (g3_parameterized("lencv", by_stock=by_stock, value=0.1, optimise=False)
            if cv_f is None else wrap(cv_f))
    return g3a_initialconditions_normalparam(stock, factor_f, mean_f, mean_f * cv_f, alpha_f, beta_f,
                                             age_offset, by_stock, by_age, wgt_by_stock)


def _default_rec_factor(by_stock):
    return g3_parameterized("rec", by_stock=by_stock, by_year=True,
                            scale=g3_parameterized("rec.scalar", by_stock=by_stock),
                            ifmissing=g3_parameterized("rec.proj", optimise=False, by_stock=by_stock))


def g3a_renewal_normalparam(stock, factor_f=None, mean_f=None, stddev_f=None, alpha_f=None,
                            beta_f=None, by_stock=True, wgt_by_stock=True, run_age="minage",
                            run_projection=True, run_step=1):
    """R/action_renewal.R:190-232. ``run_f`` is restricted to the age/step/projection form."""
    factor_f = _default_rec_factor(by_stock) if factor_f is None else wrap(factor_f)
    mean_f = g3a_renewal_vonb_t0(by_stock=by_stock) if mean_f is None else wrap(mean_f)
    stddev_f = (g3_parameterized("rec.sd", value=10, by_stock=by_stock)
                if stddev_f is None else wrap(stddev_f))
    alpha_f = g3_parameterized("walpha", by_stock=wgt_by_stock) if alpha_f is None else wrap(alpha_f)
    beta_f = g3_parameterized("wbeta", by_stock=wgt_by_stock) if beta_f is None else wrap(beta_f)
    if run_age == "minage":
        run_age = stock.minage
    return RenewalAction("renewal", stock, factor_f, mean_f, stddev_f, alpha_f, beta_f,
                         run_age=run_age, run_step=run_step, run_projection=run_projection)


def g3a_renewal_normalcv(stock, factor_f=None, mean_f=None, cv_f=None, alpha_f=None, beta_f=None,
                         by_stock=True, wgt_by_stock=True, run_age="minage", run_projection=True,
                         run_step=1):
    """R/action_renewal.R:234-272."""
    mean_f = g3a_renewal_vonb_t0(by_stock=by_stock) if mean_f is None else wrap(mean_f)
    cv_f = (g3_parameterized("lencv", by_stock=by_stock, value=0.1, optimise=False)
            if cv_f is None else wrap(cv_f))
    return g3a_renewal_normalparam(stock, factor_f, mean_f, mean_f * cv_f, alpha_f, beta_f,
                                   by_stock, wgt_by_stock, run_age, run_projection, run_step)


# ------------------------------------------------------------------ natural mortality
class OtherfoodAction(Action):
    """g3a_otherfood (R/action_renewal.R:276-308): at the start of every step num and wgt of the whole stock are
    ASSIGNED from formulas -- a prey without dynamics of its own."""
    kind = "otherfood"

    def __init__(self, stock, num_f, wgt_f):
        self.stock, self.num_f, self.wgt_f = stock, wrap(num_f), wrap(wgt_f)


def g3a_otherfood(stock, num_f=None, wgt_f=None, by_stock=True):
    """R/action_renewal.R:276-308. Defaults as the reference's: abundance ``of_abund`` by year, scaled by
    ``of_abund.step`` by step (``of_abund.proj`` outside the model's years), mean weight ``of_meanwgt``. The number is
    the same in every length group (force_lengthvector: ``n + 0 * stock__midlen``)."""
    if num_f is None:
        num_f = g3_parameterized("of_abund", by_year=True, by_stock=by_stock,
                                 scale=g3_parameterized("of_abund.step", value=1, by_step=True, by_stock=by_stock),
                                 ifmissing=g3_parameterized("of_abund.proj", by_stock=by_stock))
    if wgt_f is None:
        wgt_f = g3_parameterized("of_meanwgt", by_stock=by_stock)
    return OtherfoodAction(stock, num_f, wgt_f)


class MortalityAction(Action):
    kind = "naturalmortality"

    def __init__(self, stock, param_f):
        self.stock, self.param_f = stock, wrap(param_f)


class _MortExp:
    def __init__(self, param_f):
        self.param_f = param_f


def g3a_naturalmortality_exp(param_f=None, by_stock=True):
    """R/action_naturalmortality.R:1-9: exp(-param_f * cur_step_size)."""
    return _MortExp(g3_parameterized("M", by_stock=by_stock, by_age=True) if param_f is None else wrap(param_f))


def g3a_naturalmortality(stock, mortality_f=None):
    """R/action_naturalmortality.R:11-34. Only the g3a_naturalmortality_exp form is supported."""
    if mortality_f is None:
        mortality_f = g3a_naturalmortality_exp()
    if not isinstance(mortality_f, _MortExp):
        raise G3Unsupported("g3a_naturalmortality: only g3a_naturalmortality_exp() mortality is supported")
    return MortalityAction(stock, mortality_f.param_f)


# ------------------------------------------------------------------ growth / maturity
class _LengthVbSimple:
    def __init__(self, linf_f, kappa_f):
        self.linf_f, self.kappa_f = linf_f, kappa_f


class _WeightSimple:
    def __init__(self, alpha_f, beta_f):
        self.alpha_f, self.beta_f = alpha_f, beta_f


def g3a_grow_lengthvbsimple(linf_f=None, kappa_f=None, by_stock=True):
    """R/action_grow.R:44-55."""
    return _LengthVbSimple(
        g3_parameterized("Linf", by_stock=by_stock) if linf_f is None else wrap(linf_f),
        g3_parameterized("K", by_stock=by_stock) if kappa_f is None else wrap(kappa_f))


def g3a_grow_weightsimple(alpha_f=None, beta_f=None, by_stock=True):
    """R/action_grow.R:58-71."""
    return _WeightSimple(
        g3_parameterized("walpha", by_stock=by_stock) if alpha_f is None else wrap(alpha_f),
        g3_parameterized("wbeta", by_stock=by_stock) if beta_f is None else wrap(beta_f))


class _ImplBbinom:
    def __init__(self, delta_len_f, delta_wgt_f, beta_f, maxlengthgroupgrowth):
        self.delta_len_f, self.delta_wgt_f, self.beta_f = delta_len_f, delta_wgt_f, beta_f
        self.maxlengthgroupgrowth = int(maxlengthgroupgrowth)


def g3a_grow_impl_bbinom(delta_len_f=None, delta_wgt_f=None, beta_f=None, maxlengthgroupgrowth=None,
                         by_stock=True):
    """R/action_grow.R:145-230."""
    if maxlengthgroupgrowth is None:
        raise ValueError("maxlengthgroupgrowth is required")
    delta_len_f = g3a_grow_lengthvbsimple(by_stock=by_stock) if delta_len_f is None else delta_len_f
    delta_wgt_f = g3a_grow_weightsimple(by_stock=by_stock) if delta_wgt_f is None else delta_wgt_f
    if not isinstance(delta_len_f, _LengthVbSimple) or not isinstance(delta_wgt_f, _WeightSimple):
        raise G3Unsupported("g3a_grow_impl_bbinom: only lengthvbsimple / weightsimple growth is supported")
    beta_f = g3_parameterized("bbin", by_stock=by_stock) if beta_f is None else wrap(beta_f)
    return _ImplBbinom(delta_len_f, delta_wgt_f, beta_f, maxlengthgroupgrowth)


class _Maturity:
    def __init__(self, kind, alpha, l50, beta, a50):
        self.kind, self.alpha, self.l50, self.beta, self.a50 = kind, alpha, l50, beta, a50


def _opt(x):
    if x is None:
        return None
    if isinstance(x, (int, float)) and x == 0:
        return None
    return wrap(x)


def g3a_mature_continuous(alpha=None, l50=None, beta=0, a50=0, bounded=True, by_stock=True):
    """R/action_mature.R:1-42."""
    if not bounded:
        raise G3Unsupported("g3a_mature_continuous(bounded = FALSE)")
    alpha = g3_parameterized("mat.alpha", by_stock=by_stock) if alpha is None else wrap(alpha)
    l50 = g3_parameterized("mat.l50", by_stock=by_stock) if l50 is None else wrap(l50)
    return _Maturity("continuous", alpha, l50, _opt(beta), _opt(a50))


def g3a_mature_constant(alpha=None, l50=None, beta=None, a50=None, gamma=None, k50=None):
    """R/action_mature.R:44-74 (the gamma/k50 weight term is not supported)."""
    if _opt(gamma) is not None:
        raise G3Unsupported("g3a_mature_constant: gamma/k50 term")
    return _Maturity("constant", _opt(alpha), _opt(l50), _opt(beta), _opt(a50))


class GrowMatureAction(Action):
    kind = "growmature"

    def __init__(self, stock, impl_f, maturity_f, output_stocks, output_ratios, transition_f):
        self.stock, self.impl_f, self.maturity_f = stock, impl_f, maturity_f
        self.output_stocks, self.output_ratios = list(output_stocks), list(output_ratios)
        self.transition_f = transition_f


def g3a_growmature(stock, impl_f, maturity_f=None, output_stocks=(), output_ratios=None,
                   transition_f="cur_step_final"):
    """R/action_grow.R:340-497. ``transition_f``: 'cur_step_final' (default), True, or a list of
    1-based steps on which maturing fish move."""
    output_stocks = list(output_stocks)
    if output_ratios is None:
        output_ratios = [1.0 / max(len(output_stocks), 1)] * len(output_stocks)
    if not isinstance(impl_f, _ImplBbinom):
        raise G3Unsupported("g3a_growmature: impl_f must come from g3a_grow_impl_bbinom()")
    if maturity_f is not None and not isinstance(maturity_f, _Maturity):
        raise G3Unsupported("g3a_growmature: maturity_f must be g3a_mature_continuous/constant")
    if len(output_stocks) == 0:
        maturity_f = None  # R/action_grow.R:372-374: all maturity code is thrown away
    return GrowMatureAction(stock, impl_f, maturity_f, output_stocks, output_ratios, transition_f)


# ------------------------------------------------------------------ ageing
class AgeAction(Action):
    kind = "age"

    def __init__(self, stock, output_stocks, output_ratios):
        self.stock = stock
        self.output_stocks, self.output_ratios = list(output_stocks), list(output_ratios)


def g3a_age(stock, output_stocks=(), output_ratios=None):
    """R/action_age.R:2-106 (run_f fixed to cur_step_final)."""
    output_stocks = list(output_stocks)
    if output_ratios is None:
        output_ratios = [1.0 / max(len(output_stocks), 1)] * len(output_stocks)
    return AgeAction(stock, output_stocks, output_ratios)


# ------------------------------------------------------------------ predation
class _Suitability:
    def __init__(self, kind, args):
        self.kind, self.args = kind, args


def g3_suitability_exponentiall50(alpha=None, l50=None, by_stock=True, by_predator=True):
    """R/suitability.R:5-13."""
    alpha = g3_parameterized("alpha", by_stock=by_stock, by_predator=by_predator) if alpha is None else wrap(alpha)
    l50 = g3_parameterized("l50", by_stock=by_stock, by_predator=by_predator) if l50 is None else wrap(l50)
    return _Suitability("expl50", [alpha, l50])


def g3_suitability_andersen(p0, p1, p2, p3=None, p4=None, p5="predator_length"):
    """R/suitability.R:15-30 (p5 must stay the predator length)."""
    if p5 != "predator_length":
        raise G3Unsupported("g3_suitability_andersen: p5 other than predator_length")
    if p4 is None:
        raise ValueError("p4 is required")
    if p3 is None:
        p3 = p4
    return _Suitability("andersen", [wrap(p0), wrap(p1), wrap(p2), wrap(p3), wrap(p4)])


def g3_suitability_constant(suit=None, by_stock=True, by_predator=True):
    """R/suitability.R (constant)."""
    suit = g3_parameterized("suit", by_stock=by_stock, by_predator=by_predator) if suit is None else wrap(suit)
    return _Suitability("constant", [suit])


def g3_suitability_andersenfleet(p0=None, p1=None, p2=None, p3=None, p4=None, by_stock=True, by_predator=True,
                                 exponentiate=True):
    """R/suitability.R:32-58 (p5 stays the prey's largest midlen)."""
    import math
    gp = g3_parameterized
    p0 = gp("andersen.p0", value=0, optimise=False, by_stock=by_stock) if p0 is None else wrap(p0)
    p1 = gp("andersen.p1", value=math.log(2), by_stock=by_stock, by_predator=by_predator) if p1 is None else wrap(p1)
    p2 = gp("andersen.p2", value=1, optimise=False, by_stock=by_stock) if p2 is None else wrap(p2)
    p3 = gp("andersen.p3", value=0.1, exponentiate=exponentiate, by_stock=by_stock, by_predator=by_predator) if p3 is None else wrap(p3)
    p4 = gp("andersen.p4", value=0.1, exponentiate=exponentiate, by_stock=by_stock, by_predator=by_predator) if p4 is None else wrap(p4)
    return _Suitability("andersenfleet", [p0, p1, p2, p3, p4])


def g3_suitability_gamma(alpha, beta, gamma):
    """R/suitability.R:60-67."""
    return _Suitability("gamma", [wrap(alpha), wrap(beta), wrap(gamma)])


def g3_suitability_exponential(alpha, beta, gamma, delta):
    """R/suitability.R:69-75 (reads predator_length: stock predators only)."""
    return _Suitability("exponential", [wrap(alpha), wrap(beta), wrap(gamma), wrap(delta)])


def g3_suitability_straightline(alpha, beta):
    """R/suitability.R:77-79."""
    return _Suitability("straightline", [wrap(alpha), wrap(beta)])


def g3_suitability_richards(p0, p1, p2, p3, p4):
    """R/suitability.R:90-94."""
    return _Suitability("richards", [wrap(p0), wrap(p1), wrap(p2), wrap(p3), wrap(p4)])


class TimeAreaData:
    """g3_timeareadata (R/timedata.R:89-157): year/step/[area] -> value lookup, default 0."""

    def __init__(self, lookup_name, df, value_field="total_weight", areas=None):
        self.name = lookup_name
        cols = {k: np.asarray(v) for k, v in dict(df).items()}
        if value_field not in cols:
            raise ValueError(f"No {value_field} field in g3_timeareadata data.frame")
        n = len(cols[value_field])
        self.year = cols["year"].astype(int) if "year" in cols else None
        self.step = cols["step"].astype(int) if "step" in cols else None
        area = cols.get("area")
        if area is not None and areas is not None and not all(str(a) == str(areas.get(str(a), None)) for a in area):
            area = np.asarray([int(areas[str(a)]) for a in area])
        self.area = None if area is None else np.asarray(area).astype(int)
        self.value = cols[value_field].astype(float)
        # single area -> condition outside the table (R/timedata.R:106-112)
        self.our_area = None
        if self.area is not None and len(set(self.area.tolist())) == 1:
            self.our_area = int(self.area[0])
            self.area = None
        self.table = {}
        for i in range(n):
            key = (None if self.year is None else int(self.year[i]),
                   None if self.step is None else int(self.step[i]),
                   None if self.area is None else int(self.area[i]))
            self.table[key] = float(self.value[i])

    def lookup(self, year, step, area) -> float:
        if self.our_area is not None and area != self.our_area:
            return 0.0
        key = (None if self.year is None else year, None if self.step is None else step,
               None if self.area is None else area)
        return self.table.get(key, 0.0)


def g3_timeareadata(lookup_name, df, value_field="total_weight", areas=None):
    return TimeAreaData(lookup_name, df, value_field, areas)


class _Catchability:
    def __init__(self, kind, E=None, **kw):
        self.kind, self.E, self.kw = kind, E, kw


def g3a_predate_catchability_totalfleet(E):
    """R/action_predate.R:1-5. ``E`` must be a g3_timeareadata() table."""
    if not isinstance(E, TimeAreaData):
        raise G3Unsupported("g3a_predate_catchability_totalfleet: E must be g3_timeareadata()")
    return _Catchability("totalfleet", E)


def _fleet_E(E, what):
    if not isinstance(E, TimeAreaData):
        raise G3Unsupported(f"{what}: E must be g3_timeareadata()")
    return E


def g3a_predate_catchability_numberfleet(E):
    """R/action_predate.R:7-11: landings in individuals."""
    return _Catchability("numberfleet", _fleet_E(E, "g3a_predate_catchability_numberfleet"))


def g3a_predate_catchability_linearfleet(E):
    """R/action_predate.R:13-18: E is a harvest rate per year, scaled by cur_step_size."""
    return _Catchability("linearfleet", _fleet_E(E, "g3a_predate_catchability_linearfleet"))


def g3a_predate_catchability_effortfleet(catchability_fs, E):
    """R/action_predate.R:117-126: cons = catchability_fs[prey] * E * cur_step_size * suit.
    ``catchability_fs``: one formula, or a dict prey stock name -> formula."""
    return _Catchability("effortfleet", _fleet_E(E, "g3a_predate_catchability_effortfleet"), catchability_fs=catchability_fs)


def g3a_predate_catchability_predator(prey_preferences=1, energycontent=None, half_feeding_f=None,
                                      max_consumption=None, temperature=0, by_predator=True, by_stock=True):
    """R/action_predate.R:207-238."""
    energycontent = (g3_parameterized("energycontent", value=1, by_stock=by_stock, optimise=False)
                     if energycontent is None else wrap(energycontent))
    half_feeding_f = (g3_parameterized("halffeeding", by_predator=by_predator, optimise=False)
                      if half_feeding_f is None else wrap(half_feeding_f))
    if max_consumption is None:
        max_consumption = g3a_predate_maxconsumption(temperature=temperature)
    return _Catchability("predator", None, prey_preferences=prey_preferences, energycontent=energycontent,
                         half_feeding_f=half_feeding_f, max_consumption=max_consumption)


def g3a_predate_maxconsumption(m0=None, m1=None, m2=None, m3=None, temperature=0):
    """R/action_predate.R:190-205."""
    P = lambda n, v: g3_parameterized("consumption." + n, value=v, by_predator=True, optimise=False)
    return dict(m0=P("m0", 1) if m0 is None else wrap(m0), m1=P("m1", 0) if m1 is None else wrap(m1),
                m2=P("m2", 0) if m2 is None else wrap(m2), m3=P("m3", 0) if m3 is None else wrap(m3),
                temperature=wrap(temperature))


class PredateAction(Action):
    kind = "predate"

    def __init__(self, predstock, prey_stocks, suitabilities, catchability_f):
        self.predstock, self.prey_stocks = predstock, list(prey_stocks)
        self.suitabilities, self.catchability_f = suitabilities, catchability_f

    def suit_for(self, stock):
        s = self.suitabilities
        if isinstance(s, dict):
            return s[stock.name]
        return s


def g3a_predate(predstock, prey_stocks, suitabilities, catchability_f):
    """R/action_predate.R:240-420 (default overconsumption_f, run_f = TRUE)."""
    if not isinstance(catchability_f, _Catchability):
        raise G3Unsupported("g3a_predate: unsupported catchability_f")
    return PredateAction(predstock, prey_stocks, suitabilities, catchability_f)


def g3a_predate_fleet(fleet_stock, prey_stocks, suitabilities, catchability_f):
    """R/action_predate.R:423-476."""
    return g3a_predate(fleet_stock, prey_stocks, suitabilities, catchability_f)


# ------------------------------------------------------------------ spawning
class _Recruitment:
    def __init__(self, kind, s_args, r_args, R_by_year=None):
        self.kind, self.s_args, self.r_args, self.R_by_year = kind, [wrap(x) for x in s_args], [wrap(x) for x in r_args], R_by_year


def _sp(name, by_stock, value=0.0, **kw):
    return g3_parameterized("spawn." + name, value=value, by_stock=by_stock, **kw)


def g3a_spawn_recruitment_fecundity(p0=None, p1=None, p2=None, p3=None, p4=None, by_stock=True):
    """R/action_spawn.R:1-17: s = sum(midlen^p1 * age^p2 * spawningnum^p3 * wgt^p4), r = p0 * s."""
    ps = [_sp(f"p{i}", by_stock, 1) if x is None else x for i, x in enumerate((p0, p1, p2, p3, p4))]
    return _Recruitment("fecundity", ps[1:], ps[:1])


def g3a_spawn_recruitment_simplessb(mu=None, by_stock=True):
    """R/action_spawn.R:19-28: s = sum(wgt * spawningnum), r = mu * s."""
    return _Recruitment("simplessb", [], [_sp("mu", by_stock) if mu is None else mu])


def g3a_spawn_recruitment_ricker(mu=None, lam=None, by_stock=True):
    """R/action_spawn.R:32-43: r = mu * s * exp(-lambda * s)."""
    return _Recruitment("ricker", [], [_sp("mu", by_stock) if mu is None else mu, _sp("lambda", by_stock) if lam is None else lam])


def g3a_spawn_recruitment_bevertonholt(mu=None, lam=None, by_stock=True):
    """R/action_spawn.R:45-56: r = mu * s / (lambda + s)."""
    return _Recruitment("bevertonholt", [], [_sp("mu", by_stock) if mu is None else mu, _sp("lambda", by_stock) if lam is None else lam])


def g3a_spawn_recruitment_bevertonholt_ss3(h=None, R=None, B0=None, by_stock=True):
    """R/action_spawn.R:59-74: r = 4 h s R / (B0 (1 - h) + s (5 h - 1)); R = exp(spawn.R.<year>) * spawn.R0 by default."""
    h = _sp("h", by_stock, 0.5, lower=0.1, upper=1) if h is None else h
    R = (g3_parameterized("spawn.R", by_year=True, exponentiate=True, scale="spawn.R0", by_stock=by_stock) if R is None else R)
    B0 = _sp("B0", by_stock) if B0 is None else B0
    return _Recruitment("bevertonholt_ss3", [], [h, B0], R_by_year=wrap(R))


def g3a_spawn_recruitment_hockeystick(r0=None, blim=None, by_stock=True):
    """R/action_spawn.R:76-88: r = r0 * logspace_add(-1000 * s / blim, -1000) / -1000 (a differentiable r0 * pmin(s / blim, 1))."""
    return _Recruitment("hockeystick", [], [_sp("r0", by_stock) if r0 is None else r0, _sp("blim", by_stock, 1) if blim is None else blim])


class SpawnAction(Action):
    kind = "spawn"

    def __init__(self, stock, recruitment_f, proportion_f, mortality_f, output_stocks, output_ratios, mean_f, stddev_f,
                 alpha_f, beta_f, run_step):
        self.stock, self.recruitment_f, self.proportion_f, self.mortality_f = stock, recruitment_f, proportion_f, mortality_f
        self.output_stocks, self.output_ratios = list(output_stocks), list(output_ratios)
        self.mean_f, self.stddev_f, self.alpha_f, self.beta_f, self.run_step = mean_f, stddev_f, alpha_f, beta_f, run_step
        self.weightloss = None      # (rel_loss, min_weight) formulas


def g3a_spawn(stock, recruitment_f, proportion_f=1, mortality_f=0, weightloss_f=0, weightloss_args=None, output_stocks=(),
              output_ratios=None, mean_f=None, stddev_f=None, alpha_f=None, beta_f=None, by_stock=True, wgt_by_stock=True,
              run_step=None):
    """R/action_spawn.R:86-233. ``proportion_f`` / ``mortality_f``: a formula, or a g3_suitability_*() of the parent's
    length (as the reference's documentation uses them). mean_f .. beta_f are evaluated for each OUTPUT stock
    (R/action_spawn.R:187-189). Spawning weight loss: the relative form (weightloss_f, or weightloss_args with rel_loss and
    min_weight); abs_loss is not supported."""
    if not isinstance(recruitment_f, _Recruitment):
        raise G3Unsupported("g3a_spawn: recruitment_f must come from g3a_spawn_recruitment_*()")
    # weightloss_f is rewritten as weightloss_args = list(rel_loss = weightloss_f, min_weight = 0) (R/action_spawn.R:166-171)
    wl_args = dict(weightloss_args or {})
    if not (isinstance(weightloss_f, (int, float)) and weightloss_f == 0):
        wl_args = dict(rel_loss=weightloss_f, min_weight=0)
    if wl_args.get("abs_loss") is not None:
        raise G3Unsupported("g3a_spawn(weightloss_args = list(abs_loss = ...))")
    if wl_args and wl_args.get("rel_loss") is None:
        raise G3Unsupported("g3a_spawn(weightloss_args) without rel_loss")
    output_stocks = list(output_stocks)
    if output_ratios is None:
        output_ratios = [1.0 / len(output_stocks)] * len(output_stocks) if output_stocks else []
    output_ratios = list(output_ratios)
    if len(output_stocks) != len(output_ratios):
        raise ValueError("length(output_stocks) == length(output_ratios) is not TRUE")
    if all(isinstance(x, (int, float)) for x in output_ratios) and output_ratios and abs(sum(output_ratios) - 1) >= 1e-4:
        raise ValueError("output_ratios must sum to 1")
    mean_f = g3a_renewal_vonb_t0(by_stock=by_stock) if mean_f is None else wrap(mean_f)
    stddev_f = g3_parameterized("rec.sd", value=10, by_stock=by_stock) if stddev_f is None else wrap(stddev_f)
    alpha_f = g3_parameterized("walpha", by_stock=wgt_by_stock) if alpha_f is None else wrap(alpha_f)
    beta_f = g3_parameterized("wbeta", by_stock=wgt_by_stock) if beta_f is None else wrap(beta_f)
    act = SpawnAction(stock, recruitment_f, proportion_f, mortality_f, output_stocks, output_ratios, mean_f, stddev_f,
                      alpha_f, beta_f, run_step)
    if wl_args:     # g3a_weightloss applied to the spawning part of every cell (R/action_weightloss.R:19-44; min_weight defaults to 1e-7)
        act.weightloss = (wrap(wl_args["rel_loss"]), wrap(wl_args.get("min_weight", 1e-7)))
    return act


# ------------------------------------------------------------------ migration
class MigrateAction(Action):
    kind = "migrate"

    def __init__(self, stock, migrate_f, run_steps):
        self.stock, self.migrate_f, self.run_steps = stock, migrate_f, run_steps


def g3a_migrate(stock, migrate_f, run_steps=None):
    """R/action_migrate.R:17-82. ``migrate_f(src_name, dest_name) -> Expr`` gives the
    proportion moving from src to dest; ``run_steps``: 1-based steps it runs on (None = all)."""
    return MigrateAction(stock, migrate_f, run_steps)


# ------------------------------------------------------------------ reports
class ReportDetailAction(Action):
    kind = "report_detail"


def g3a_report_detail(actions=None):
    """R/action_report.R:177-213: registers the report_detail parameter; the dstart_* histories are
    always produced by g3cuda_report()."""
    return ReportDetailAction()
