"""Likelihood actions: host-side mirror of g3l_* for the supported subset
(R/likelihood_distribution.R, R/likelihood_data.R, R/likelihood_understocking.R, R/likelihood_bounds.R)."""
from __future__ import annotations

import math
import re

import numpy as np

from .actions import Action
from .formula import G3Unsupported, g3_parameterized, wrap
from .stock import g3s_length, Stock


class _DistFn:
    def __init__(self, name, **kw):
        self.name, self.kw = name, kw


def g3l_distribution_sumofsquares(over=("area", "predator", "predator_tag", "predator_age", "predator_length")):
    """R/likelihood_distribution.R:3-39. With 'length' in ``over`` the sum of squares is STRATIFIED: proportions are taken
    within each length group (total = avoid_zero(rowSums(...)) over the other dimensions, dist_prop() :8-16) -- the
    usual form for maturity-at-length data."""
    if "area" not in over:
        raise G3Unsupported("g3l_distribution_sumofsquares: 'area' must be in over")
    return _DistFn("sumofsquares", stratified="length" in over)


def g3l_distribution_multinomial(epsilon=10):
    """R/likelihood_distribution.R:41-60."""
    return _DistFn("multinomial", epsilon=float(epsilon))


def g3l_distribution_sumofsquaredlogratios(epsilon=10):
    """R/likelihood_distribution.R:127-136 (catch-in-kilos style sum of squared log ratios)."""
    return _DistFn("sumofsquaredlogratios", epsilon=float(epsilon))


def g3l_distribution_surveyindices_log(alpha=None, beta=1):
    """R/likelihood_distribution.R:82-125."""
    return _DistFn("surveyindices_log", alpha=alpha, beta=beta)


def g3l_distribution_surveyindices_linear(alpha=None, beta=1):
    return _DistFn("surveyindices_linear", alpha=alpha, beta=beta)


# ------------------------------------------------------------------ observation data -> dense arrays
_CUT_RE = re.compile(r"^(\[|\()(.*),(.*)(\]|\))$")
_RANGE_RE = re.compile(r"^(.+):(.+)$")


def _parse_levels(lvls):
    """R/likelihood_data.R:1-33: returns list of (name, lower, upper) and open_ended flag."""
    try:
        # numeric lower bounds: sorted, each group ends where the next begins, the last one is open-ended
        # (parse_levels() on the sorted factor levels, R/likelihood_data.R:1-11,275-281)
        order = sorted(range(len(lvls)), key=lambda i: float(lvls[i]))
        m = [float(lvls[i]) for i in order]
        ups = m[1:] + [math.inf]
        return [(str(lvls[i]), lo, up) for i, lo, up in zip(order, m, ups)], True
    except (TypeError, ValueError):
        pass
    out = []
    for s in lvls:
        s = str(s)
        m = _CUT_RE.match(s)
        if m:
            if m.group(1) != "[" or m.group(4) == "]":
                raise ValueError("length intervals should be inclusive-lower, i.e. cut(..., right=FALSE): " + s)
            out.append((s, float(m.group(2)), float(m.group(3))))
            continue
        m = _RANGE_RE.match(s)
        if m:
            out.append((s, float(m.group(1)), float(m.group(2))))
            continue
        raise ValueError("Unknown form of length levels, see ?cut for formatting: " + s)
    out.sort(key=lambda t: (t[1], t[2]))
    return out, math.isinf(out[-1][2])


class LikelihoodData:
    """Mirror of g3l_likelihood_data() (R/likelihood_data.R:35-223) for data given as a dict of
    equal-length columns: year, [step], [area], [age], [length], [stock], and number | weight."""

    def __init__(self, nll_name, data, all_stocks, area_group=None, missing_val=0.0):
        cols = {k: list(v) for k, v in dict(data).items()}
        n = len(next(iter(cols.values())))
        if "year" not in cols:
            raise ValueError("Data must contain a year column")
        self.has_step = "step" in cols
        years = [int(y) for y in cols.pop("year")]
        steps = [int(s) for s in cols.pop("step")] if self.has_step else [0] * n
        times = [y * 100 + s for y, s in zip(years, steps)]
        self.times = sorted(set(times))          # R/stock_time.R:1-22, keys year*100+step
        time_idx = {t: i for i, t in enumerate(self.times)}

        # length dimension (R/likelihood_data.R:225-291)
        if "length" in cols:
            raw = cols.pop("length")
            seen = []
            for x in raw:
                if x not in seen:
                    seen.append(x)
            lv, open_ended = _parse_levels(seen)
            for a, b in zip(lv[:-1], lv[1:]):
                if not math.isclose(a[2], b[1]):
                    raise ValueError("Gaps in length groups are not supported")
            bounds = [t[1] for t in lv] + ([] if open_ended else [lv[-1][2]])
            self.lenstock = g3s_length(Stock("x", {}), bounds, open_ended=open_ended)
            lmap = {t[0]: i for i, t in enumerate(lv)}
            lmap.update({float(t[0]): i for i, t in enumerate(lv) if _is_num(t[0])})
            len_idx = [lmap[x if x in lmap else str(x)] for x in raw]
            self.has_length = True
        else:
            self.lenstock = g3s_length(Stock("x", {}), [0.0])
            len_idx = [0] * n
            self.has_length = False
        self.n_bins = self.lenstock.L

        # age dimension: numeric ages, no grouping (R/likelihood_data.R:293-330)
        if "age" in cols:
            ages = [int(a) for a in cols.pop("age")]
            self.minage, self.maxage = min(ages), max(ages)
            age_idx = [a - self.minage for a in ages]
            self.n_age = self.maxage - self.minage + 1
        else:
            self.minage = self.maxage = None
            age_idx = [0] * n
            self.n_age = 1

        # stock dimension (R/likelihood_data.R ld_dim_g3stock): names may be partial ("imm")
        if "stock" in cols:
            import itertools
            names = [str(x) for x in cols.pop("stock")]
            self.stock_groups = []            # unique, in order of first appearance (R/likelihood_data.R:384)
            for x in names:
                if x not in self.stock_groups:
                    self.stock_groups.append(x)
            stock_idx = [self.stock_groups.index(x) for x in names]
            # shortest name-part combinations first, longer matches override (R/likelihood_data.R:391-406)
            self.stock_map = {s.name: -1 for s in all_stocks}
            maxparts = max(len(s.name.split("_")) for s in all_stocks)
            for k in range(1, maxparts + 1):
                for s in all_stocks:
                    parts = s.name.split("_")
                    if k > len(parts):
                        continue
                    combos = ["_".join(c) for c in itertools.combinations(parts, k)]
                    hits = [i for i, g in enumerate(self.stock_groups) if g in combos]
                    if hits:
                        self.stock_map[s.name] = hits[0]
            unused = set(range(len(self.stock_groups))) - set(self.stock_map.values())
            if unused:
                raise ValueError("stock groups matched no stock in likelihood data: "
                                 + ", ".join(self.stock_groups[i] for i in sorted(unused)))
        else:
            self.stock_groups = None
            stock_idx = [0] * n
            self.stock_map = None
        self.n_stockg = len(self.stock_groups) if self.stock_groups else 1

        # area dimension (R/likelihood_data.R:170-190)
        if "area" in cols:
            areas = [str(a) for a in cols.pop("area")]
            used = sorted(set(areas))
            if area_group is None:
                area_group = {a: int(a) for a in used}
            self.area_group = {a: area_group[a] for a in sorted(used)}
            anames = list(self.area_group.keys())
            area_idx = [anames.index(a) for a in areas]
        else:
            self.area_group = None
            area_idx = [0] * n
        self.n_areag = len(self.area_group) if self.area_group else 1

        obs_cols = [c for c in cols if c in ("number", "weight")]
        if len(obs_cols) != 1:
            raise G3Unsupported("likelihood data must have exactly one of number / weight columns, got %r" % list(cols))
        self.unit = "num" if obs_cols[0] == "number" else "wgt"
        vals = [float(v) for v in cols[obs_cols[0]]]

        self.n_slices = self.n_stockg * self.n_age
        self.n_dest = self.n_areag * self.n_slices * self.n_bins
        obs = np.full((len(self.times), self.n_dest), float(missing_val))
        for i in range(n):
            e = (area_idx[i] * self.n_slices + stock_idx[i] * self.n_age + age_idx[i]) * self.n_bins + len_idx[i]
            obs[time_idx[times[i]], e] = vals[i]
        self.obs = obs

    def dest(self, areag, stockg, ageg, b):
        return (areag * self.n_slices + stockg * self.n_age + ageg) * self.n_bins + b


def _is_num(x):
    try:
        float(x)
        return True
    except (TypeError, ValueError):
        return False


def _check_transforms(transform_fs, stocks):
    """transform_fs (R/likelihood_distribution.R:184,298-347): 'length' -- the collected vector of a stock is multiplied by an
    L x L matrix before it is reshaped onto the observation's length groups; 'age' -- what is collected at age `preage`
    lands in age `age` with the factor matrix[preage_idx, age_idx] (an age-reading error matrix). One matrix, or one by
    stock name (list_to_stock_switch; a stock without an entry is not transformed). The matrices are DATA here: they
    are folded into the gather coefficients at lowering time; a matrix of parameters (g3_param_array) is refused.
    Returns {stock name: {'length': ndarray | None, 'age': ndarray | None}}."""
    out = {s.name: {"length": None, "age": None} for s in stocks}
    for dim, tf in (transform_fs or {}).items():
        if dim not in ("length", "age"):
            raise G3Unsupported(f"Transforms for {dim} dimensions not supported yet")
        for s in stocks:
            m = tf.get(s.name) if isinstance(tf, dict) else tf
            if m is None:
                continue
            try:
                m = np.asarray(m, dtype=float)
            except (TypeError, ValueError):
                raise G3Unsupported("g3l_distribution(transform_fs = ...): only constant matrices are supported") from None
            n = s.L if dim == "length" else s.nage
            if m.shape != (n, n):
                raise ValueError(f"transform_fs${dim} for {s.name} must be a {n} x {n} matrix, got {m.shape}")
            out[s.name][dim] = m
    return out


class DistributionAction(Action):
    kind = "distribution"

    def __init__(self, nll_name, obs_data, fleets, stocks, function_f, area_group, report, nll_breakdown,
                 weight, missing_val, transform_fs=None):
        self.fleets, self.stocks = list(fleets), list(stocks)
        self.transform_fs = _check_transforms(transform_fs, self.stocks)
        self.function_f = function_f
        self.report, self.nll_breakdown = report, nll_breakdown
        # R/likelihood_distribution.R:266-271
        self.nll_name = "_".join(["cdist" if self.fleets else "adist", function_f.name, nll_name])
        self.weight = (g3_parameterized(self.nll_name + "_weight", optimise=False, value=1)
                       if weight is None else wrap(weight))
        self.ld = LikelihoodData(self.nll_name, obs_data, self.stocks, area_group, missing_val)


def g3l_distribution(nll_name, obs_data, fleets=(), stocks=(), function_f=None, area_group=None,
                     report=False, nll_breakdown=False, weight=None, missing_val=0, transform_fs=None,
                     predators=()):
    """R/likelihood_distribution.R:177-440."""
    # predators (stock predators) are collected exactly as fleets are (fleetpreds <- c(fleets, predators),
    # R/likelihood_distribution.R:200,275-292): predprey__cons, summed over the predator's own dimensions unless the data
    # carries predator_length / predator_age / predator_tag columns -- those are not supported
    if predators:
        cols = obs_data.keys() if hasattr(obs_data, "keys") else ()
        bad = [c for c in cols if str(c).startswith("predator_")]
        if bad:
            raise G3Unsupported("g3l_distribution(predators = ...) with %s columns in the data" % ", ".join(bad))
        fleets = list(fleets) + list(predators)
    if not isinstance(function_f, _DistFn):
        raise G3Unsupported("g3l_distribution: unsupported function_f")
    return DistributionAction(nll_name, obs_data, fleets, stocks, function_f, area_group, report,
                              nll_breakdown, weight, missing_val, transform_fs)


def g3l_catchdistribution(nll_name, obs_data, fleets=(), stocks=(), function_f=None, **kw):
    if len(fleets) == 0 and len(kw.get("predators", ())) == 0:
        raise ValueError("fleets/predators must be supplied for g3l_catchdistribution")
    return g3l_distribution(nll_name, obs_data, fleets, stocks, function_f, **kw)


def g3l_abundancedistribution(nll_name, obs_data, fleets=(), stocks=(), function_f=None, **kw):
    if len(fleets) > 0 or len(kw.get("predators", ())) > 0:
        raise ValueError("fleets/predators must not be supplied for g3l_abundancedistribution")
    return g3l_distribution(nll_name, obs_data, (), stocks, function_f, **kw)


class UnderstockingAction(Action):
    kind = "understocking"

    def __init__(self, prey_stocks, power, weight, nll_breakdown):
        self.prey_stocks, self.power, self.weight = list(prey_stocks), float(power), float(weight)
        self.nll_breakdown = nll_breakdown


def g3l_understocking(prey_stocks, power_f=2, nll_breakdown=False, weight=1e8):
    """R/likelihood_understocking.R:1-53."""
    return UnderstockingAction(prey_stocks, power_f, weight, nll_breakdown)


class BoundsPenaltyAction(Action):
    kind = "bounds_penalty"

    def __init__(self, weight, scale):
        self.weight, self.scale = float(weight), float(scale)


def g3l_bounds_penalty(actions_or_parameter_template=None, weight=1, scale=1e6):
    """R/likelihood_bounds.R:2-65, "fromactions" form: one penalty per parameter with finite
    lower & upper bound; the bounds are bound at g3_cuda_adfun() time like the __lower/__upper
    DATA_SCALARs (R/run_tmb.R:1234-1235)."""
    from .formula import ParameterTemplate
    if isinstance(actions_or_parameter_template, ParameterTemplate):
        # the parameter_template form penalises only optimise = TRUE rows and bakes today's bounds into the action
        # (R/likelihood_bounds.R:44-62): a different nll from the one this backend would compute
        raise G3Unsupported("g3l_bounds_penalty(parameter_template): only the actions form is supported "
                            "(bounds are bound at g3_cuda_adfun() time)")
    return BoundsPenaltyAction(weight, scale)


# --------------------------------------------------------------------------- random-effect penalties
def _param_columns(e):
    """The table columns of the g3_parameterized() calls inside ``e`` (what g3_parameterized_breakdown reports,
    R/likelihood_random.R:1-12)."""
    from .formula import BinOp, ParamExpr, UnOp
    cols = set()

    def walk(x):
        if isinstance(x, ParamExpr):
            for flag, col in ((x.by_stock, "stock"), (x.by_predator, "predator"), (x.by_year, "cur_year"),
                              (x.by_step, "cur_step"), (x.by_age, "age"), (x.by_area, "area")):
                if flag:
                    cols.add(col)
            if x.ifmissing is not None and not isinstance(x.ifmissing, (int, float)):
                walk(x.ifmissing)
        elif isinstance(x, BinOp):
            walk(x.a), walk(x.b)
        elif isinstance(x, UnOp):
            walk(x.a)
    walk(e)
    return sorted(cols)


def _guess_period(param_f):
    """R/likelihood_random.R:1-12."""
    cols = _param_columns(param_f)
    if not cols:
        return "single"
    if "stock" in cols:
        raise ValueError("Per-stock random parameters not supported yet")
    if cols == ["cur_step", "cur_year"]:
        return "step"
    if cols == ["cur_year"]:
        return "year"
    if cols == ["cur_step"]:
        raise ValueError("Seasonal random parameters not supported yet")
    raise ValueError("Cannot use period='auto', set period manually: %s" % ", ".join(cols))


class RandomAction(Action):
    kind = "random"

    def __init__(self, walk, nll_name, param_f, mean_f, sigma_f, log_f, period, nll_breakdown, weight):
        from .formula import g3_parameterized, wrap
        if period not in ("year", "step", "single", "auto"):
            raise ValueError("period must be one of 'year', 'step', 'single', 'auto'")
        if not isinstance(log_f, bool):
            raise ValueError("log_f must be logical")
        self.walk, self.nll_name = bool(walk), nll_name
        self.param_f, self.sigma_f = wrap(param_f), wrap(sigma_f)
        self.mean_f = None if walk else wrap(mean_f)
        self.log_f = log_f
        self.period = _guess_period(self.param_f) if period == "auto" else period
        self.nll_breakdown = nll_breakdown
        self.weight = wrap(g3_parameterized(nll_name + "_weight", optimise=False, value=1) if weight is None else weight)

    @property
    def report_name(self):
        return ("nll_random_walk_" if self.walk else "nll_random_dnorm_") + self.nll_name


def g3l_random_dnorm(nll_name, param_f, mean_f=0, sigma_f=1, log_f=True, period="auto", nll_breakdown=False, weight=None):
    """R/likelihood_random.R:14-60: on every step of the period, nll += weight * -dnorm(param_f, mean_f, sigma_f, log)
    (the sign is +1 when log_f is FALSE). The Laplace machinery of a random effect stays in TMB; this is the term the
    objective function itself adds. weight defaults to the fixed parameter <nll_name>_weight = 1."""
    return RandomAction(False, nll_name, param_f, mean_f, sigma_f, log_f, period, nll_breakdown, weight)


def g3l_random_walk(nll_name, param_f, sigma_f=1, log_f=True, period="auto", nll_breakdown=False, weight=None):
    """R/likelihood_random.R:62-109: as g3l_random_dnorm with the mean = param_f on the previous step of the period;
    the first step only remembers the value."""
    return RandomAction(True, nll_name, param_f, None, sigma_f, log_f, period, nll_breakdown, weight)
