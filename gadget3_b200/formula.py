"""Parameter formulas: the host-side mirror of gadget3's R formulas for scalar slots.

In gadget3 every numeric argument of an action is an R formula that may mention
``g3_param()``/``g3_param_table()`` calls and loop variables (``age``, ``area``,
``cur_year``, ``cur_step``, ``cur_step_size``) -- see R/params.R:43-223 and
R/run_tmb.R:764-883. Here such a formula is an :class:`Expr` tree. The lowering step
(:class:`Lowerer`) instantiates it for concrete loop-variable values, resolves
``g3_param_table`` lookups to a parameter index, folds constants, and emits a tiny postfix
program ("slot") that the device evaluates once per parameter vector.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np

from .spec import K


class G3Unsupported(Exception):
    """Raised when a model uses something outside the supported action subset
    (mirrors the ``stop()`` a ``g3_to_cuda()`` R function would raise: no CPU fallback)."""


# --------------------------------------------------------------------------- expressions
class Expr:
    def __add__(self, o): return BinOp("add", self, wrap(o))
    def __radd__(self, o): return BinOp("add", wrap(o), self)
    def __sub__(self, o): return BinOp("sub", self, wrap(o))
    def __rsub__(self, o): return BinOp("sub", wrap(o), self)
    def __mul__(self, o): return BinOp("mul", self, wrap(o))
    def __rmul__(self, o): return BinOp("mul", wrap(o), self)
    def __truediv__(self, o): return BinOp("div", self, wrap(o))
    def __rtruediv__(self, o): return BinOp("div", wrap(o), self)
    def __pow__(self, o): return BinOp("pow", self, wrap(o))
    def __rpow__(self, o): return BinOp("pow", wrap(o), self)
    def __neg__(self): return UnOp("neg", self)


def wrap(x) -> Expr:
    if isinstance(x, Expr):
        return x
    if isinstance(x, (int, float, np.integer, np.floating)):
        return Const(float(x))
    raise TypeError(f"cannot use {x!r} in a g3 formula")


class Const(Expr):
    def __init__(self, v):
        self.v = float(v)


class Var(Expr):
    """Loop variable: age, area, cur_year, cur_step, cur_step_size."""
    def __init__(self, name):
        self.name = name


class BinOp(Expr):
    def __init__(self, op, a, b):
        self.op, self.a, self.b = op, a, b


class UnOp(Expr):
    def __init__(self, op, a):
        self.op, self.a = op, a


class Lookup(Expr):
    """Data-vector lookup ``values[[key - base + 1]]`` (e.g. ``ling_imm_stddev[[age - 3 + 2]]``,
    inttest/codegeneration/run.R:49)."""
    def __init__(self, values, key: Expr, base: float):
        self.values, self.key, self.base = list(values), wrap(key), base


class TimeVar(Expr):
    """g3_timevariable (R/timevariable.R:1-31): the value of `init` until the first (year, step) listed, then each entry's
    formula from its (year, step) on."""
    def __init__(self, name, init, entries):
        self.name, self.init, self.entries = name, init, entries       # entries: [(year, step, Expr)] in time order


class ParamExpr(Expr):
    """Result of :func:`g3_parameterized` before scale/offset wrapping: one ``g3_param`` or
    ``g3_param_table`` call (R/params.R:170-186)."""
    def __init__(self, name, by_stock, by_predator, by_year, by_step, by_age, by_area,
                 ifmissing, value, lower, upper, optimise, prepend_extra=()):
        self.name = name
        self.by_stock, self.by_predator = by_stock, by_predator
        self.by_year, self.by_step, self.by_age, self.by_area = by_year, by_step, by_age, by_area
        self.ifmissing = ifmissing
        self.value, self.lower, self.upper, self.optimise = value, lower, upper, optimise
        self.prepend_extra = tuple(prepend_extra)


def g3_exp(x): return UnOp("exp", wrap(x))
def g3_log(x): return UnOp("log", wrap(x))
def g3_sqrt(x): return UnOp("sqrt", wrap(x))
def g3_avoid_zero(x): return UnOp("avoid_zero", wrap(x))
def g3_as_integer(x): return UnOp("floor", wrap(x))     # as_integer() of R/aab_env.R:9 (loop variables only)


age = Var("age")
area = Var("area")
cur_year = Var("cur_year")
cur_step = Var("cur_step")
cur_step_size = Var("cur_step_size")


def g3_timevariable(lookup_name, fs) -> Expr:
    """Mirror of ``g3_timevariable()`` (R/timevariable.R:1-31). ``fs``: an ordered mapping; the first key must be 'init', the
    others 'YYYY' (step 1) or 'YYYY-SS'."""
    import re as _re
    items = list(fs.items())
    if not items or items[0][0] != "init":
        raise ValueError("First formula should be named 'init'")
    entries = []
    for k, f in items[1:]:
        m = _re.match(r"^([0-9]{4})-?([0-9]{1,2})?$", str(k))
        if not m:
            raise ValueError(f"Invalid formula identifier(s): {k}")
        entries.append((int(m.group(1)), int(m.group(2)) if m.group(2) else 1, wrap(f)))
    return TimeVar("tv_" + lookup_name, wrap(items[0][1]), entries)


def g3_parameterized(name, by_stock=False, by_predator=False, by_year=False, by_step=False,
                     by_age=False, by_area=False, prepend_extra=(), exponentiate=False,
                     avoid_zero=False, scale=1, offset=0, ifmissing=None,
                     value=0.0, lower=float("nan"), upper=float("nan"), optimise=None) -> Expr:
    """Mirror of ``g3_parameterized()`` (R/params.R:43-223)."""
    if isinstance(name, (list, tuple)):
        name = ".".join(name)
    if by_stock is False and by_age:
        raise ValueError("!by_stock && by_age doesn't make sense")
    if exponentiate:
        name = name + "_exp"
    if isinstance(ifmissing, str):
        ifmissing = g3_parameterized(ifmissing, by_stock=by_stock)
    if optimise is None:
        optimise = bool(np.isfinite(lower) and np.isfinite(upper))
    out: Expr = ParamExpr(name, by_stock, by_predator, by_year, by_step, by_age, by_area,
                          ifmissing, value, lower, upper, optimise, prepend_extra)
    if isinstance(scale, str):
        scale = g3_parameterized(scale, by_stock=by_stock)
    if isinstance(offset, str):
        offset = g3_parameterized(offset, by_stock=by_stock)
    if exponentiate:
        out = g3_exp(out)
    if isinstance(scale, Expr) or scale != 1:
        out = out * scale
    if avoid_zero:
        out = g3_avoid_zero(out)
    if isinstance(offset, Expr) or offset != 0:
        out = out + offset
    return out


# --------------------------------------------------------------------------- parameter template
class ParameterTemplate:
    """Mirror of the ``parameter_template`` data.frame attribute (R/run_tmb.R:789-803):
    columns switch, value, optimise, random, lower, upper, parscale, in PARAMETER order."""

    COLUMNS = ("switch", "value", "optimise", "random", "lower", "upper", "parscale")

    def __init__(self):
        self.switch, self.value, self.optimise = [], [], []
        self.lower, self.upper, self.parscale, self.random = [], [], [], []
        self._index = {}

    def __len__(self):
        return len(self.switch)

    def index(self, name) -> int:
        return self._index[name]

    def __contains__(self, name):
        return name in self._index

    def add(self, name, value=0.0, lower=float("nan"), upper=float("nan"), optimise=False) -> int:
        if name in self._index:
            return self._index[name]
        self._index[name] = len(self.switch)
        self.switch.append(name)
        self.value.append(float(value))
        self.lower.append(float(lower))
        self.upper.append(float(upper))
        self.optimise.append(bool(optimise))
        self.parscale.append(float("nan"))
        self.random.append(False)
        return self._index[name]

    def copy(self) -> "ParameterTemplate":
        out = ParameterTemplate()
        for c in ("switch", "value", "optimise", "lower", "upper", "parscale", "random"):
            setattr(out, c, list(getattr(self, c)))
        out._index = dict(self._index)
        return out

    def values(self) -> np.ndarray:
        return np.asarray(self.value, dtype=np.float64)

    def optimised_indices(self) -> np.ndarray:
        return np.nonzero(np.asarray(self.optimise, dtype=bool))[0].astype(np.int32)

    def as_dict(self) -> dict:
        return {c: list(getattr(self, c)) for c in self.COLUMNS}


def _name_re(name_spec: str):
    """Wildcard syntax of ``g3_init_val`` (R/init_val.R:23-51)."""
    import re
    parts = []
    for part in name_spec.split("."):
        m = re.match(r"^\[(\d+)[:-](\d+)\]$", part)
        if m:
            parts.append("(?:" + "|".join(str(i) for i in range(int(m.group(1)), int(m.group(2)) + 1)) + ")")
            continue
        rx, buf = "", ""
        for ch in part:
            if ch in "#*|":
                rx += re.escape(buf)
                buf = ""
                rx += {"#": r"\d+", "*": r".*", "|": "|"}[ch]
            else:
                buf += ch
        rx += re.escape(buf)
        parts.append("(?:" + rx + ")")
    return re.compile("^" + r"\.".join(parts) + "(_exp)?$")


def g3_init_val(param_template: ParameterTemplate, name_spec: str, value=None, spread=None,
                lower=None, upper=None, optimise=None, parscale=None) -> ParameterTemplate:
    """Mirror of ``g3_init_val()`` (R/init_val.R:1-132) for the data.frame form."""
    if spread is not None:
        if lower is None:
            lower = min(value * (1 - spread), value * (1 + spread))
        if upper is None:
            upper = max(value * (1 - spread), value * (1 + spread))
    if optimise is None:
        optimise = (lower is not None) and (upper is not None)
    out = param_template.copy()
    rx = _name_re(name_spec)
    hit = False
    for i, n in enumerate(out.switch):
        m = rx.match(n)
        if not m:
            continue
        hit = True
        is_exp = m.group(1) == "_exp"
        f = (lambda v: math.log(v)) if is_exp else (lambda v: v)
        if value is not None:
            out.value[i] = f(float(value))
        if lower is not None:
            out.lower[i] = f(float(lower))
        if upper is not None:
            out.upper[i] = f(float(upper))
        out.optimise[i] = bool(optimise)
        if lower is not None and upper is not None:
            out.parscale[i] = out.upper[i] - out.lower[i]
        elif parscale is not None:
            out.parscale[i] = float(parscale)
    if not hit:
        import warnings
        warnings.warn(f"g3_init_val('{name_spec}') didn't match any parameters")
    return out


# --------------------------------------------------------------------------- lowering
_BIN = {"add": K["G3OP_ADD"], "sub": K["G3OP_SUB"], "mul": K["G3OP_MUL"], "div": K["G3OP_DIV"],
        "pow": K["G3OP_POW"]}
_UN = {"neg": K["G3OP_NEG"], "exp": K["G3OP_EXP"], "log": K["G3OP_LOG"], "sqrt": K["G3OP_SQRT"],
       "avoid_zero": K["G3OP_AVOID_ZERO"]}


def _avoid_zero(x: float) -> float:
    p = x * 1000.0
    return (max(p, 0.0) + math.log1p(math.exp(-abs(p)))) / 1000.0


class Ctx:
    """Values of the loop variables a formula is instantiated for."""
    def __init__(self, stock=None, predstock=None, age=None, area=None, area_name=None,
                 cur_year=None, cur_step=None, cur_step_size=None):
        self.stock, self.predstock = stock, predstock
        self.age, self.area, self.area_name = age, area, area_name
        self.cur_year, self.cur_step, self.cur_step_size = cur_year, cur_step, cur_step_size

    def replace(self, **kw) -> "Ctx":
        out = Ctx(**self.__dict__)
        for k, v in kw.items():
            setattr(out, k, v)
        return out


class Lowerer:
    """Turns (Expr, Ctx) into slot ids; owns the parameter template being discovered."""

    def __init__(self, builder, template: ParameterTemplate, start_year, end_year, n_steps):
        self.b = builder
        self.template = template
        self.start_year, self.end_year, self.n_steps = start_year, end_year, n_steps
        self.slots = []          # list of code tuples
        self._slot_index = {}
        self._tables_done = set()

    # -- parameter names ----------------------------------------------------
    @staticmethod
    def _stock_part(stock, by):
        if by is True:
            return stock.name
        if isinstance(by, str):
            if by not in stock.name_parts:
                raise ValueError(f"stock {stock.name} has no name part {by!r}")
            return stock.name_parts[by]
        # explicit stock / list of stocks: common part of the names
        stocks = by if isinstance(by, (list, tuple)) else [by]
        parts = [s.name.split("_") for s in stocks]
        common = [p for p in parts[0] if all(p in q for q in parts[1:])]
        return "_".join(common)

    def _base_name(self, p: ParamExpr, ctx: Ctx) -> str:
        parts = []
        for ex in reversed(p.prepend_extra):
            parts.append(str(ex))
        if p.by_stock is not False:
            if ctx.stock is None and (p.by_stock is True or isinstance(p.by_stock, str)):    # explicit stocks name themselves
                raise G3Unsupported(f"parameter {p.name}: by_stock used outside a stock context")
            parts.append(self._stock_part(ctx.stock, p.by_stock))
        if p.by_predator is not False:
            if ctx.predstock is None:
                raise G3Unsupported(f"parameter {p.name}: by_predator used outside a predator context")
            parts.append(self._stock_part(ctx.predstock, p.by_predator))
        parts.append(p.name)
        return ".".join(parts)

    def _table_keys(self, p: ParamExpr, ctx: Ctx):
        """All (suffix tuple) rows of the g3_param_table, expand.grid order (first varies fastest)."""
        dims = []
        if p.by_year is True:
            dims.append(("cur_year", list(range(self.start_year, self.end_year + 1))))
        elif p.by_year is not False:
            dims.append(("cur_year", [int(y) for y in p.by_year]))
        if p.by_step:
            dims.append(("cur_step", list(range(1, self.n_steps + 1))))
        if p.by_age:
            s = ctx.stock
            if isinstance(p.by_stock, (list, tuple)):
                lo = min(x.minage for x in p.by_stock)
                hi = max(x.maxage for x in p.by_stock)
            else:
                lo, hi = s.minage, s.maxage
            dims.append(("age", list(range(lo, hi + 1))))
        if p.by_area:
            owner = ctx.stock if p.by_stock is not False else ctx.predstock
            dims.append(("area", list(owner.areas.keys())))
        return dims

    def _register_table(self, p: ParamExpr, base: str, ctx: Ctx):
        dims = self._table_keys(p, ctx)
        if base in self._tables_done:
            return dims
        self._tables_done.add(base)
        # expand.grid: first dimension varies fastest
        idx = [0] * len(dims)
        total = 1
        for _, vals in dims:
            total *= len(vals)
        for _ in range(total):
            suffix = ".".join(str(dims[d][1][idx[d]]) for d in range(len(dims)))
            self.template.add(base + "." + suffix, p.value, p.lower, p.upper, p.optimise)
            for d in range(len(dims)):
                idx[d] += 1
                if idx[d] < len(dims[d][1]):
                    break
                idx[d] = 0
        return dims

    # -- tree -> (is_const, value | code) ------------------------------------
    def _var(self, name, ctx: Ctx) -> float:
        v = getattr(ctx, name)
        if v is None:
            raise G3Unsupported(f"formula depends on '{name}', which is not available for this slot")
        return float(v)

    def _lower(self, e: Expr, ctx: Ctx):
        """Returns ('c', float) or ('x', [code ints])."""
        if isinstance(e, Const):
            return ("c", e.v)
        if isinstance(e, Var):
            return ("c", self._var(e.name, ctx))
        if isinstance(e, Lookup):
            k = self._lower(e.key, ctx)
            if k[0] != "c":
                raise G3Unsupported("data lookup keyed by a parameter")
            i = int(math.floor(k[1] - e.base + 0.5))
            if not (0 <= i < len(e.values)):
                raise G3Unsupported(f"data lookup index {i} out of range")
            return ("c", float(e.values[i]))
        if isinstance(e, TimeVar):
            # every formula is lowered (its parameters are declared whatever the step), the step picks one: the global formula
            # is re-evaluated every step and keeps its last value (R/timevariable.R:20-29)
            now = (int(self._var("cur_year", ctx)), int(self._var("cur_step", ctx)))
            out = self._lower(e.init, ctx)
            for y, st, f in e.entries:
                v = self._lower(f, ctx)
                if (y, st) <= now:
                    out = v
            return out
        if isinstance(e, ParamExpr):
            base = self._base_name(e, ctx)
            is_table = (e.by_year is not False) or e.by_step or e.by_age or e.by_area
            if not is_table:
                return ("x", [K["G3OP_PARAM"], self.template.add(base, e.value, e.lower, e.upper, e.optimise)])
            # ifmissing is declared before the table rows (R/run_tmb.R:808-812)
            miss = self._lower(e.ifmissing, ctx) if e.ifmissing is not None else None
            dims = self._register_table(e, base, ctx)
            suffix = []
            for dname, vals in dims:
                if dname == "area":
                    key = ctx.area_name
                    if key is None:
                        raise G3Unsupported(f"parameter {base}: by_area outside an area loop")
                else:
                    key = int(math.floor(self._var(dname, ctx) + 0.5))
                if key not in vals:
                    if miss is not None:
                        return miss
                    return ("x", [K["G3OP_NAN"]])
                suffix.append(str(key))
            return ("x", [K["G3OP_PARAM"], self.template.index(base + "." + ".".join(suffix))])
        if isinstance(e, UnOp):
            a = self._lower(e.a, ctx)
            if a[0] == "c":
                v = a[1]
                f = {"neg": lambda x: -x, "exp": math.exp, "log": math.log, "sqrt": math.sqrt,
                     "avoid_zero": _avoid_zero, "floor": math.floor}[e.op]
                try:
                    return ("c", f(v))
                except (ValueError, OverflowError):
                    return ("c", float("nan"))
            if e.op == "floor":
                raise G3Unsupported("as_integer() of a parameter-dependent value")
            return ("x", a[1] + [_UN[e.op]])
        if isinstance(e, BinOp):
            a, b = self._lower(e.a, ctx), self._lower(e.b, ctx)
            if a[0] == "c" and b[0] == "c":
                x, y = a[1], b[1]
                try:
                    v = {"add": lambda: x + y, "sub": lambda: x - y, "mul": lambda: x * y,
                         "div": lambda: x / y if y != 0 else math.copysign(math.inf, x) if x != 0 else math.nan,
                         "pow": lambda: math.pow(x, y)}[e.op]()
                except (ValueError, OverflowError):
                    v = float("nan")
                return ("c", v)
            return ("x", self._code(a) + self._code(b) + [_BIN[e.op]])
        raise TypeError(f"not a g3 formula: {e!r}")

    def _code(self, r):
        if r[0] == "c":
            return [K["G3OP_CONST"], self.b.const(r[1])]
        return r[1]

    @staticmethod
    def _depth(code) -> int:
        d = m = 0
        i = 0
        while i < len(code):
            op = code[i]
            if op in (K["G3OP_CONST"], K["G3OP_PARAM"]):
                d += 1
                i += 2
            elif op == K["G3OP_NAN"]:
                d += 1
                i += 1
            elif op in _BIN.values():
                d -= 1
                i += 1
            else:
                i += 1
            m = max(m, d)
        return m

    def slot(self, e, ctx: Ctx) -> int:
        """Slot id holding the value of formula ``e`` instantiated for ``ctx``."""
        code = tuple(self._code(self._lower(wrap(e), ctx)))
        if self._depth(code) > K["G3_SLOT_STACK"]:
            raise G3Unsupported("formula too deep for the slot evaluator")
        if code not in self._slot_index:
            self._slot_index[code] = len(self.slots)
            self.slots.append(code)
        return self._slot_index[code]

    def const_value(self, e, ctx: Ctx) -> Optional[float]:
        r = self._lower(wrap(e), ctx)
        return r[1] if r[0] == "c" else None

    def finish(self):
        """Write the slot table into the builder; returns (n_slot, slot_off)."""
        table = []
        for code in self.slots:
            off = self.b.alloc_ints(code)
            table += [off, len(code)]
        return len(self.slots), self.b.alloc_ints(table)
