"""Stock definitions: host-side mirror of g3_stock / g3s_* (R/stock.R, R/stock_age.R, R/stock_areas.R)."""
from __future__ import annotations

import copy
from collections import OrderedDict

import numpy as np


def g3_areas(area_names):
    """R/stock_areas.R:1-3: named vector name -> 1..n."""
    return OrderedDict((n, i + 1) for i, n in enumerate(area_names))


class Stock:
    def __init__(self, name, name_parts):
        self.name = name
        self.name_parts = dict(name_parts)
        self.dims = []            # order in which g3s_* were applied, e.g. ['length', 'age', 'area']
        self.lengthgroups = None
        self.minage = self.maxage = None
        self.areas = None         # OrderedDict area name -> model area id

    def _clone(self):
        return copy.deepcopy(self)

    @property
    def L(self):
        return len(self.minlen)

    @property
    def nage(self):
        return self.maxage - self.minage + 1

    def __repr__(self):
        return f"<g3_stock {self.name} dims={self.dims}>"


def _stat_mode(vec):
    # rle()-based mode of R/stock.R:78-81: first longest run
    best, best_len = vec[0], 0
    i = 0
    while i < len(vec):
        j = i
        while j < len(vec) and vec[j] == vec[i]:
            j += 1
        if j - i > best_len:
            best, best_len = vec[i], j - i
        i = j
    return best


def g3s_length(stock: Stock, lengthgroups, open_ended=True, plus_dl=None) -> Stock:
    """R/stock.R:68-153."""
    s = stock._clone()
    lg = [float(x) for x in lengthgroups]
    if plus_dl is not None:
        plusdl = float(plus_dl)
    elif len(lg) == 1:
        plusdl = 1.0
    else:
        plusdl = float(_stat_mode(list(np.diff(lg))))
    if open_ended:
        dl = np.diff(lg + [lg[-1] + plusdl])
        upperlen = float("inf")
    else:
        dl = np.diff(lg)
        upperlen = lg[-1]
        lg = lg[:-1]
    s.lengthgroups = lg
    s.open_ended = bool(open_ended)
    s.plusdl = plusdl
    s.dl = np.asarray(dl, dtype=np.float64)
    s.upperlen = upperlen
    s.minlen = np.asarray(lg, dtype=np.float64)
    s.midlen = s.minlen + s.dl / 2
    s.maxlen = np.asarray(lg[1:] + [upperlen], dtype=np.float64)
    if "length" not in s.dims:
        s.dims.append("length")
    return s


def g3_stock(var_name, lengthgroups, open_ended=True) -> Stock:
    """R/stock.R:156: var_name may be a dict/list of name parts, e.g. {'species': 'fish', '': 'imm'}."""
    if isinstance(var_name, dict):
        parts = var_name
        name = "_".join(str(v) for v in var_name.values())
    elif isinstance(var_name, (list, tuple)):
        parts = {str(i): v for i, v in enumerate(var_name)}
        name = "_".join(var_name)
    else:
        parts, name = {}, str(var_name)
    parts = dict(parts)
    parts["full"] = name
    return g3s_length(Stock(name, parts), lengthgroups, open_ended)


def g3_fleet(var_name) -> Stock:
    """R/stock.R:155: a fleet is a stock without length/age dimensions."""
    if isinstance(var_name, (list, tuple)):
        name = "_".join(var_name)
    else:
        name = str(var_name)
    return Stock(name, {"full": name})


def g3s_age(stock: Stock, minage, maxage) -> Stock:
    """R/stock_age.R:1-50."""
    s = stock._clone()
    s.minage, s.maxage = int(minage), int(maxage)
    if "age" not in s.dims:
        s.dims.append("age")
    return s


def g3s_livesonareas(stock: Stock, areas) -> Stock:
    """R/stock_areas.R:5-100. ``areas``: dict name -> id (a slice of g3_areas())."""
    s = stock._clone()
    if not isinstance(areas, dict):
        areas = OrderedDict((f"area{a}", int(a)) for a in areas)
    s.areas = OrderedDict((str(k), int(v)) for k, v in areas.items())
    if "area" not in s.dims:
        s.dims.append("area")
    return s


def g3_stock_def(stock: Stock, name):
    """R/stock.R:43-58."""
    return getattr(stock, name)
