"""Read a TMB truth bundle written by tools/compare_with_tmb.R and rebuild the
same model through the host API, so g3cuda results can be compared with obj$fn / obj$gr / obj$report
of the reference's TMB backend (R/run_tmb.R:1254-1325) on a machine that has R.

The bundle stores the observation tables themselves, so no RNG has to agree between R and Python.
"""
from __future__ import annotations

import csv
import os

import numpy as np

from .configs import build_c1, build_c2


def _read_table(path):
    with open(path, newline="") as f:
        rows = list(csv.reader(f))
    head, body = rows[0], rows[1:]
    cols = {h: [r[i] for r in body] for i, h in enumerate(head)}
    for h in cols:
        if h in ("year", "step", "age"):
            cols[h] = [int(float(x)) for x in cols[h]]
        elif h in ("number", "weight"):
            cols[h] = [float(x) for x in cols[h]]
    return cols


def _num(x):
    return float("nan") if x in ("NA", "NaN", "") else float(x)


def _read_matrix(path):
    with open(path) as f:
        return np.array([[float(x) for x in ln.strip().split(",")] for ln in f if ln.strip()])


class TmbBundle:
    """model, parameters (template with the bundle's values / bounds / optimise flags), par [N][n_par],
    fn [N], gr [N][n_optimised], reports {i: {name: array}}."""

    def __init__(self, directory):
        d = directory
        self.kind = open(os.path.join(d, "model.txt")).read().strip()
        data = {k: _read_table(os.path.join(d, k + ".csv")) for k in ("landings", "ldist", "aldist", "si")}
        if self.kind == "multi":
            data["matp"] = _read_table(os.path.join(d, "matp.csv"))
        elif self.kind != "single":
            raise ValueError(f"unknown bundle model {self.kind!r}")
        self.model, p = (build_c2 if self.kind == "multi" else build_c1)(data=data)
        t = _read_table(os.path.join(d, "template.csv"))
        if list(t["switch"]) != list(p.switch):
            bad = [(a, b) for a, b in zip(t["switch"], p.switch) if a != b][:3]
            raise ValueError(f"Parameters not in expected order (first differences: {bad})")   # R/run_tmb.R:1156-1158
        for i in range(len(p.switch)):
            p.value[i] = float(t["value"][i])
            p.optimise[i] = bool(int(t["optimise"][i]))
            p.lower[i] = _num(t["lower"][i])        # NA: no bound
            p.upper[i] = _num(t["upper"][i])
        self.parameters = p
        self.par = _read_matrix(os.path.join(d, "par.csv"))
        self.fn = _read_matrix(os.path.join(d, "fn.csv")).ravel()
        self.gr = _read_matrix(os.path.join(d, "gr.csv"))
        self.reports = {}
        for i in range(len(self.par)):
            path = os.path.join(d, f"report_{i}.csv")
            if not os.path.exists(path):
                continue
            rows = _read_table(path)
            rep = {}
            for n, k, v in zip(rows["name"], rows["index"], rows["value"]):
                rep.setdefault(n, []).append(float(v))
            self.reports[i] = {n: np.array(v) for n, v in rep.items()}


def write_bundle(directory, kind, data, parameters, par, fn, gr, reports=None):
    """The Python twin of the R script's writer (used by the tests to exercise the reader)."""
    os.makedirs(directory, exist_ok=True)
    with open(os.path.join(directory, "model.txt"), "w") as f:
        f.write(kind + "\n")

    def table(name, cols):
        with open(os.path.join(directory, name + ".csv"), "w", newline="") as f:
            w = csv.writer(f, quoting=csv.QUOTE_NONNUMERIC)
            keys = list(cols)
            w.writerow(keys)
            for i in range(len(cols[keys[0]])):
                w.writerow([repr(float(cols[k][i])) if k in ("weight",) else cols[k][i] for k in keys])

    for k, v in data.items():
        table(k, v)
    p = parameters
    na = lambda x: "NA" if not np.isfinite(x) else repr(float(x))          # noqa: E731
    table("template", dict(switch=list(p.switch), value=[repr(float(v)) for v in p.values()],
                           optimise=[int(o) for o in p.optimise], lower=[na(x) for x in p.lower],
                           upper=[na(x) for x in p.upper]))
    for name, m in (("par", par), ("fn", np.asarray(fn).reshape(-1, 1)), ("gr", gr)):
        with open(os.path.join(directory, name + ".csv"), "w") as f:
            for row in np.atleast_2d(m):
                f.write(",".join(repr(float(x)) for x in row) + "\n")
    for i, rep in (reports or {}).items():
        with open(os.path.join(directory, f"report_{i}.csv"), "w") as f:
            f.write("name,index,value\n")
            for n, arr in rep.items():
                for k, v in enumerate(np.asarray(arr).ravel()):
                    f.write(f"{n},{k},{float(v)!r}\n")
