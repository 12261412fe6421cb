"""Packed model specification ("the IR") shared by the CUDA library and the CPU oracle.

The layout constants live in ``include/g3cuda_spec.h`` (the single source of truth);
this module parses the ``#define``s so Python, CUDA and the oracle can never drift.

Reference: replaces the generated TMB translation unit + ``model_data`` environment that
``g3_to_tmb()`` returns (R/run_tmb.R:1076-1091).
"""
from __future__ import annotations

import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "g3cuda_spec.h")


def _parse_header(path: str) -> dict:
    out = {}
    pat = re.compile(r"^#define\s+(G3\w+)\s+(-?\d+)\b")
    with open(path) as fh:
        for line in fh:
            m = pat.match(line)
            if m:
                out[m.group(1)] = int(m.group(2))
    return out


K = _parse_header(HEADER_PATH)
globals().update(K)  # G3H_*, G3S_*, ... available as module attributes


class SpecBuilder:
    """Append-only builder for the ints[] / reals[] arrays."""

    def __init__(self):
        self.ints = [0] * K["G3H_LEN"]
        self.reals = []
        self._const_cache = {}

    # -- ints ---------------------------------------------------------------
    def alloc_ints(self, values) -> int:
        off = len(self.ints)
        self.ints.extend(int(v) for v in values)
        return off

    def set_int(self, off: int, v: int):
        self.ints[off] = int(v)

    # -- reals --------------------------------------------------------------
    def alloc_reals(self, values) -> int:
        off = len(self.reals)
        self.reals.extend(float(v) for v in np.asarray(values, dtype=np.float64).ravel())
        return off

    def const(self, v: float) -> int:
        """Offset of a (deduplicated) scalar constant in reals[]."""
        key = np.float64(v).tobytes()
        if key not in self._const_cache:
            self._const_cache[key] = self.alloc_reals([v])
        return self._const_cache[key]

    def finish(self):
        self.ints[K["G3H_VERSION"]] = K["G3CUDA_SPEC_VERSION"]
        return (np.asarray(self.ints, dtype=np.int32), np.asarray(self.reals, dtype=np.float64))
