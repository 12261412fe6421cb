#!/usr/bin/env python
"""Replay a TMB truth bundle (tools/compare_with_tmb.R) through the C ABI and compare.

    python tools/replay_tmb_bundle.py <bundle_dir> [--device 0]

nll is compared at 1e-10 relative, gradients at 1e-8 of the largest entry, per-step nll_* arrays
at 1e-9. Exit status 0 = every vector within tolerance.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from gadget3_b200.run_cuda import g3_cuda_adfun, unpack_report      # noqa: E402
from gadget3_b200.tmb_bundle import TmbBundle                       # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("bundle")
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args()
    b = TmbBundle(a.bundle)
    fn = g3_cuda_adfun(b.model, b.parameters, device=a.device)
    nll, comp, _ = fn.handle.eval_batch(b.par, want_comp=True)
    _, _, grad = fn.handle.eval_batch(b.par, tangent_idx=fn.opt_idx.astype(np.int32))
    rel = np.abs(nll - b.fn) / np.abs(b.fn)
    gscale = np.abs(b.gr).max(axis=1, keepdims=True)
    grel = (np.abs(grad - b.gr) / gscale).max(axis=1)
    print(f"{len(nll)} vectors: nll max rel diff {rel.max():.3e} (median {np.median(rel):.3e}); "
          f"gradient max rel diff {grel.max():.3e}")
    ok = bool((rel < 1e-10).all() and (grel < 1e-8).all())
    for i, want in b.reports.items():
        rep = unpack_report(b.model, fn.handle.report(b.par[i]))
        for name, arr in want.items():
            if name not in rep or name.endswith("__weight"):
                continue
            mine = np.asarray(rep[name]).ravel()
            err = np.abs(mine - arr).max() / max(np.abs(arr).max(), 1e-300)
            print(f"  vector {i} {name}: max rel diff {err:.3e}")
            ok = ok and err < 1e-9
    print("PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
