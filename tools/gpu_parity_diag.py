"""Parity numbers per config and kernel on a GPU box (what tests/test_gpu_parity.py asserts, printed instead of asserted).
Usage: python tools/gpu_parity_diag.py [CONFIG ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from _util import US_WEIGHT, consumed_biomass_per_step, us_tolerance  # noqa: E402
from gadget3_b200.configs import CONFIGS, parameter_batch  # noqa: E402
from gadget3_b200.run_cuda import g3_cuda_adfun  # noqa: E402
from oracle.oracle_lib import Oracle  # noqa: E402

names = sys.argv[1:] or ["C1", "C1_MULTI", "C2", "C4", "C5"]
for name in names:
    model, p = CONFIGS[name]()
    fn = g3_cuda_adfun(model, p, device=0)
    orc = Oracle(*model.pack(p))
    nvec = 63 if sum(L * A_ * R for L, A_, R in model.stock_shapes) > 400 else 255
    P = np.vstack([fn.par[None, :], parameter_batch(p, nvec, seed=2026)])
    full = fn.full_par(P)
    o_nll, o_comp, _ = orc.eval_batch(full, want_comp=True, nthreads=8)
    S = consumed_biomass_per_step(model, p)
    print(f"== {name}: nll median {np.median(o_nll):.4g}; us share of nll max {np.max(US_WEIGHT * o_comp[:, -2] / np.abs(o_nll)):.3g} "
          f"median {np.median(US_WEIGHT * o_comp[:, -2] / np.abs(o_nll)):.3g}; consumed biomass/step max {S.max():.4g}")
    for kernel in (4, 3):
        fn.handle.set_kernel(kernel)
        g_nll, g_comp, _ = fn.handle.eval_batch(full, want_comp=True)
        fn.handle.set_kernel(0)
        rel = np.abs(g_nll - o_nll) / np.abs(o_nll)
        ex = np.abs((g_nll - US_WEIGHT * g_comp[:, -2]) - (o_nll - US_WEIGHT * o_comp[:, -2])) / np.abs(o_nll)
        du = np.abs(g_comp[:, -2] - o_comp[:, -2])
        tol = us_tolerance(model, p, o_comp[:, -2])
        ncomp = len(model.comp_names)
        crel = np.abs(g_comp[:, :ncomp] - o_comp[:, :ncomp]) / np.maximum(np.abs(o_comp[:, :ncomp]), 1e-300)
        i = int(np.argmax(du / np.maximum(tol, 1e-300)))
        print(f"  kernel {kernel}: nll rel max {rel.max():.3g} (n>1e-10: {(rel > 1e-10).sum()}), excl. understocking max {ex.max():.3g} "
              f"median {np.median(ex):.3g}; comps rel max {crel.max():.3g}; us |diff|/tol max {np.max(du / np.maximum(tol, 1e-300)):.3g} "
              f"at vec {i}: us {o_comp[i, -2]:.6g} diff {du[i]:.3g} tol {tol[i]:.3g} sqrt(u) {np.sqrt(o_comp[i, -2]):.3g}")
