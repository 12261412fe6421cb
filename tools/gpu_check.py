"""Debug helper (GPU box): compare the CUDA path with the oracle step by step on C1/C2."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gadget3_b200.configs import CONFIGS, parameter_batch
from gadget3_b200.run_cuda import g3_cuda_adfun, unpack_report
from oracle.oracle_lib import Oracle


def main():
    names = sys.argv[1:] or ["C1", "C2"]
    for name in names:
        model, params = CONFIGS[name]()
        fn = g3_cuda_adfun(model, params)
        ints, reals = model.pack(params)
        orc = Oracle(ints, reals)
        P = np.vstack([fn.par[None, :], parameter_batch(params, 255)])
        full = fn.full_par(P)
        t = time.time(); o_nll, o_comp, _ = orc.eval_batch(full, want_comp=True, nthreads=8); to = time.time() - t
        t = time.time(); g_nll, g_comp, _ = fn.handle.eval_batch(full, want_comp=True); tg = time.time() - t
        rel = np.abs(g_nll - o_nll) / np.maximum(np.abs(o_nll), 1e-300)
        print(f"{name}: oracle {to*1e3:.1f} ms, gpu {tg*1e3:.1f} ms (kernel {fn.handle.last_kernel_ms():.3f} ms), max rel err nll {rel.max():.3e}")
        print("  nll[:4] gpu", g_nll[:4], "oracle", o_nll[:4])
        wo_g, wo_o = g_nll - 1e8 * g_comp[:, -2], o_nll - 1e8 * o_comp[:, -2]
        print("  max rel err nll without understocking", (np.abs(wo_g - wo_o) / np.abs(o_nll)).max(),
              " us share of nll at worst vector", (1e8 * o_comp[:, -2] / o_nll)[int(np.argmax(rel))])
        crel = np.abs(g_comp - o_comp) / np.maximum(np.abs(o_comp), 1e-12)
        print("  comp max rel err per row", crel.max(axis=0))
        if rel.max() > 1e-10:
            i = int(np.argmax(rel))
            ro = unpack_report(model, orc.report(full[i]))
            rg = unpack_report(model, fn.handle.report(full[i]))
            for k in ro:
                a, b = np.asarray(ro[k]), np.asarray(rg[k])
                d = np.abs(a - b) / np.maximum(np.abs(a), 1e-12)
                if np.nanmax(d) > 1e-10 or np.isnan(a).any() != np.isnan(b).any():
                    idx = np.unravel_index(np.nanargmax(d), d.shape) if d.ndim else ()
                    print(f"  MISMATCH {k}: max rel {np.nanmax(d):.3e} at {idx}: oracle {a[idx] if d.ndim else a} gpu {b[idx] if d.ndim else b}")
                    if k.startswith("dstart") and d.ndim == 4:
                        bad_t = np.where(np.nanmax(d, axis=(0, 1, 2)) > 1e-10)[0]
                        print("    first bad time", bad_t[:5])
        # gradient
        idx = fn.opt_idx[:6]
        _, _, og = orc.eval_batch(full[:2], tangent_idx=idx)
        _, _, gg = fn.handle.eval_batch(full[:2], tangent_idx=idx)
        print("  grad oracle", og[0], "\n  grad gpu   ", gg[0])
        B = 4096
        Pb = fn.full_par(parameter_batch(params, B, seed=7))
        fn.handle.eval_batch(Pb)
        t = time.time(); fn.handle.eval_batch(Pb); tg = time.time() - t
        ms = fn.handle.last_kernel_ms()
        print(f"  B={B}: e2e {tg*1e3:.1f} ms, kernel {ms:.2f} ms -> {B/ms*1e3:.0f} evals/s")


if __name__ == "__main__":
    main()
