"""(value + gradient) evaluations per second of both kernels per config (host buffers, wall clock)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gadget3_b200.configs import CONFIGS, parameter_batch
from gadget3_b200.run_cuda import g3_cuda_adfun
for name, B in (("C1", 8192), ("C2", 4096), ("C4", 512), ("C5", 2048)):
    model, p = CONFIGS[name]()
    fn = g3_cuda_adfun(model, p, device=0)
    full = fn.full_par(parameter_batch(p, B, seed=5))
    tix = fn.opt_idx.astype(np.int32)
    res = {}
    for k in (3, 4):
        try:
            fn.handle.set_kernel(k)
            fn.handle.eval_batch(full[:256], tangent_idx=tix)
            t = time.perf_counter(); out = fn.handle.eval_batch(full, tangent_idx=tix); dt = time.perf_counter() - t
            res[k] = out
            print(f"{name} kernel {k}: {B} vectors x {len(tix)} directions in {dt:.3f} s = {B / dt:,.0f} (value+gradient)/s", flush=True)
        except Exception as e:
            print(name, k, "failed:", e)
    if 3 in res and 4 in res:
        sc = np.abs(res[3][2]).max(axis=1, keepdims=True)
        print(f"   kernels agree: nll {np.max(np.abs(res[3][0] - res[4][0]) / np.abs(res[3][0])):.2g}, grad {np.max(np.abs(res[3][2] - res[4][2]) / sc):.2g}")
