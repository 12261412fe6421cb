"""GPU box helper: time value+gradient batches with the CTA and the thread kernel."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gadget3_b200.configs import CONFIGS, parameter_batch
from gadget3_b200.run_cuda import g3_cuda_adfun
model, params = CONFIGS["C2"]()
fn = g3_cuda_adfun(model, params)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
P = parameter_batch(params, B, seed=3)
for k in (1, 3):
    fn.handle.set_kernel(k)
    fn.gr_batch(P[:64])
    t = time.time(); nll, g = fn.gr_batch(P); dt = time.time() - t
    print(f"kernel {k}: B={B} K={g.shape[1]} {dt*1e3:.1f} ms -> {B/dt:.0f} (value+gradient)/s, kernel ms {fn.handle.last_kernel_ms():.1f}")
