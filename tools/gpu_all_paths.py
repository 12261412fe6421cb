"""GPU box: one small pass of every kernel path of every config (value, gradient, report on the thread-, CTA-
and warp-per-evaluation kernels): a kernel that does not fit a model must answer with an error, never crash.
(compute-sanitizer is closed on the GPU pool this was developed on; where it is available, run this under
`compute-sanitizer --tool memcheck`.)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gadget3_b200.configs import CONFIGS, parameter_batch          # noqa: E402
from gadget3_b200.run_cuda import g3_cuda_adfun                    # noqa: E402

for name in sys.argv[1:] or ["C1", "C2", "MINI_ANDERSEN", "MINI_SUITS", "MINI_FLEETS", "LING"]:
    model, p = CONFIGS[name]()
    fn = g3_cuda_adfun(model, p, device=0)
    P = parameter_batch(p, 40, seed=1)
    for kernel in (3, 1, 2):
        try:
            fn.handle.set_kernel(kernel)
            nll = fn.fn_batch(P)
            _, g = fn.gr_batch(P[:2])
            rep = fn.report(P[0])
            print(name, "kernel", kernel, "nll[0]", nll[0], "grad ok", np.isfinite(g).mean(), "report keys", len(rep))
        except Exception as e:          # kernels that do not fit a model answer with an error code, not a crash
            print(name, "kernel", kernel, "refused:", str(e)[:80])
        finally:
            fn.handle.set_kernel(0)
