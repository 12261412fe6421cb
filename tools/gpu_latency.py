"""GPU box: kernel time of small batches on the CTA- and the thread-per-evaluation kernels (dispatch threshold)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gadget3_b200.configs import CONFIGS, parameter_batch          # noqa: E402
from gadget3_b200.run_cuda import g3_cuda_adfun                    # noqa: E402

for name in sys.argv[1:] or ["C1", "C2"]:
    model, p = CONFIGS[name]()
    fn = g3_cuda_adfun(model, p, device=0)
    for B in (1, 32, 256, 1024, 2048, 4096, 8192, 16384, 32768):
        full = fn.full_par(parameter_batch(p, B, seed=3))
        row = []
        for kernel in (1, 3, 0):
            fn.handle.set_kernel(kernel)
            try:
                fn.handle.eval_batch(full)
                ms = []
                for _ in range(3):
                    fn.handle.eval_batch(full)
                    ms.append(fn.handle.last_kernel_ms())
                row.append(min(ms))
            except Exception:
                row.append(float("nan"))
            finally:
                fn.handle.set_kernel(0)
        print(f"{name} B={B:6d}  cta {row[0]:9.3f} ms  thread {row[1]:9.3f} ms  auto {row[2]:9.3f} ms")
