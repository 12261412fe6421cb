import sys
sys.path.insert(0, "/root/repo")
import numpy as np
from gadget3_b200.configs import CONFIGS, parameter_batch
from gadget3_b200.run_cuda import g3_cuda_adfun
for name in ("C1", "C2"):
    model, p = CONFIGS[name]()
    fn = g3_cuda_adfun(model, p)
    tidx = fn.opt_idx.astype(np.int32)
    for B in (1, 8, 64, 256, 512, 1024, 2048):
        full = fn.full_par(parameter_batch(p, B, seed=3))
        row = []
        for kernel in (1, 3, 0):
            fn.handle.set_kernel(kernel)
            try:
                fn.handle.eval_batch(full, tangent_idx=tidx)
                ms = []
                for _ in range(2):
                    fn.handle.eval_batch(full, tangent_idx=tidx); ms.append(fn.handle.last_kernel_ms())
                row.append(min(ms))
            except Exception as e:
                row.append(float("nan"))
            finally:
                fn.handle.set_kernel(0)
        print(f"{name} grad B={B:5d} work={B*((len(tidx)+3)//4):6d}  cta {row[0]:9.2f} ms  thread {row[1]:9.2f} ms  auto {row[2]:9.2f} ms")
