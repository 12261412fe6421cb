import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
from gadget3_b200.configs import CONFIGS
from gadget3_b200.run_cuda import g3_cuda_adfun
for name in ("C1", "C2", "C5"):
    model, p = CONFIGS[name]()
    obj = g3_cuda_adfun(model, p)
    for what, f in (("fn", lambda: obj.fn(obj.par)), ("gr", lambda: obj.gr(obj.par)), ("report", lambda: obj.report(obj.par))):
        f()
        t = time.perf_counter()
        for _ in range(5): f()
        print(name, what, "%.2f ms per call (wall, host buffers)" % ((time.perf_counter() - t) / 5 * 1e3), "n_opt", len(obj.par))
