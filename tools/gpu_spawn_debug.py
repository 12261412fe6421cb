"""Thread-kernel report vs oracle report on the spawning rigs: first array / step that differs."""
import sys
import numpy as np
sys.path.insert(0, ".")
from gadget3_b200.configs import CONFIGS
from gadget3_b200.run_cuda import g3_cuda_adfun, unpack_report
from oracle.oracle_lib import Oracle, build

build()
for name in sys.argv[1:] or ["SPAWN_RIG", "MINI_SPAWN"]:
    model, p = CONFIGS[name]()
    fn = g3_cuda_adfun(model, p, device=0)
    orc = Oracle(*model.pack(p))
    rg, ro = fn.report(fn.par), unpack_report(model, orc.report(fn.full_par(fn.par[None, :])[0]))
    print("==", name, "nll", float(np.asarray(rg["nll"])), float(np.asarray(ro["nll"])))
    for k in ro:
        a, b = np.atleast_1d(np.asarray(rg[k], dtype=float)), np.atleast_1d(np.asarray(ro[k], dtype=float))
        if a.shape != b.shape:
            print(k, "shape", a.shape, b.shape); continue
        err = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
        err[(a == b)] = 0
        if err.size and err.max() > 1e-9:
            if k.startswith("dstart") or k.startswith("detail"):
                bt = err.reshape(-1, err.shape[-1]).max(axis=0)
                first = int(np.argmax(bt > 1e-9))
                print(f"{k}: max rel {err.max():.2e}, first bad step {first}; gpu {a[..., first].ravel()[:6]} oracle {b[..., first].ravel()[:6]}")
            else:
                print(f"{k}: max rel {err.max():.2e}; gpu {a.ravel()[:6]} oracle {b.ravel()[:6]}")
