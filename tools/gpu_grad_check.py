import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gadget3_b200.configs import CONFIGS, parameter_batch
from gadget3_b200.run_cuda import g3_cuda_adfun
from oracle.oracle_lib import Oracle
name = sys.argv[1] if len(sys.argv) > 1 else "C4"
model, p = CONFIGS[name]()
fn = g3_cuda_adfun(model, p)
orc = Oracle(*model.pack(p))
P = parameter_batch(p, 2, seed=77)
full = fn.full_par(P)
fn.handle.set_kernel(3)
g_nll, g = fn.gr_batch(P)
_, _, o = orc.eval_batch(full, tangent_idx=fn.opt_idx.astype(np.int32), nthreads=8)
scale = np.abs(o).max(axis=1, keepdims=True)
err = np.abs(g - o) / scale
print("max err", err.max(), "at", np.unravel_index(err.argmax(), err.shape))
for i in np.argsort(-err[0])[:8]:
    print(p.switch[fn.opt_idx[i]], g[0, i], o[0, i], err[0, i])
