import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gadget3_b200.configs import CONFIGS, parameter_batch
from gadget3_b200.run_cuda import g3_cuda_adfun
from oracle.oracle_lib import Oracle
model, p = CONFIGS["MINI_ANDERSEN"]()
fn = g3_cuda_adfun(model, p)
orc = Oracle(*model.pack(p))
P = np.vstack([fn.par[None, :], parameter_batch(p, 7, seed=12)])
full = fn.full_par(P)
g, gc, _ = fn.handle.eval_batch(full, want_comp=True)
o, oc, _ = orc.eval_batch(full, want_comp=True)
print("nll gpu", g); print("nll orc", o)
print("comp gpu", gc[:3]); print("comp orc", oc[:3])
