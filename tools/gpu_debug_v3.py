"""C2 golden vectors on the tile kernel under several partitions: per-component errors against the oracle."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from _util import golden
from gadget3_b200.configs import CONFIGS
from gadget3_b200.run_cuda import g3_cuda_adfun
from oracle.oracle_lib import Oracle
model, p = CONFIGS["C2"]()
fn = g3_cuda_adfun(model, p, device=0)
orc = Oracle(*model.pack(p))
g = golden("C2")
par = g["par"] if len(sys.argv) < 2 else g["par"][:int(sys.argv[1])]
o_nll, o_comp, _ = orc.eval_batch(par, want_comp=True, nthreads=8)
fn.handle.set_kernel(4)
for warps, seg in ((0, 0), (0, -1), (7, 0), (1, -1), (0, 4)):
    fn.handle.set_tile_partition(warps, seg)
    for rep in range(2):
        nll, comp, _ = fn.handle.eval_batch(par, want_comp=True)
        rel = np.abs(nll - o_nll) / np.abs(o_nll)
        crel = (np.abs(comp - o_comp) / np.maximum(np.abs(o_comp), 1e-300))[:, :4].max(axis=0)
        print(f"warps {warps} seg {seg} rep {rep}: nll rel max {rel.max():.3g} at {rel.argmax()}; comps {crel}")
