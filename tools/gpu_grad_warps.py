"""C3 (value + gradient over all optimised parameters) on the tile kernel at several warp counts per tile: the Dual<N>
instantiations with fewer threads get more registers per thread (256 threads: 255 registers, no spills)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gadget3_b200.configs import CONFIGS, parameter_batch
from gadget3_b200.run_cuda import g3_cuda_adfun
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
warps = [int(w) for w in sys.argv[3].split(",")] if len(sys.argv) > 3 else [15, 11, 7]
model, p = CONFIGS[name]()
fn = g3_cuda_adfun(model, p, device=0)
full = fn.full_par(parameter_batch(p, B, seed=5))
tix = fn.opt_idx.astype(np.int32)
print(name, fn.handle.tile_info(), flush=True)
ref = None
for w in warps:
    fn.handle.set_tile_warps(w)
    fn.handle.eval_batch(full[:512], tangent_idx=tix)
    best = 1e9
    for _ in range(2):
        t0 = time.perf_counter()
        out = fn.handle.eval_batch(full, tangent_idx=tix)
        best = min(best, (time.perf_counter() - t0) * 1e3)
    msg = ""
    if ref is None:
        ref = out
    else:
        sc = np.abs(ref[2]).max(axis=1, keepdims=True)
        msg = f" vs first: nll {np.max(np.abs(ref[0] - out[0]) / np.abs(ref[0])):.2g}, grad {np.max(np.abs(ref[2] - out[2]) / sc):.2g}"
    print(f"{os.environ.get('G3CUDA_LIBRARY', 'default')[-20:]} {name} warps {w}: {B} x {len(tix)} directions, host call {best:.1f} ms = {B / best * 1e3:,.0f} (value+gradient)/s{msg}", flush=True)
