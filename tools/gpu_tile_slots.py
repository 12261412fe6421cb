"""Throughput of the tile kernel against the number of tiles in flight (g3cuda_set_tile_slots): the per-tile slabs of the
tiles in flight are the kernel's L2 working set; where that exceeds the 126 MB of L2 the slabs stream from HBM."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gadget3_b200.configs import CONFIGS, parameter_batch
from gadget3_b200.run_cuda import g3_cuda_adfun
for name, B, grad in (("C4", 16384, False), ("C5", 32768, False), ("C2", 2048, True), ("C2", 65536, False)):
    model, p = CONFIGS[name]()
    fn = g3_cuda_adfun(model, p, device=0)
    full = fn.full_par(parameter_batch(p, B, seed=5))
    tix = fn.opt_idx.astype(np.int32) if grad else None
    info = fn.handle.tile_info()
    ref = None
    for slots in (0, 132, 116, 100, 84, 68):
        fn.handle.set_tile_slots(slots)
        fn.handle.eval_batch(full[:256], tangent_idx=tix)
        best = 1e9
        for _ in range(2):
            t0 = time.perf_counter(); out = fn.handle.eval_batch(full, tangent_idx=tix); best = min(best, time.perf_counter() - t0)
        if ref is None: ref = out
        same = np.array_equal(ref[0], out[0])
        print(f"{name}{' grad' if grad else ''} slots {slots or 148}: {B / best:,.0f} evals/s (tables {info['table_entries']}, state {info['state_entries']}, in smem {info['state_in_smem']}; same bits {same})", flush=True)
