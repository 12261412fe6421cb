"""Throughput of the tile kernel under different partitions (warps per tile, rows per unit) on one config.
Usage: python tools/gpu_tile_sweep.py CONFIG BATCH "W:seg,W:seg,..." """
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gadget3_b200.configs import CONFIGS, parameter_batch
from gadget3_b200.run_cuda import g3_cuda_adfun
name, B = sys.argv[1], int(sys.argv[2])
combos = [tuple(int(v) for v in c.split(":")) for c in sys.argv[3].split(",")]
model, p = CONFIGS[name]()
fn = g3_cuda_adfun(model, p, device=0)
full = fn.full_par(parameter_batch(p, B, seed=11))
d_par = torch.from_numpy(full).cuda()
d_nll = torch.empty(B, dtype=torch.float64, device="cuda")
fn.handle.set_kernel(4)
ref = None
for W, seg in combos:
    try:
        fn.handle.set_tile_partition(W, seg)
    except Exception as e:
        print(W, seg, "refused:", e); continue
    st = torch.cuda.current_stream()
    for _ in range(2):
        fn.handle.eval_batch_device(d_par.data_ptr(), B, d_nll.data_ptr(), stream=st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        fn.handle.eval_batch_device(d_par.data_ptr(), B, d_nll.data_ptr(), stream=st.cuda_stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    got = d_nll.cpu().numpy()
    if ref is None: ref = got.copy()
    info = fn.handle.tile_info()
    print(f"{name} B={B} W={W} seg={seg}: {ms:.2f} ms, {B / ms * 1e3:,.0f} evals/s; warps {info['warps']}; max rel diff vs first {np.max(np.abs(got - ref) / np.abs(ref)):.2g}", flush=True)
