"""Where a tile's time goes, phase by phase: runs C2 on a profiling build of the library (-DG3T_PHASE_CLOCKS: warp 0 of every
CTA sums clock64() differences between the phase boundaries) and prints the shares. Usage: python tools/gpu_phase_clocks.py LIB [CONFIG] [B]"""
import ctypes as C
import os, sys
import numpy as np
os.environ["G3CUDA_LIBRARY"] = os.path.abspath(sys.argv[1])
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gadget3_b200 import _lib
from gadget3_b200.configs import CONFIGS, parameter_batch
from gadget3_b200.run_cuda import g3_cuda_adfun
name = sys.argv[2] if len(sys.argv) > 2 else "C2"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4736
model, p = CONFIGS[name]()
fn = g3_cuda_adfun(model, p, device=0)
full = fn.full_par(parameter_batch(p, B, seed=11))
lib = _lib.load()
lib.g3cuda_debug_phase_cycles.argtypes = [C.POINTER(C.c_uint64)]
out = (C.c_uint64 * 16)()
fn.handle.eval_batch(full); lib.g3cuda_debug_phase_cycles(out)
fn.handle.eval_batch(full); lib.g3cuda_debug_phase_cycles(out)
v = np.array(list(out)[:11], dtype=float)
names = ["tile set-up", "migration", "predation: per-unit suitable biomass", "predation: totals", "pass 1: predation + ghost rows",
         "pass 2: growth", "hand-over blocks (steps that also collect)", "collect tasks", "compare / hand-over blocks", "end-of-step barrier wait",
         "keeper / ageing"]
print(f"{name}, {B} vectors, kernel {fn.handle.last_kernel_ms():.3f} ms; cycles as warp 0 of each CTA sees them:")
for n, x in zip(names, v):
    print(f"  {n:45s} {100 * x / v.sum():5.1f} %   {x / 148 / model.n_steps:9.0f} cycles per step and CTA")
