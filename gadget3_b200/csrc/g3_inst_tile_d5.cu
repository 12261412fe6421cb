// instantiation unit: tile kernel, T = double, band width 5
#include "g3_ntan.h"
#include "g3_tile.cuh"
#include "g3_launch.h"
const void* g3i::tile_d5(bool smem) {
    return smem ? (const void*)g3t::eval_tile_kernel<double, 5, true, 512> : (const void*)g3t::eval_tile_kernel<double, 5, false, 512>;
}
// more warps per tile: 24 at 80 registers, 32 at 64, with the avoid_zero chains of 2 rows interleaved instead of 4
const void* g3i::tile_d5_w24(bool smem) {
    return smem ? (const void*)g3t::eval_tile_kernel<double, 5, true, 768, 2> : (const void*)g3t::eval_tile_kernel<double, 5, false, 768, 2>;
}
const void* g3i::tile_d5_w32(bool smem) {
    return smem ? (const void*)g3t::eval_tile_kernel<double, 5, true, 1024, 2> : (const void*)g3t::eval_tile_kernel<double, 5, false, 1024, 2>;
}
#ifdef G3T_PHASE_CLOCKS
// profiling build only: read and reset the per-phase cycle sums of this unit's kernels (tools/gpu_phase_clocks.py)
#include <cuda_runtime.h>
extern "C" int g3cuda_debug_phase_cycles(unsigned long long* out) {
    unsigned long long z[16] = {0};
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(out, g3t::g3t_phase_cycles, sizeof z) != cudaSuccess) return -1;
    return cudaMemcpyToSymbol(g3t::g3t_phase_cycles, z, sizeof z) == cudaSuccess ? 0 : -1;
}
#endif
