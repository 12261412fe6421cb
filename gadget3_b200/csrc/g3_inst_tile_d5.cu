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
