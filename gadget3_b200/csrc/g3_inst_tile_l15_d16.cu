// instantiation unit: tile kernel, T = double, band width 16, LEAN = 15 (features compiled out: see g3_tile.cuh)
#include "g3_ntan.h"
#include "g3_tile.cuh"
#include "g3_launch.h"
const void* g3i::tile_l15_d16(bool smem) {
    return smem ? (const void*)g3t::eval_tile_kernel<double, 16, true, 512, 0, false, 15> : (const void*)g3t::eval_tile_kernel<double, 16, false, 512, 0, false, 15>;
}
