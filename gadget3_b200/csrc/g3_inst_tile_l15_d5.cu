// instantiation unit: tile kernel, T = double, band width 5, LEAN = 15 (features compiled out: see g3_tile.cuh)
#include "g3_ntan.h"
#include "g3_tile.cuh"
#include "g3_launch.h"
const void* g3i::tile_l15_d5(bool smem) {
    return smem ? (const void*)g3t::eval_tile_kernel<double, 5, true, 512, 0, false, 15> : (const void*)g3t::eval_tile_kernel<double, 5, false, 512, 0, false, 15>;
}
