// instantiation unit: tile kernel with the XF code (g3a_otherfood, g3a_spawn, prey_preferences), T = g3::Dual<G3_NTAN>, band width 5
#include "g3_ntan.h"
#include "g3_tile.cuh"
#include "g3_launch.h"
const void* g3i::tile_x_g5(bool smem) {
    return smem ? (const void*)g3t::eval_tile_kernel<g3::Dual<G3_NTAN>, 5, true, 512, 0, true> : (const void*)g3t::eval_tile_kernel<g3::Dual<G3_NTAN>, 5, false, 512, 0, true>;
}
