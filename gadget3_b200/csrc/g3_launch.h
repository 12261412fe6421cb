// g3_launch.h -- kernel entry points of the instantiation units (g3_inst_*.cu), as host-side function pointers
// for cudaLaunchKernel / cudaFuncSetAttribute / the occupancy query in g3cuda.cu. One unit per scalar type and
// band width so that nvcc compiles them in parallel.
#pragma once
namespace g3i {
constexpr unsigned REPORT_STRIDE = 32;      // workspace stride of the single-vector report variant
// tile kernel (g3_tile.cuh): T = double | Dual<NTAN>, NW = 5 | 16, state in shared memory or in the slab
const void* tile_d5(bool smem);
const void* tile_d5_w24(bool smem);
const void* tile_d5_w32(bool smem);
const void* tile_d16(bool smem);
const void* tile_g5(bool smem);
const void* tile_g16(bool smem);
// ... the same with the XF code (models with g3a_otherfood, g3a_spawn or prey_preferences), 16 warps per tile
const void* tile_x_d5(bool smem);
const void* tile_x_d16(bool smem);
const void* tile_x_g5(bool smem);
const void* tile_x_g16(bool smem);
// ... plain values with the features a model does not use compiled out (LEAN masks of g3_tile.cuh), 16 warps per tile
const void* tile_l15_d5(bool smem);
const void* tile_l12_d5(bool smem);
const void* tile_l15_d16(bool smem);
const void* tile_l11_d16(bool smem);
// thread kernel (g3_thread_kernel.cuh): (big = 1024-thread variant, plain doubles only), constmat
const void* thread_d5(bool big, bool constmat);
const void* thread_d16(bool constmat);
const void* thread_g5(bool constmat);
const void* thread_g16(bool constmat);
const void* thread_report(bool wide, bool constmat);
}  // namespace g3i
