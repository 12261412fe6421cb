// instantiation unit: thread kernel, Dual<G3_NTAN>, band width 5
#include "g3_ntan.h"
#include "g3_thread_kernel.cuh"
#include "g3_launch.h"
const void* g3i::thread_g5(bool cm) {
    using D = g3::Dual<G3_NTAN>;
    return cm ? (const void*)g3::eval_thread_kernel<D, g3::TPB, 5, true> : (const void*)g3::eval_thread_kernel<D, g3::TPB, 5, false>;
}
