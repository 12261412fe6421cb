// instantiation unit: thread kernel, Dual<G3_NTAN>, band width 16
#include "g3_ntan.h"
#include "g3_thread_kernel.cuh"
#include "g3_launch.h"
const void* g3i::thread_g16(bool cm) {
    using D = g3::Dual<G3_NTAN>;
    return cm ? (const void*)g3::eval_thread_kernel<D, g3::TPB, 16, true> : (const void*)g3::eval_thread_kernel<D, g3::TPB, 16, false>;
}
