// instantiation unit: thread kernel, single-vector report run (one warp, small workspace stride)
#include "g3_thread_kernel.cuh"
#include "g3_launch.h"
const void* g3i::thread_report(bool wide, bool cm) {
    constexpr unsigned S = g3i::REPORT_STRIDE;
    if (wide) return cm ? (const void*)g3::eval_thread_kernel<double, g3::TPB, 16, true, true, S> : (const void*)g3::eval_thread_kernel<double, g3::TPB, 16, false, true, S>;
    return cm ? (const void*)g3::eval_thread_kernel<double, g3::TPB, 5, true, true, S> : (const void*)g3::eval_thread_kernel<double, g3::TPB, 5, false, true, S>;
}
