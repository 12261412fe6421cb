// instantiation unit: thread kernel, plain doubles, band width 5
#include "g3_thread_kernel.cuh"
#include "g3_launch.h"
const void* g3i::thread_d5(bool big, bool cm) {
    if (big) return cm ? (const void*)g3::eval_thread_kernel<double, g3::TPB_BIG, 5, true> : (const void*)g3::eval_thread_kernel<double, g3::TPB_BIG, 5, false>;
    return cm ? (const void*)g3::eval_thread_kernel<double, g3::TPB, 5, true> : (const void*)g3::eval_thread_kernel<double, g3::TPB, 5, false>;
}
