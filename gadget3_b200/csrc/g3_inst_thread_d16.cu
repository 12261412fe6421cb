// instantiation unit: thread kernel, plain doubles, band width 16
#include "g3_thread_kernel.cuh"
#include "g3_launch.h"
const void* g3i::thread_d16(bool cm) {
    return cm ? (const void*)g3::eval_thread_kernel<double, g3::TPB, 16, true> : (const void*)g3::eval_thread_kernel<double, g3::TPB, 16, false>;
}
