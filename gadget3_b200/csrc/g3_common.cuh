// g3_common.cuh -- device helpers shared by the evaluation kernels: parameter seeding for forward-mode
// tangents, the slot-program interpreter (include/g3cuda_spec.h G3OP_*), the suitability functions.
// Compiles under nvcc (product) and under g++ (tests/emu: the kernel-logic test harness, never loaded by the product).
#pragma once
#include "g3_math.cuh"
#include "../../include/g3cuda_spec.h"

namespace g3 {

template <class T> __device__ __forceinline__ T mkpar_g(const double* __restrict__ prow, const int* seed, int i) {
    T r(prow[i]);
    if constexpr (Tan<T>::n > 0) {
#pragma unroll
        for (int k = 0; k < Tan<T>::n; k++) if (seed[k] == i) r.d[k] = 1.0;
    }
    return r;
}

// slot programs (include/g3cuda_spec.h G3OP_*), parameters read from the vector's row in global memory
template <class T> __device__ G3_NOINLINE T eval_slot_g(const int32_t* __restrict__ I, const double* __restrict__ R,
                                                         const double* __restrict__ prow, const int* seed, int slot) {
    const int32_t* tab = I + I[G3H_SLOT_OFF] + 2 * slot;
    const int32_t* c = I + tab[0];
    int n = tab[1];
    T st[G3_SLOT_STACK];
    int sp = 0;
    for (int i = 0; i < n; i++) {
        int op = c[i];
        switch (op) {
        case G3OP_CONST: st[sp++] = T(R[c[++i]]); break;
        case G3OP_PARAM: st[sp++] = mkpar_g<T>(prow, seed, c[++i]); break;
        case G3OP_NAN: st[sp++] = T(nan("")); break;
        case G3OP_ADD: sp--; st[sp - 1] = st[sp - 1] + st[sp]; break;
        case G3OP_SUB: sp--; st[sp - 1] = st[sp - 1] - st[sp]; break;
        case G3OP_MUL: sp--; st[sp - 1] = st[sp - 1] * st[sp]; break;
        case G3OP_DIV: sp--; st[sp - 1] = st[sp - 1] / st[sp]; break;
        case G3OP_POW: sp--; st[sp - 1] = xpow(st[sp - 1], st[sp]); break;
        case G3OP_NEG: st[sp - 1] = -st[sp - 1]; break;
        case G3OP_EXP: st[sp - 1] = xexp(st[sp - 1]); break;
        case G3OP_LOG: st[sp - 1] = xlog(st[sp - 1]); break;
        case G3OP_SQRT: st[sp - 1] = xsqrt(st[sp - 1]); break;
        case G3OP_AVOID_ZERO: st[sp - 1] = avoid_zero(st[sp - 1]); break;
        default: break;
        }
    }
    return st[0];
}

// g3_suitability_* of prey length `ml` for predator length `pl` (R/suitability.R:5-94); p[] = the pair's slot values
template <class T> __device__ __forceinline__ T suitability_of(int kind, const T* p, double ml, double pl) {
    if (kind == G3SUIT_EXPL50) return 1.0 / (1.0 + xexp(-p[0] * (ml - p[1])));
    if (kind == G3SUIT_ANDERSEN) {
        const T x = xlog(avoid_zero(T(pl / ml)));
        const T d = x - p[1];
        const T lo = 1.0 / (1.0 + xexp(1000.0 * (p[1] - x)));        // bounded_vec(., 0, 1), R/aab_env.R:51-55
        const T hi = 1.0 / (1.0 + xexp(1000.0 * (x - p[1])));
        const T a2 = avoid_zero(p[2]);
        return p[0] + a2 * xexp(-(d * d) / avoid_zero(p[3])) * lo + a2 * xexp(-(d * d) / avoid_zero(p[4])) * hi;
    }
    if (kind == G3SUIT_ANDERSENFLEET) {     // pl = the prey's largest midlen; R/suitability.R:32-58
        const T x = xlog(T(pl / ml));
        const T d = x - p[1];
        const T lo = 1.0 / (1.0 + xexp(1000.0 * (p[1] - x)));
        const T hi = 1.0 / (1.0 + xexp(1000.0 * (x - p[1])));
        return p[0] + p[2] * xexp(-(d * d) / p[3]) * lo + p[2] * xexp(-(d * d) / p[4]) * hi;
    }
    if (kind == G3SUIT_GAMMA)               // R/suitability.R:60-67
        return xpow(T(ml) / ((p[0] - 1.0) * p[1] * p[2]), p[0] - 1.0) * xexp(p[0] - 1.0 - T(ml) / (p[1] * p[2]));
    if (kind == G3SUIT_EXPONENTIAL || kind == G3SUIT_RICHARDS) {       // R/suitability.R:69-75,90-94
        const T e = p[3] / (1.0 + xexp(-p[0] - p[1] * ml - p[2] * pl));
        return kind == G3SUIT_RICHARDS ? xpow(e, 1.0 / p[4]) : e;
    }
    if (kind == G3SUIT_STRAIGHTLINE) return p[0] + p[1] * ml;           // R/suitability.R:77-79
    return p[0];
}

}  // namespace g3
