// g3cuda.cu -- C ABI (include/g3cuda.h) over the evaluation kernel: spec validation, derived
// device tables, shared-memory layout, launches. No CPU evaluation path exists in this library:
// without a CUDA device every compute entry point returns G3CUDA_ENODEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/g3cuda.h"
#include "g3_kernel.cuh"
#include "g3_warp_kernel.cuh"
#include "g3_thread_kernel.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}
#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(G3CUDA_ECUDA, "%s: %s", #x, cudaGetErrorString(e_)); } while (0)

#ifndef G3_NTAN
#define G3_NTAN 4
#endif
constexpr int NTAN = G3_NTAN;           // tangent directions per gradient pass
using DualT = g3::Dual<NTAN>;

double h_avoid_zero(double a) {         // R/aab_env.R:24-31 on the host, for parameter-independent terms
    double p = a * 1000.0;
    return (std::fmax(p, 0.0) + std::log1p(std::exp(-std::fabs(p)))) / 1000.0;
}

template <class T> T* dev_copy(const std::vector<T>& v) {
    T* d = nullptr;
    size_t n = std::max<size_t>(v.size(), 1) * sizeof(T);
    if (cudaMalloc(&d, n) != cudaSuccess) return nullptr;
    if (!v.empty() && cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return nullptr; }
    return d;
}

}  // namespace

struct g3cuda_model {
    int device = 0;
    std::vector<int32_t> I; std::vector<double> R;
    g3::KArgs base{};                   // everything except per-call pointers and the smem layout
    std::vector<void*> owned;           // device allocations
    int nthreads = 0;
    int num_sms = 0;
    int64_t report_len = 0, rep_detail = 0;      // rep_detail: first double of the g3a_report_detail section
    int64_t launches = 0;
    // per-call scratch (grow-only)
    double* d_par = nullptr; double* d_nll = nullptr; double* d_comp = nullptr; double* d_grad = nullptr;
    int32_t* d_tidx = nullptr; double* d_report = nullptr; double* d_out = nullptr;
    size_t cap_par = 0, cap_nll = 0, cap_comp = 0, cap_grad = 0, cap_tidx = 0, cap_out = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    // smem element counts
    int sm_elems = 0;
    size_t smem_optin = 0;
    bool only_thread = false;           // model uses actions only the thread kernel implements (migration, stock predators)
    bool cta_ok_double = false, cta_ok_dual = false;    // CTA kernel: shared-memory layout fits (values / duals)
    bool thread_ok = false;             // model fits the thread-per-evaluation kernel
    int t_entries = 0, t_maxn = 0;      // workspace entries per thread; widest growth band
    void* t_ws = nullptr; size_t t_ws_bytes = 0;
    cudaEvent_t ws_free = nullptr;      // recorded after every launch that uses the workspace: the next one waits for it,
                                        // whichever stream it is on (the workspace is per handle, not per stream)
    bool warp_ok = false;               // model fits the warp-per-evaluation kernel
    int force_kernel = 0;               // 0 = automatic, 1 = CTA-per-evaluation kernel, 2 = warp kernel (debug / tests)
};

namespace {

struct Layout { g3::KArgs a; size_t bytes; };

// Shared-memory layout for scalar type of `tsize` bytes.
Layout make_layout(const g3cuda_model* m, size_t tsize) {
    using namespace g3;
    Layout out; out.a = m->base;
    KArgs& A = out.a;
    const int32_t* I = m->I.data();
    int off = 0;
    auto take = [&](int n) { int o = off; off += n; return o; };
    A.sm_par_bytes = (int)(((size_t)A.NPAR * 8 + 15) / 16 * 16);
    A.o_slot = take(A.NSLOT);
    A.o_num = take(A.NC); A.o_wgt = take(A.NC); A.o_tn = take(A.NC); A.o_tw = take(A.NC);
    A.o_alt = A.any_const_mat ? take(A.NC) : A.o_tn;
    for (int x = 0; x < MAXX; x++) A.o_xslot[x] = -1;
    for (int c = 0; c < A.NCOMP; c++) {
        const int32_t* C = I + I[G3H_COMP_OFF] + c * G3C_LEN;
        if (C[G3C_MODE] == 0 && A.o_xslot[C[G3C_XBUF]] < 0) A.o_xslot[C[G3C_XBUF]] = take(A.NC);
    }
    A.o_xg = 0;
    A.o_pw = off;
    for (int s = 0; s < A.NS; s++) A.o_pws[s] = take(I[I[G3H_STOCK_OFF] + s * G3S_LEN + G3S_L]);
    A.o_suit = take(A.NPP * A.maxL);
    int maxg = 0;
    for (int s = 0; s < A.NS; s++) {
        const int32_t* St = I + I[G3H_STOCK_OFF] + s * G3S_LEN;
        int n = St[G3S_GROW_N], L = St[G3S_L];
        if (n < 0) { A.o_g[s] = A.o_gr[s] = A.o_gn[s] = 0; continue; }
        int g = (n + 1) * L; maxg = std::max(maxg, g);
        A.o_g[s] = take(g);
        if (St[G3S_MAT_KIND] == G3MAT_CONTINUOUS && St[G3S_N_OUT] > 0) { A.o_gr[s] = take(g); A.o_gn[s] = take(g); }
        else { A.o_gr[s] = A.o_gn[s] = A.o_g[s]; }
    }
    A.o_gscr = take(3 * maxg);
    int rd0 = off;
    for (int s = 0; s < A.NS; s++) {
        const int32_t* St = I + I[G3H_STOCK_OFF] + s * G3S_LEN;
        A.o_rd[s] = take(St[G3S_N_RENEW] * St[G3S_L]);
    }
    A.o_rwt = take(off - rd0);
    for (int c = 0; c < A.NCOMP; c++) {
        const int32_t* C = I + I[G3H_COMP_OFF] + c * G3C_LEN;
        bool si = C[G3C_FUNC] == G3F_SI_LOG || C[G3C_FUNC] == G3F_SI_LINEAR;
        A.o_ms[c] = take(C[G3C_N_DEST] * (si ? C[G3C_N_TIMES] : 1));
    }
    A.o_red = take((NRED + MAXTS + 2) * 32);
    out.bytes = (size_t)A.sm_par_bytes + (size_t)off * tsize;
    return out;
}

// Shared-memory layout of ONE warp's slice for the warp-per-evaluation kernel.
Layout make_wlayout(const g3cuda_model* m, size_t tsize) {
    using namespace g3;
    Layout out; out.a = m->base;
    KArgs& A = out.a;
    const int32_t* I = m->I.data();
    int off = 0;
    auto take = [&](int n) { int o = off; off += n; return o; };
    A.sm_par_bytes = (int)(((size_t)A.NPAR * 8 + 15) / 16 * 16);
    A.o_slot = take(A.NSLOT);
    A.o_num = take(A.NC); A.o_wgt = take(A.NC);
    int ntr = 0, nmv = 0, nmort = 0, maxg = 0;
    for (int s = 0; s < A.NS; s++) {
        const int32_t* St = I + I[G3H_STOCK_OFF] + s * G3S_LEN;
        int ncell = St[G3S_L] * St[G3S_NAGE] * St[G3S_NAREA];
        A.w_tr_off[s] = ntr;                 // (same numbering as w_inc_tr, built at model_create)
        if (St[G3S_GROW_N] >= 0 && St[G3S_N_OUT] > 0) ntr += ncell;
        A.w_mv_off[s] = nmv;
        if (St[G3S_AGE_ENABLE] && St[G3S_AGE_N_OUT] > 0) nmv += 2 * St[G3S_NAREA] * St[G3S_L];
        A.w_mort_off[s] = nmort;
        if (St[G3S_MORT_OFF] >= 0) nmort += St[G3S_NAGE] * St[G3S_NAREA] * A.w_ndt;
    }
    A.o_tn = take(ntr); A.o_tw = take(ntr);
    A.o_alt = A.any_const_mat ? take(A.NC) : A.o_tn;
    A.w_o_mv = take(nmv);
    A.w_o_mort = take(nmort);
    for (int x = 0; x < MAXX; x++) A.o_xslot[x] = x < A.NX ? take(A.NC) : -1;
    for (int s = 0; s < A.NS; s++) A.o_pws[s] = take(I[I[G3H_STOCK_OFF] + s * G3S_LEN + G3S_L]);
    A.o_suit = take(A.NPP * A.maxL);
    for (int s = 0; s < A.NS; s++) {
        const int32_t* St = I + I[G3H_STOCK_OFF] + s * G3S_LEN;
        int n = St[G3S_GROW_N], L = St[G3S_L];
        A.o_rd[s] = take(St[G3S_N_RENEW] * L);
        A.w_o_rw[s] = take(St[G3S_N_RENEW] * L);
        if (n < 0) { A.o_g[s] = A.o_gr[s] = A.o_gn[s] = A.w_o_dw[s] = 0; continue; }
        int g = (n + 1) * L; maxg = std::max(maxg, g);
        A.w_o_dw[s] = take(g);
        A.o_g[s] = take(g * A.w_ndt);
        if (St[G3S_MAT_KIND] == G3MAT_CONTINUOUS && St[G3S_N_OUT] > 0) { A.o_gr[s] = take(g * A.w_ndt); A.o_gn[s] = take(g * A.w_ndt); }
        else { A.o_gr[s] = A.o_gn[s] = A.o_g[s]; }
    }
    A.o_gscr = take(std::max(3 * maxg, 32));
    for (int c = 0; c < A.NCOMP; c++) {
        const int32_t* C = I + I[G3H_COMP_OFF] + c * G3C_LEN;
        bool si = C[G3C_FUNC] == G3F_SI_LOG || C[G3C_FUNC] == G3F_SI_LINEAR;
        A.o_ms[c] = take(C[G3C_N_DEST] * (si ? C[G3C_N_TIMES] : 1));
    }
    A.w_o_tot = take(MAXTS + A.NNLL);
    out.bytes = ((size_t)A.sm_par_bytes + (size_t)off * tsize + 15) / 16 * 16;
    A.w_warp_bytes = (int)out.bytes;
    return out;
}

// Workspace entries of the thread-per-evaluation kernel (one T per entry per thread).
int make_tlayout(g3cuda_model* m, g3::KArgs& A) {
    using namespace g3;
    const int32_t* I = m->I.data();
    int off = 0;
    auto take = [&](int n) { int o = off; off += n; return o; };
    A.t_slot = take(A.NSLOT);
    A.t_num = take(A.NC); A.t_wgt = take(A.NC);
    int ntr = 0, nmv = 0, maxg = 0;
    for (int s = 0; s < A.NS; s++) {
        const int32_t* St = I + I[G3H_STOCK_OFF] + s * G3S_LEN;
        int ncell = St[G3S_L] * St[G3S_NAGE] * St[G3S_NAREA];
        A.w_tr_off[s] = ntr;
        if (St[G3S_GROW_N] >= 0 && St[G3S_N_OUT] > 0) ntr += ncell;
        A.w_mv_off[s] = nmv;
        if (St[G3S_AGE_ENABLE] && St[G3S_AGE_N_OUT] > 0) nmv += 2 * St[G3S_NAREA] * St[G3S_L];
    }
    A.t_tn = take(ntr); A.t_tw = take(ntr); A.t_mv = take(nmv);
    A.t_suit = off;
    for (int q = 0; q < A.NPP && q < 32; q++) {
        const int32_t* Q = I + I[G3H_PP_OFF] + q * G3PP_LEN;
        const int32_t* P = I + I[G3H_PRED_OFF] + Q[G3PP_PRED] * G3P_LEN;
        int PL = P[G3P_KIND] == G3PRED_STOCK ? I[I[G3H_STOCK_OFF] + P[G3P_STOCK] * G3S_LEN + G3S_L] : 1;
        A.t_suitq[q] = take(PL * I[I[G3H_STOCK_OFF] + Q[G3PP_PREY] * G3S_LEN + G3S_L]);
    }
    for (int p = 0; p < I[G3H_NPRED] && p < MAXS; p++) {
        const int32_t* P = I + I[G3H_PRED_OFF] + p * G3P_LEN;
        A.t_pts[p] = A.t_pmc[p] = 0;
        if (P[G3P_KIND] != G3PRED_STOCK) continue;
        const int32_t* Pp = I + I[G3H_STOCK_OFF] + P[G3P_STOCK] * G3S_LEN;
        A.t_pts[p] = take(Pp[G3S_NAREA] * Pp[G3S_L]);
        A.t_pmc[p] = take(Pp[G3S_L]);
    }
    for (int s = 0; s < A.NS; s++) {
        const int32_t* St = I + I[G3H_STOCK_OFF] + s * G3S_LEN;
        A.t_mig[s] = St[G3S_MIGRATE_OFF] >= 0 ? take(A.NSTEPS * St[G3S_NAREA] * St[G3S_NAREA]) : 0;
    }
    for (int s = 0; s < A.NS; s++) {
        const int32_t* St = I + I[G3H_STOCK_OFF] + s * G3S_LEN;
        int n = St[G3S_GROW_N], L = St[G3S_L];
        A.t_mort[s] = take(St[G3S_MORT_OFF] >= 0 ? St[G3S_NAGE] * St[G3S_NAREA] * A.w_ndt : 0);
        A.t_rd[s] = take(St[G3S_N_RENEW] * L);
        A.t_rw[s] = take(St[G3S_N_RENEW] * L);
        A.t_pw[s] = take(n >= 0 ? L : 0);
        if (n < 0) { A.t_g[s] = A.t_gr[s] = A.t_gn[s] = 0; continue; }
        int g = (n + 1) * L; maxg = std::max(maxg, g);
        A.t_g[s] = take(g * A.w_ndt);
        if (St[G3S_MAT_KIND] == G3MAT_CONTINUOUS && St[G3S_N_OUT] > 0) { A.t_gr[s] = take(g * A.w_ndt); A.t_gn[s] = take(g * A.w_ndt); }
        else { A.t_gr[s] = A.t_gn[s] = A.t_g[s]; }
    }
    A.t_scr = take(3 * maxg);
    for (int c = 0; c < A.NCOMP; c++) {
        const int32_t* C = I + I[G3H_COMP_OFF] + c * G3C_LEN;
        bool si = C[G3C_FUNC] == G3F_SI_LOG || C[G3C_FUNC] == G3F_SI_LINEAR;
        A.t_ms[c] = take(C[G3C_N_DEST] * (si ? C[G3C_N_TIMES] : 1));
    }
    A.t_ctot = take(A.NNLL);
    for (int x = 0; x < MAXX; x++) A.t_x[x] = x < A.NX ? take(A.NC) : 0;
    int ncolg = 0;
    for (int s = 0; s < A.NS; s++) {
        const int32_t* St = I + I[G3H_STOCK_OFF] + s * G3S_LEN;
        A.t_colbase[s] = ncolg; ncolg += St[G3S_NAGE] * St[G3S_NAREA];
        A.t_cmat[s] = (St[G3S_GROW_N] >= 0 && St[G3S_N_OUT] > 0 && St[G3S_MAT_KIND] == G3MAT_CONSTANT) ? take(St[G3S_L] * St[G3S_NAGE] * St[G3S_NAREA]) : 0;
    }
    A.t_remf = take(ncolg);
    { const char* e = getenv("G3_CBLK"); A.t_cblk = e && atoi(e) > 0 ? atoi(e) : 2; }     // (C4 on B200: 1: 48.9 k, 2: 51.9 k, 3: 50.4 k, all columns: 46.3 k evals/s)
    return off;
}

template <class T, int CB, int NW, bool CM>
int launch_thread_variant(g3cuda_model* m, g3::KArgs& A, long long work, cudaStream_t st) {
    auto kern = g3::eval_thread_kernel<T, CB, NW, CM>;
    const size_t smem = sizeof(T) * g3::MAXTS * CB;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CB, smem));
    if (per_sm < 1) return fail(G3CUDA_EUNSUPP, "thread kernel does not fit one SM");
    // One block per SM, sized so that the batch splits into equally full waves: an evaluation is
    // one thread for its whole life, so a ragged last wave would idle most of the machine.
    const int sms = std::min<int>(m->num_sms, (int)(g3::WS_STRIDE / CB));
    const long long slots = (long long)CB * sms;
    const long long waves = (work + slots - 1) / slots;
    long long per_wave = (work + waves - 1) / waves;
    long long grid = std::min<long long>(sms, (per_wave + 31) / 32);
    if (grid < 1) grid = 1;
    int block = (int)(((per_wave + grid - 1) / grid + 31) / 32 * 32);
    block = std::max(32, std::min(block, CB));
    if ((double)m->t_entries * (double)g3::WS_STRIDE >= 4294967295.0)
        return fail(G3CUDA_EUNSUPP, "evaluation workspace too large for 32-bit element indexing (%d entries)", m->t_entries);
    const size_t need = (size_t)m->t_entries * sizeof(T) * (size_t)g3::WS_STRIDE;
    if (need > m->t_ws_bytes) {
        // the workspace may still be in use by an earlier asynchronous launch on another stream
        CUDA_TRY(cudaDeviceSynchronize());
        if (m->t_ws) cudaFree(m->t_ws);
        m->t_ws = nullptr; m->t_ws_bytes = 0;
        if (cudaMalloc(&m->t_ws, need) != cudaSuccess) { cudaGetLastError(); return fail(G3CUDA_ENOMEM, "cudaMalloc of the %zu-byte evaluation workspace failed", need); }
        m->t_ws_bytes = need;
    }
    A.t_ws = m->t_ws;
    if (!m->ws_free) CUDA_TRY(cudaEventCreateWithFlags(&m->ws_free, cudaEventDisableTiming));
    else CUDA_TRY(cudaStreamWaitEvent(st, m->ws_free, 0));
    kern<<<(unsigned)grid, block, smem, st>>>(A);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(m->ws_free, st));
    m->launches++;
    return G3CUDA_OK;
}

// single-vector report run of the thread kernel (one warp; lane 0 is the evaluation)
template <int NW, bool CM>
int launch_thread_report(g3cuda_model* m, g3::KArgs& A, cudaStream_t st) {
    auto kern = g3::eval_thread_kernel<double, g3::TPB, NW, CM, true>;
    const size_t smem = sizeof(double) * g3::MAXTS * 32;
    const size_t need = (size_t)m->t_entries * sizeof(double) * (size_t)g3::WS_STRIDE;
    if (need > m->t_ws_bytes) {
        CUDA_TRY(cudaDeviceSynchronize());
        if (m->t_ws) cudaFree(m->t_ws);
        m->t_ws = nullptr; m->t_ws_bytes = 0;
        if (cudaMalloc(&m->t_ws, need) != cudaSuccess) { cudaGetLastError(); return fail(G3CUDA_ENOMEM, "cudaMalloc of the %zu-byte evaluation workspace failed", need); }
        m->t_ws_bytes = need;
    }
    A.t_ws = m->t_ws;
    if (!m->ws_free) CUDA_TRY(cudaEventCreateWithFlags(&m->ws_free, cudaEventDisableTiming));
    else CUDA_TRY(cudaStreamWaitEvent(st, m->ws_free, 0));
    kern<<<1, 32, smem, st>>>(A);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(m->ws_free, st));
    m->launches++;
    return G3CUDA_OK;
}

template <class T>
int launch_thread(g3cuda_model* m, g3::KArgs& A, long long work, cudaStream_t st) {
    const bool cm = A.any_const_mat != 0;
    if (A.report) {
        if constexpr (std::is_same<T, double>::value) {
            if (m->t_maxn <= 4) return cm ? launch_thread_report<5, true>(m, A, st) : launch_thread_report<5, false>(m, A, st);
            return cm ? launch_thread_report<16, true>(m, A, st) : launch_thread_report<16, false>(m, A, st);
        }
        return fail(G3CUDA_EINVAL, "report runs are plain-double");
    }
    // plain doubles and more work than one 512-thread wave: the 1024-thread (64-register) variant
    const bool big = std::is_same<T, double>::value && work > (long long)g3::TPB * m->num_sms;
    if (m->t_maxn <= 4) {
        if (big) {
            if constexpr (std::is_same<T, double>::value) {
                if (cm) return launch_thread_variant<T, g3::TPB_BIG, 5, true>(m, A, work, st);
                return launch_thread_variant<T, g3::TPB_BIG, 5, false>(m, A, work, st);
            }
        }
        if (cm) return launch_thread_variant<T, g3::TPB, 5, true>(m, A, work, st);
        return launch_thread_variant<T, g3::TPB, 5, false>(m, A, work, st);
    }
    if (cm) return launch_thread_variant<T, g3::TPB, 16, true>(m, A, work, st);
    return launch_thread_variant<T, g3::TPB, 16, false>(m, A, work, st);
}

template <class T, int WPC>
int launch_warp_variant(g3cuda_model* m, const Layout& lay, long long work, cudaStream_t st) {
    auto kern = g3::eval_warp_kernel<T, WPC>;
    size_t bytes = lay.bytes * WPC;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WPC * 32, bytes));
    if (per_sm < 1) return fail(G3CUDA_EUNSUPP, "warp kernel does not fit one SM: %zu bytes of shared memory per CTA", bytes);
    long long grid = std::min<long long>((work + WPC - 1) / WPC, (long long)per_sm * m->num_sms);
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, WPC * 32, bytes, st>>>(lay.a);
    CUDA_TRY(cudaGetLastError());
    m->launches++;
    return G3CUDA_OK;
}

template <class T, int MAXT>
int launch_variant(g3cuda_model* m, const Layout& lay, int nblocks_hint, cudaStream_t st) {
    auto kern = g3::eval_kernel<T, MAXT>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.bytes));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, m->nthreads, lay.bytes));
    if (per_sm < 1) return fail(G3CUDA_EUNSUPP, "model does not fit one SM: %zu bytes of shared memory, %d threads", lay.bytes, m->nthreads);
    long long grid = std::min<long long>(nblocks_hint, (long long)per_sm * m->num_sms);
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, m->nthreads, lay.bytes, st>>>(lay.a);
    CUDA_TRY(cudaGetLastError());
    m->launches++;
    return G3CUDA_OK;
}

// Below this many evaluations the CTA-per-evaluation kernel wins: one evaluation is one thread's whole life in
// the thread kernel (C2: 67 ms whatever the batch), but 1.3 ms when a CTA shares it (an optimiser's single fn()
// call). Measured crossover on B200: ~11 k vectors (C1), ~14 k (C2) -- tools/gpu_latency.py.
static const long long SMALL_BATCH = 12288;

template <class T>
int launch(g3cuda_model* m, const g3::KArgs& call, long long work, cudaStream_t st) {
    const bool is_double = std::is_same<T, double>::value;
    if (m->thread_ok && (m->force_kernel == 0 || m->force_kernel == 3) &&
        (call.report ? true : (m->force_kernel == 3 ||
                               (is_double ? (work >= SMALL_BATCH || !m->cta_ok_double) : (work >= 16384 || !m->cta_ok_dual))))) {
        g3::KArgs a = m->base;
        a.par = call.par; a.B = call.B; a.tangent_idx = call.tangent_idx; a.K = call.K;
        a.nll = call.nll; a.comp = call.comp; a.grad = call.grad; a.report = call.report;
        return launch_thread<T>(m, a, work, st);
    }
    if (m->warp_ok && !call.report && m->force_kernel == 2) {       // (kept as measured evidence: never chosen automatically)
        Layout lay = make_wlayout(m, sizeof(T));
        lay.a.par = call.par; lay.a.B = call.B; lay.a.tangent_idx = call.tangent_idx; lay.a.K = call.K;
        lay.a.nll = call.nll; lay.a.comp = call.comp; lay.a.grad = call.grad; lay.a.report = nullptr;
        if (lay.bytes * 2 <= m->smem_optin) {
            if (lay.bytes * 4 <= m->smem_optin / 2) return launch_warp_variant<T, 4>(m, lay, work, st);
            return launch_warp_variant<T, 2>(m, lay, work, st);
        }
    }
    if (!(is_double ? m->cta_ok_double : m->cta_ok_dual))
        return fail(G3CUDA_EUNSUPP, "the CTA-per-evaluation kernel does not fit this model%s", call.report ? " (needed for g3cuda_report)" : "");
    Layout lay = make_layout(m, sizeof(T));
    // per-call fields
    lay.a.par = call.par; lay.a.B = call.B; lay.a.tangent_idx = call.tangent_idx; lay.a.K = call.K;
    lay.a.nll = call.nll; lay.a.comp = call.comp; lay.a.grad = call.grad; lay.a.report = call.report;
    int hint = (int)std::min<long long>(work, 1 << 30);
    int n = m->nthreads;
    if (n <= 128) return launch_variant<T, 128>(m, lay, hint, st);
    if (n <= 320) return launch_variant<T, 320>(m, lay, hint, st);
    if (n <= 704) return launch_variant<T, 704>(m, lay, hint, st);
    return launch_variant<T, 1024>(m, lay, hint, st);
}

template <class T> int ensure(T** p, size_t* cap, size_t n) {
    if (n <= *cap) return G3CUDA_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    if (cudaMalloc(p, n * sizeof(T)) != cudaSuccess) return fail(G3CUDA_ENOMEM, "cudaMalloc of %zu bytes failed", n * sizeof(T));
    *cap = n;
    return G3CUDA_OK;
}

int check_spec(const g3cuda_spec* spec) {
    if (!spec || !spec->ints || !spec->reals) return fail(G3CUDA_EINVAL, "null spec");
    if (spec->version != G3CUDA_SPEC_VERSION || spec->n_ints < G3H_LEN || spec->ints[G3H_VERSION] != G3CUDA_SPEC_VERSION)
        return fail(G3CUDA_EINVAL, "spec version mismatch (library built for %d)", G3CUDA_SPEC_VERSION);
    const int32_t* I = spec->ints;
    auto in_i = [&](int64_t off, int64_t n) { return off >= 0 && off + n <= spec->n_ints; };
    if (!in_i(I[G3H_STOCK_OFF], (int64_t)I[G3H_NSTOCK] * G3S_LEN) || !in_i(I[G3H_COMP_OFF], (int64_t)I[G3H_NCOMP] * G3C_LEN) ||
        !in_i(I[G3H_PP_OFF], (int64_t)I[G3H_NPP] * G3PP_LEN) || !in_i(I[G3H_PRED_OFF], (int64_t)I[G3H_NPRED] * G3P_LEN) ||
        !in_i(I[G3H_SLOT_OFF], 2LL * I[G3H_NSLOT]) || !in_i(I[G3H_XBUF_OFF], (int64_t)I[G3H_NXBUF] * G3X_LEN))
        return fail(G3CUDA_EINVAL, "spec section offsets out of range");
    if (I[G3H_NSTOCK] < 1 || I[G3H_NSTOCK] > g3::MAXS) return fail(G3CUDA_EUNSUPP, "number of stocks %d outside 1..%d", I[G3H_NSTOCK], g3::MAXS);
    if (I[G3H_NCOMP] > g3::MAXC) return fail(G3CUDA_EUNSUPP, "more than %d likelihood components", g3::MAXC);
    if (I[G3H_NXBUF] > g3::MAXX) return fail(G3CUDA_EUNSUPP, "more than %d collect vectors", g3::MAXX);

    if (I[G3H_NSTEPS] > 31) return fail(G3CUDA_EUNSUPP, "more than 31 steps per year");
    return G3CUDA_OK;
}

}  // namespace

// DFMA-chain microbenchmark: the fp64 roofline denominator (MEASURED_PEAKS.json has no fp64 entry).
__global__ void fp64_peak_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void math_probe_kernel(int which, const double* x, const double* y, long long n, double* out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (which == 0) out[i] = g3::lsa(x[i], y[i]);
    else if (which == 1) out[i] = g3::avoid_zero(x[i]);
    else out[i] = g3::lsa(x[i] * -1e3, y[i] * -1e3) / -1e3;
}

extern "C" {

int g3cuda_math_probe(int device, int which, const double* x, const double* y, int64_t n, double* out) {
    if (!x || !out || n < 0 || which < 0 || which > 2 || (which != 1 && !y)) return fail(G3CUDA_EINVAL, "bad argument");
    if (n == 0) return G3CUDA_OK;
    CUDA_TRY(cudaSetDevice(device));
    double *dx = nullptr, *dy = nullptr, *dout = nullptr;
    const size_t bytes = sizeof(double) * (size_t)n;
    if (cudaMalloc(&dx, bytes) != cudaSuccess || cudaMalloc(&dy, bytes) != cudaSuccess || cudaMalloc(&dout, bytes) != cudaSuccess) {
        cudaFree(dx); cudaFree(dy); cudaFree(dout); cudaGetLastError();
        return fail(G3CUDA_ENOMEM, "cudaMalloc failed");
    }
    cudaMemcpy(dx, x, bytes, cudaMemcpyHostToDevice);
    if (y) cudaMemcpy(dy, y, bytes, cudaMemcpyHostToDevice); else cudaMemset(dy, 0, bytes);
    math_probe_kernel<<<(unsigned)((n + 255) / 256), 256>>>(which, dx, dy, (long long)n, dout);
    cudaError_t e = cudaMemcpy(out, dout, bytes, cudaMemcpyDeviceToHost);
    cudaFree(dx); cudaFree(dy); cudaFree(dout);
    if (e != cudaSuccess) return fail(G3CUDA_ECUDA, "math probe failed: %s", cudaGetErrorString(e));
    return G3CUDA_OK;
}

double g3cuda_fp64_peak_tflops(int device) {
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return -1.0; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1.0;
    const int threads = 512, blocks = prop.multiProcessorCount * 4, iters = 20000;
    double* d = nullptr;
    if (cudaMalloc(&d, sizeof(double) * threads * blocks) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        fp64_peak_kernel<<<blocks, threads>>>(d, iters);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 64.0 * iters * (double)threads * blocks;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    return best;
}

const char* g3cuda_last_error(void) { return g_err.c_str(); }
int32_t g3cuda_abi_version(void) { return G3CUDA_SPEC_VERSION; }
int32_t g3cuda_n_par(const g3cuda_model* m) { return m ? m->base.NPAR : 0; }
int32_t g3cuda_n_nll(const g3cuda_model* m) { return m ? m->base.NNLL : 0; }
int32_t g3cuda_n_steps(const g3cuda_model* m) { return m ? m->base.NT : 0; }
int64_t g3cuda_report_len(const g3cuda_model* m) { return m ? m->report_len : 0; }
int64_t g3cuda_launch_count(const g3cuda_model* m) { return m ? m->launches : 0; }

void g3cuda_model_free(g3cuda_model* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    for (void* p : m->owned) cudaFree(p);
    for (void* p : {(void*)m->d_par, (void*)m->d_nll, (void*)m->d_comp, (void*)m->d_grad, (void*)m->d_tidx, (void*)m->d_report, (void*)m->d_out, m->t_ws})
        if (p) cudaFree(p);
    if (m->ev0) cudaEventDestroy(m->ev0);
    if (m->ev1) cudaEventDestroy(m->ev1);
    if (m->ws_free) cudaEventDestroy(m->ws_free);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
}

int g3cuda_model_create(const g3cuda_spec* spec, int device, g3cuda_model** out) {
    using namespace g3;
    if (!out) return fail(G3CUDA_EINVAL, "null out pointer");
    *out = nullptr;
    int rc = check_spec(spec);
    if (rc) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1 || device < 0 || device >= ndev) {
        cudaGetLastError();
        return fail(G3CUDA_ENODEVICE, "no usable CUDA device (requested %d, found %d); this library has no CPU fallback", device, ndev);
    }
    CUDA_TRY(cudaSetDevice(device));
    g3cuda_model* m = new g3cuda_model();
    m->device = device;
    m->I.assign(spec->ints, spec->ints + spec->n_ints);
    m->R.assign(spec->reals, spec->reals + spec->n_reals);
    const int32_t* I = m->I.data();
    const double* R = m->R.data();
    KArgs& A = m->base;
    A.NC = I[G3H_NCELL]; A.NS = I[G3H_NSTOCK]; A.NPP = I[G3H_NPP]; A.NCOMP = I[G3H_NCOMP]; A.NX = I[G3H_NXBUF];
    A.NSLOT = I[G3H_NSLOT]; A.NPAR = I[G3H_NPAR]; A.NT = I[G3H_NT]; A.NSTEPS = I[G3H_NSTEPS]; A.NNLL = I[G3H_NNLL];
    A.NBOUND = I[G3H_NBOUND]; A.maxL = I[G3H_MAXL];
    auto stock = [&](int s) { return I + I[G3H_STOCK_OFF] + s * G3S_LEN; };
    auto bail = [&](int code) { g3cuda_model_free(m); return code; };

    // ---- per-cell tables
    int maxnarea = 1;
    std::vector<int4> cell(A.NC);
    std::vector<int32_t> inc_src(A.NC, -1), inc_slot(A.NC, -1), inc_mask(A.NC, 0), age_src(A.NC, -1), age_slot(A.NC, -1);
    std::vector<int32_t> rem_off(A.NC, 0), rem_slots;
    A.any_const_mat = 0;
    for (int s = 0; s < A.NS; s++) {
        const int32_t* St = stock(s);
        int L = St[G3S_L], nage = St[G3S_NAGE], narea = St[G3S_NAREA], c0 = St[G3S_CELL_OFF];
        maxnarea = std::max(maxnarea, narea);
        if (St[G3S_N_RENEW] > MAXRN) return bail(fail(G3CUDA_EUNSUPP, "more than %d renewal actions on one stock", MAXRN));
        if (St[G3S_MIGRATE_OFF] >= 0) m->only_thread = true;        // g3a_migrate: thread kernel only
        if (St[G3S_GROW_N] >= 0 && St[G3S_MAT_KIND] == G3MAT_CONSTANT && St[G3S_N_OUT] > 0) A.any_const_mat = 1;
        if (St[G3S_GROW_N] >= 0 && St[G3S_MAT_KIND] == G3MAT_CONTINUOUS && St[G3S_N_OUT] > 0 && I[St[G3S_MAT_OFF] + 2] >= 0)
            return bail(fail(G3CUDA_EUNSUPP, "g3a_mature_continuous with a beta/a50 (age) term is not supported yet"));
        for (int r = 0; r < narea; r++) for (int a = 0; a < nage; a++) for (int l = 0; l < L; l++)
            cell[c0 + (r * nage + a) * L + l] = make_int4(s, l, r * nage + a, r);
        int ncell = L * nage * narea;
        // maturation: destination cells learn their source; sources learn which ratios have a destination
        bool trans = St[G3S_GROW_N] >= 0 && St[G3S_N_OUT] > 0;
        for (int i = 0; i < ncell; i++) {
            rem_off[c0 + i] = (int)rem_slots.size();
            if (trans) for (int o = 0; o < St[G3S_N_OUT]; o++) {
                int d = I[St[G3S_OUT_MAP_OFF] + o * ncell + i];
                if (d < 0) continue;
                int slot = I[St[G3S_OUT_OFF] + 2 * o + 1];
                if (inc_src[d] >= 0) return bail(fail(G3CUDA_EUNSUPP, "a cell receives maturing fish from more than one source"));
                inc_src[d] = c0 + i; inc_slot[d] = slot; inc_mask[d] = St[G3S_TRANS_MASK];
                rem_slots.push_back(slot);
            }
            rem_slots.push_back(-1);
        }
        if (St[G3S_AGE_ENABLE]) for (int o = 0; o < St[G3S_AGE_N_OUT]; o++) for (int r = 0; r < narea; r++) for (int l = 0; l < L; l++) {
            int d = I[St[G3S_AGE_MAP_OFF] + (o * narea + r) * L + l];
            if (d < 0) continue;
            if (age_src[d] >= 0) return bail(fail(G3CUDA_EUNSUPP, "a cell receives ageing fish from more than one source"));
            age_src[d] = c0 + (r * nage + nage - 1) * L + l;
            age_slot[d] = I[St[G3S_AGE_OUT_OFF] + 2 * o + 1];
        }
    }
    A.maxnarea = maxnarea;
    for (int st = 0; st < 32; st++) {
        A.trans_any[st] = 0;
        for (int s = 0; s < A.NS; s++) {
            const int32_t* St = stock(s);
            if (St[G3S_GROW_N] >= 0 && St[G3S_N_OUT] > 0 && ((St[G3S_TRANS_MASK] >> st) & 1)) A.trans_any[st] = 1;
        }
    }
    // ---- predation tables
    std::vector<int32_t> tsid((size_t)std::max(A.NPP, 1) * maxnarea, -1), eoff(tsid.size(), 0), estr(tsid.size(), 0);
    std::vector<int> ts_of_pred_area;   // accumulator id per (pred, pred area idx)
    int nts = 0;
    std::vector<int> pred_ts_base(I[G3H_NPRED], 0);
    for (int p = 0; p < I[G3H_NPRED]; p++) {
        const int32_t* P = I + I[G3H_PRED_OFF] + p * G3P_LEN;
        if (P[G3P_KIND] == G3PRED_STOCK) { m->only_thread = true; pred_ts_base[p] = -1; continue; }     // stock predator: thread kernel only
        if (P[G3P_KIND] == G3PRED_NUMBERFLEET || P[G3P_KIND] == G3PRED_LINEARFLEET || P[G3P_KIND] == G3PRED_EFFORTFLEET)
            m->only_thread = true;      // the CTA / warp kernels know the totalfleet form only
        else if (P[G3P_KIND] != G3PRED_TOTALFLEET) return bail(fail(G3CUDA_EUNSUPP, "predator kind %d is not supported", P[G3P_KIND]));
        pred_ts_base[p] = nts; nts += P[G3P_NAREA];
    }
    if (nts > MAXTS) return bail(fail(G3CUDA_EUNSUPP, "more than %d (predator, area) pairs", MAXTS));
    A.NTS = nts;
    std::vector<std::vector<int>> pp_of_stock(A.NS);
    for (int q = 0; q < A.NPP; q++) {
        const int32_t* Q = I + I[G3H_PP_OFF] + q * G3PP_LEN;
        const int32_t* P = I + I[G3H_PRED_OFF] + Q[G3PP_PRED] * G3P_LEN;
        const int32_t* St = stock(Q[G3PP_PREY]);
        const int sk = Q[G3PP_SUIT_KIND];
        if (sk == G3SUIT_ANDERSEN || sk == G3SUIT_EXPONENTIAL || sk == G3SUIT_RICHARDS) {
            if (P[G3P_KIND] != G3PRED_STOCK) return bail(fail(G3CUDA_EUNSUPP, "suitability kind %d reads predator_length: it needs a stock predator", sk));
        } else if (sk < 0 || sk > G3SUIT_RICHARDS)
            return bail(fail(G3CUDA_EUNSUPP, "suitability kind %d is not supported", sk));
        if (sk != G3SUIT_EXPL50 && sk != G3SUIT_CONSTANT) m->only_thread = true;     // the CTA / warp kernels know two forms only
        pp_of_stock[Q[G3PP_PREY]].push_back(q);
        for (int r = 0; r < St[G3S_NAREA]; r++) {
            int area = I[St[G3S_AREA_OFF] + r];
            for (int k = 0; k < P[G3P_NAREA]; k++) if (I[P[G3P_AREA_OFF] + k] == area) {
                if (P[G3P_KIND] == G3PRED_STOCK) { tsid[(size_t)q * maxnarea + r] = k; continue; }     // predator area index
                tsid[(size_t)q * maxnarea + r] = pred_ts_base[Q[G3PP_PRED]] + k;
                eoff[(size_t)q * maxnarea + r] = P[G3P_E_ROFF] + k;
                estr[(size_t)q * maxnarea + r] = P[G3P_NAREA];
            }
        }
    }
    int ppo = 0;
    for (int s = 0; s < A.NS; s++) {
        A.stock_pp_off[s] = ppo;
        if ((int)pp_of_stock[s].size() > MAXQ) return bail(fail(G3CUDA_EUNSUPP, "more than %d predators on one stock", MAXQ));
        for (int q : pp_of_stock[s]) A.stock_pp[ppo++] = q;
        A.us_flag[s] = 0;
    }
    A.stock_pp_off[A.NS] = ppo;
    for (int i = 0; i < I[G3H_US_N]; i++) A.us_flag[I[I[G3H_US_OFF] + i]] = 1;

    // ---- warp-per-evaluation kernel: tiling, distinct step lengths, compact hand-over tables
    m->warp_ok = A.NPP <= 32;
    {
        A.w_ndt = 0;
        for (int st = 0; st < A.NSTEPS; st++) {
            int len = I[I[G3H_STEPLEN_OFF] + st], k = 0;
            while (k < A.w_ndt && A.w_dt_len[k] != len) k++;
            if (k == A.w_ndt) { if (A.w_ndt < 4) A.w_dt_len[A.w_ndt++] = len; else { m->warp_ok = false; k = 0; } }
            A.w_step_idt[st] = k;
        }
        std::vector<int32_t> tr_of_cell(A.NC, -1), inc_tr(A.NC, -1);
        int ntr = 0, nmv = 0;
        std::vector<int> mv_off(A.NS, 0);
        for (int s = 0; s < A.NS; s++) {
            const int32_t* St = stock(s);
            int L = St[G3S_L], ncell = L * St[G3S_NAGE] * St[G3S_NAREA];
            if (L > 32) { m->warp_ok = false; A.w_G[s] = 1; A.w_ntile[s] = 0; }
            else { A.w_G[s] = 32 / L; A.w_ntile[s] = (St[G3S_NAGE] * St[G3S_NAREA] + A.w_G[s] - 1) / A.w_G[s]; }
            if (St[G3S_GROW_N] >= 0 && St[G3S_N_OUT] > 0) { for (int i = 0; i < ncell; i++) tr_of_cell[St[G3S_CELL_OFF] + i] = ntr + i; ntr += ncell; }
            mv_off[s] = nmv;
            if (St[G3S_AGE_ENABLE] && St[G3S_AGE_N_OUT] > 0) nmv += 2 * St[G3S_NAREA] * L;
            A.w_inc_steps[s] = 0;
        }
        for (int c = 0; c < A.NC; c++) if (inc_src[c] >= 0) {
            inc_tr[c] = tr_of_cell[inc_src[c]];
            A.w_inc_steps[cell[c].x] |= inc_mask[c];
        }
        std::vector<int4> agedst;
        for (int c = 0; c < A.NC; c++) if (age_src[c] >= 0) {
            int src = age_src[c], ss = cell[src].x;
            const int32_t* St = stock(ss);
            int L = St[G3S_L], i = cell[src].w * L + cell[src].y;       // (area idx, l) of the oldest column
            agedst.push_back(make_int4(c, mv_off[ss] + i, mv_off[ss] + St[G3S_NAREA] * L + i, age_slot[c]));
        }
        A.w_n_agedst = (int)agedst.size();
        int4* d_agedst = dev_copy(agedst); int32_t* d_inc_tr = dev_copy(inc_tr);
        if (!d_agedst || !d_inc_tr) return bail(fail(G3CUDA_ENOMEM, "device allocation failed"));
        m->owned.push_back(d_agedst); m->owned.push_back(d_inc_tr);
        A.w_agedst = d_agedst; A.w_inc_tr = d_inc_tr;
        // accumulators of (predator, area): where their landings live
        std::vector<unsigned char> pact(A.NT, 0);
        for (int p = 0; p < I[G3H_NPRED]; p++) {
            const int32_t* P = I + I[G3H_PRED_OFF] + p * G3P_LEN;
            if (P[G3P_KIND] == G3PRED_STOCK) { std::fill(pact.begin(), pact.end(), 1); continue; }    // a stock predator eats every step
            for (int k = 0; k < P[G3P_NAREA]; k++) {
                A.w_ts_eoff[pred_ts_base[p] + k] = P[G3P_E_ROFF] + k;
                A.w_ts_estr[pred_ts_base[p] + k] = P[G3P_NAREA];
                A.w_ts_kind[pred_ts_base[p] + k] = P[G3P_KIND];
                for (int t = 0; t < A.NT; t++) if (R[P[G3P_E_ROFF] + (size_t)t * P[G3P_NAREA] + k] != 0.0) pact[t] = 1;
            }
        }
        unsigned char* d_pact = dev_copy(pact);
        if (!d_pact) return bail(fail(G3CUDA_ENOMEM, "device allocation failed"));
        m->owned.push_back(d_pact); A.w_pred_active = d_pact;
        for (int x = 0; x < A.NX; x++) {
            const int32_t* X = I + I[G3H_XBUF_OFF] + x * G3X_LEN;
            A.w_x_kind[x] = X[G3X_KIND]; A.w_x_unit[x] = X[G3X_UNIT]; A.w_x_ppmask[x] = 0;
            for (int j = 0; j < X[G3X_N_PP]; j++) { int q = I[X[G3X_PP_OFF] + j]; if (q < 32) A.w_x_ppmask[x] |= 1u << q; }
        }
    }

    // ---- thread-per-evaluation kernel: workspace layout, cell-oriented collect CSR, ordering check
    {
        m->thread_ok = A.NPP <= 32 && A.w_ndt <= 4 && A.NCOMP <= 32 && I[G3H_NPRED] <= MAXS;
        for (int s = 0; s < A.NS; s++) {
            const int32_t* St = stock(s);
            A.t_mig_steps[s] = 0;
            if (St[G3S_MIGRATE_OFF] < 0) continue;
            int NR = St[G3S_NAREA];
            if (NR > MAXMIG) m->thread_ok = false;
            for (int st = 0; st < A.NSTEPS; st++) for (int i = 0; i < NR * NR; i++)
                if (I[St[G3S_MIGRATE_OFF] + st * NR * NR + i] >= 0) A.t_mig_steps[s] |= 1 << st;
        }
        m->t_maxn = 0;
        for (int s = 0; s < A.NS; s++) m->t_maxn = std::max(m->t_maxn, stock(s)[G3S_GROW_N]);
        if (m->t_maxn > 15) m->thread_ok = false;
        // the fused pass hands maturing fish over within one sweep: sources must come before destinations
        for (int c = 0; c < A.NC; c++) if (inc_src[c] >= 0 && cell[inc_src[c]].x >= cell[c].x) m->thread_ok = false;
        m->t_entries = make_tlayout(m, A);
        // hand-over per COLUMN: every cell of a destination column must be fed by the same length of one source column
        {
            std::vector<int4> colinc;
            std::vector<int32_t> tr_of_cell(A.NC, -1);
            int ntr = 0;
            for (int s = 0; s < A.NS; s++) {
                const int32_t* St = stock(s);
                int ncell = St[G3S_L] * St[G3S_NAGE] * St[G3S_NAREA];
                if (St[G3S_GROW_N] >= 0 && St[G3S_N_OUT] > 0) { for (int i = 0; i < ncell; i++) tr_of_cell[St[G3S_CELL_OFF] + i] = ntr + i; ntr += ncell; }
            }
            for (int s = 0; s < A.NS; s++) {
                const int32_t* St = stock(s);
                int L = St[G3S_L];
                for (int col = 0; col < St[G3S_NAGE] * St[G3S_NAREA]; col++) {
                    int c00 = St[G3S_CELL_OFF] + col * L;
                    int4 e = make_int4(-1, 0, 0, 0);
                    if (inc_src[c00] >= 0) e = make_int4(tr_of_cell[inc_src[c00]], inc_slot[c00], inc_mask[c00], 0);
                    for (int l = 0; l < L; l++) {
                        int c = c00 + l;
                        bool ok = inc_src[c] < 0 ? e.x < 0 : (e.x >= 0 && tr_of_cell[inc_src[c]] == e.x + l && inc_slot[c] == e.y && inc_mask[c] == e.z);
                        if (!ok) m->thread_ok = false;
                        if (rem_off[c] != rem_off[c00] && l > 0) {       // remainder ratios must be the same down a column
                            int a = rem_off[c], b = rem_off[c00];
                            while (rem_slots[a] >= 0 && rem_slots[a] == rem_slots[b]) { a++; b++; }
                            if (rem_slots[a] != rem_slots[b]) m->thread_ok = false;
                        }
                    }
                    colinc.push_back(e);
                }
            }
            int4* d_ci = dev_copy(colinc);
            if (!d_ci) return bail(fail(G3CUDA_ENOMEM, "device allocation failed"));
            m->owned.push_back(d_ci); A.t_colinc = d_ci;
            m->t_entries = make_tlayout(m, A);
        }
        // one renewal per (column, step) is what the fused pass applies
        for (int s = 0; s < A.NS; s++) {
            const int32_t* St = stock(s);
            for (int k = 0; k < St[G3S_N_RENEW]; k++) for (int k2 = k + 1; k2 < St[G3S_N_RENEW]; k2++) {
                const int32_t* Ra = I + St[G3S_RENEW_OFF] + k * G3R_LEN; const int32_t* Rb = I + St[G3S_RENEW_OFF] + k2 * G3R_LEN;
                if (Ra[G3R_AGE_IDX] == Rb[G3R_AGE_IDX] && (Ra[G3R_RUN_STEP] == 0 || Rb[G3R_RUN_STEP] == 0 || Ra[G3R_RUN_STEP] == Rb[G3R_RUN_STEP]))
                    m->thread_ok = false;
            }
        }
    }

    // ---- observation terms that do not depend on parameters
    std::vector<double> obsprop, obsconst;
    int maxnd = 0, maxnt = 0;
    for (int c = 0; c < A.NCOMP; c++) {
        const int32_t* C = I + I[G3H_COMP_OFF] + c * G3C_LEN;
        int nd = C[G3C_N_DEST], nt = C[G3C_N_TIMES], nb = C[G3C_N_BINS], nsl = C[G3C_N_SLICES], nag = C[G3C_N_AREAG];
        maxnd = std::max(maxnd, nd);
        A.obs_off[c] = (int)obsprop.size(); A.obsconst_off[c] = (int)obsconst.size();
        const double* obs = R + C[G3C_OBS_ROFF];
        bool si = C[G3C_FUNC] == G3F_SI_LOG || C[G3C_FUNC] == G3F_SI_LINEAR;
        if (si) maxnt = std::max(maxnt, nt);
        if (C[G3C_FUNC] == G3F_SUMSQLOGRATIOS) m->only_thread = true;       // thread kernel only
        else if (C[G3C_FUNC] < 0 || C[G3C_FUNC] > G3F_SUMSQLOGRATIOS) return bail(fail(G3CUDA_EUNSUPP, "likelihood function %d is not supported", C[G3C_FUNC]));
        if (C[G3C_MODE] == 1 && nd > 32) return bail(fail(G3CUDA_EUNSUPP, "reduce-mode component with %d destinations", nd));
        for (int t = 0; t < nt; t++) {
            double cst = 0.0;
            for (int ag = 0; ag < nag; ag++) {
                double tot = 0.0;
                for (int i = 0; i < nsl * nb; i++) tot += obs[(size_t)t * nd + ag * nsl * nb + i];
                double az = h_avoid_zero(tot);
                for (int i = 0; i < nsl * nb; i++) {
                    double o = obs[(size_t)t * nd + ag * nsl * nb + i];
                    double v = o;
                    if (C[G3C_FUNC] == G3F_SUMOFSQUARES) v = o / az;                // R/likelihood_distribution.R:12-27
                    else if (C[G3C_FUNC] == G3F_SI_LOG) v = std::log(h_avoid_zero(o));  // R/likelihood_distribution.R:86
                    else if (C[G3C_FUNC] == G3F_SUMSQLOGRATIOS) v = std::log(o + R[C[G3C_PARAM_ROFF]]);     // R/likelihood_distribution.R:131
                    obsprop.push_back(v);
                }
                if (C[G3C_FUNC] == G3F_MULTINOMIAL)
                    for (int sl = 0; sl < nsl; sl++) {      // sum(lgamma(1 + data)) - lgamma(1 + sum(data))
                        double so = 0, slg = 0;
                        for (int b = 0; b < nb; b++) { double o = obs[(size_t)t * nd + (ag * nsl + sl) * nb + b]; so += o; slg += std::lgamma(1 + o); }
                        cst += slg - std::lgamma(1 + so);
                    }
            }
            obsconst.push_back(cst);
        }
    }
    // ---- report layout (include/g3cuda.h)
    int64_t pos = 1 + (int64_t)A.NNLL * A.NT;
    for (int s = 0; s < A.NS; s++) {
        const int32_t* St = stock(s);
        A.rep_dstart[s] = pos;
        pos += 2LL * A.NT * St[G3S_L] * St[G3S_NAGE] * St[G3S_NAREA];
    }
    for (int c = 0; c < A.NCOMP; c++) {
        const int32_t* C = I + I[G3H_COMP_OFF] + c * G3C_LEN;
        A.rep_model[c] = pos; pos += (int64_t)C[G3C_N_TIMES] * C[G3C_N_DEST];
    }
    A.rep_params = pos; pos += 2LL * A.NCOMP;
    m->rep_detail = pos;
    for (int s = 0; s < A.NS; s++) {
        const int32_t* St = stock(s);
        A.rep_renew[s] = pos;
        pos += (int64_t)A.NT * St[G3S_L] * St[G3S_NAGE] * St[G3S_NAREA];
    }
    for (int q = 0; q < A.NPP && q < 32; q++) {
        const int32_t* Q = I + I[G3H_PP_OFF] + q * G3PP_LEN;
        const int32_t* P = I + I[G3H_PRED_OFF] + Q[G3PP_PRED] * G3P_LEN;
        const int32_t* St = stock(Q[G3PP_PREY]);
        const int64_t PL = P[G3P_KIND] == G3PRED_STOCK ? stock(P[G3P_STOCK])[G3S_L] : 1;
        A.rep_pp[q] = pos;
        pos += PL * St[G3S_L] + 2LL * A.NT * St[G3S_L] * St[G3S_NAGE] * St[G3S_NAREA];
    }
    m->report_len = pos;

    int need = std::max({A.NC, maxnd, maxnt, A.maxL, 32});
    m->nthreads = (need + 31) / 32 * 32;
    // (more than 1024 cells: the CTA kernel is unavailable; the thread kernel has no such limit)

    // ---- upload
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    m->num_sms = prop.multiProcessorCount;
    m->smem_optin = prop.sharedMemPerBlockOptin;
    auto up = [&](auto& vec) { auto* d = dev_copy(vec); if (d) m->owned.push_back((void*)d); return d; };
    A.I = up(m->I); A.R = up(m->R); A.cell = up(cell);
    A.inc_src = up(inc_src); A.inc_slot = up(inc_slot); A.inc_mask = up(inc_mask);
    A.age_src = up(age_src); A.age_slot = up(age_slot); A.rem_off = up(rem_off); A.rem_slots = up(rem_slots);
    A.pp_tsid = up(tsid); A.pp_eoff = up(eoff); A.pp_estride = up(estr);
    A.obsprop = up(obsprop); A.obsconst = up(obsconst);
    if (!A.I || !A.R || !A.cell || !A.inc_src || !A.inc_slot || !A.inc_mask || !A.age_src || !A.age_slot || !A.rem_off ||
        !A.rem_slots || !A.pp_tsid || !A.pp_eoff || !A.pp_estride || !A.obsprop || !A.obsconst)
        return bail(fail(G3CUDA_ENOMEM, "device allocation failed: %s", cudaGetErrorString(cudaGetLastError())));
    if (cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&m->ev0) != cudaSuccess || cudaEventCreate(&m->ev1) != cudaSuccess)
        return bail(fail(G3CUDA_ECUDA, "stream/event creation failed"));
    if (m->only_thread) m->warp_ok = false;
    m->cta_ok_double = !m->only_thread && m->nthreads <= 1024 && make_layout(m, sizeof(double)).bytes <= (size_t)prop.sharedMemPerBlockOptin;
    m->cta_ok_dual = !m->only_thread && m->nthreads <= 1024 && make_layout(m, sizeof(DualT)).bytes <= (size_t)prop.sharedMemPerBlockOptin;
    if (!m->thread_ok && !m->cta_ok_double)
        return bail(fail(G3CUDA_EUNSUPP, "model fits neither evaluation kernel (%d cells, %zu bytes of shared memory per evaluation, limit %zu)",
                         A.NC, make_layout(m, sizeof(double)).bytes, (size_t)prop.sharedMemPerBlockOptin));
    *out = m;
    return G3CUDA_OK;
}

int g3cuda_eval_batch_device(g3cuda_model* m, const double* d_par, int64_t B, const int32_t* d_tangent_idx, int32_t K,
                             double* d_nll, double* d_comp, double* d_grad, void* stream) {
    if (!m || !d_par || !d_nll || B < 0) return fail(G3CUDA_EINVAL, "bad argument");
    if (B == 0) return G3CUDA_OK;
    CUDA_TRY(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    g3::KArgs call{};
    call.par = d_par; call.B = B; call.nll = d_nll; call.comp = d_comp; call.report = nullptr;
    m->timed = true;
    CUDA_TRY(cudaEventRecord(m->ev0, st));
    int rc;
    if (d_grad && K > 0) {
        if (!d_tangent_idx) return fail(G3CUDA_EINVAL, "grad requested without tangent_idx");
        const long long ntiles = (K + NTAN - 1) / NTAN;
        call.tangent_idx = nullptr; call.K = 0; call.grad = nullptr;
        g3::KArgs g = call;
        g.tangent_idx = d_tangent_idx; g.K = K; g.grad = d_grad;
        // The thread kernel's dual pass carries the value: its first direction tile writes nll and the component
        // sums, so no separate value launch is needed (it cost 6 % of a C3 step). The CTA kernel's dual pass
        // does not write the components: there the value pass stays.
        const bool dual_on_thread = m->thread_ok && (m->force_kernel == 0 || m->force_kernel == 3) &&
                                    (m->force_kernel == 3 || B * ntiles >= 16384 || !m->cta_ok_dual);
        if (!dual_on_thread) {
            rc = launch<double>(m, call, B, st);
            if (rc) return rc;
            rc = ensure(&m->d_nll, &m->cap_nll, (size_t)B);     // (its tile 0 writes nll too: into scratch)
            if (rc) return rc;
            g.nll = m->d_nll; g.comp = nullptr;
        }
        rc = launch<DualT>(m, g, B * ntiles, st);
    } else {
        call.tangent_idx = nullptr; call.K = 0; call.grad = nullptr;
        rc = launch<double>(m, call, B, st);
    }
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(m->ev1, st));
    return G3CUDA_OK;
}

int g3cuda_eval_batch(g3cuda_model* m, const double* par, int64_t B, const int32_t* tangent_idx, int32_t K,
                      double* nll, double* nll_comp, double* grad) {
    if (!m || !par || !nll || B < 0) return fail(G3CUDA_EINVAL, "bad argument");
    if (B == 0) return G3CUDA_OK;
    CUDA_TRY(cudaSetDevice(m->device));
    const int P = m->base.NPAR, NN = m->base.NNLL;
    int rc;
    if ((rc = ensure(&m->d_par, &m->cap_par, (size_t)B * P))) return rc;
    if ((rc = ensure(&m->d_out, &m->cap_out, (size_t)B))) return rc;
    if (nll_comp && (rc = ensure(&m->d_comp, &m->cap_comp, (size_t)B * NN))) return rc;
    bool want_grad = grad && K > 0;
    if (want_grad) {
        if (!tangent_idx) return fail(G3CUDA_EINVAL, "grad requested without tangent_idx");
        if ((rc = ensure(&m->d_grad, &m->cap_grad, (size_t)B * K)) || (rc = ensure(&m->d_tidx, &m->cap_tidx, (size_t)K))) return rc;
        CUDA_TRY(cudaMemcpyAsync(m->d_tidx, tangent_idx, sizeof(int32_t) * K, cudaMemcpyHostToDevice, m->stream));
    }
    CUDA_TRY(cudaMemcpyAsync(m->d_par, par, sizeof(double) * B * P, cudaMemcpyHostToDevice, m->stream));
    rc = g3cuda_eval_batch_device(m, m->d_par, B, want_grad ? m->d_tidx : nullptr, want_grad ? K : 0, m->d_out,
                                  nll_comp ? m->d_comp : nullptr, want_grad ? m->d_grad : nullptr, m->stream);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(nll, m->d_out, sizeof(double) * B, cudaMemcpyDeviceToHost, m->stream));
    if (nll_comp) CUDA_TRY(cudaMemcpyAsync(nll_comp, m->d_comp, sizeof(double) * B * NN, cudaMemcpyDeviceToHost, m->stream));
    if (want_grad) CUDA_TRY(cudaMemcpyAsync(grad, m->d_grad, sizeof(double) * B * K, cudaMemcpyDeviceToHost, m->stream));
    cudaError_t e = cudaStreamSynchronize(m->stream);
    if (e != cudaSuccess) return fail(G3CUDA_ECUDA, "evaluation failed: %s", cudaGetErrorString(e));
    return G3CUDA_OK;
}

int g3cuda_set_kernel(g3cuda_model* m, int which) {
    if (!m || which < 0 || which > 3) return fail(G3CUDA_EINVAL, "bad argument");
    if (which == 3 && !m->thread_ok) return fail(G3CUDA_EUNSUPP, "model does not fit the thread-per-evaluation kernel");
    if (which == 2 && !m->warp_ok) return fail(G3CUDA_EUNSUPP, "model does not fit the warp-per-evaluation kernel");
    m->force_kernel = which;
    return G3CUDA_OK;
}

float g3cuda_last_kernel_ms(g3cuda_model* m) {
    if (!m || !m->timed) return -1.0f;
    float ms = -1.0f;
    if (cudaEventSynchronize(m->ev1) != cudaSuccess) return -1.0f;
    if (cudaEventElapsedTime(&ms, m->ev0, m->ev1) != cudaSuccess) return -1.0f;
    return ms;
}

int g3cuda_report(g3cuda_model* m, const double* par, double* out) {
    if (!m || !par || !out) return fail(G3CUDA_EINVAL, "bad argument");
    CUDA_TRY(cudaSetDevice(m->device));
    int rc;
    size_t cap = 0; double* d_rep = nullptr; double* d_nll = nullptr; size_t cap2 = 0;
    if ((rc = ensure(&m->d_par, &m->cap_par, (size_t)m->base.NPAR))) return rc;
    if ((rc = ensure(&d_rep, &cap, (size_t)m->report_len))) return rc;
    if ((rc = ensure(&d_nll, &cap2, 1))) { cudaFree(d_rep); return rc; }
    cudaMemsetAsync(d_rep, 0, sizeof(double) * m->report_len, m->stream);
    cudaMemcpyAsync(m->d_par, par, sizeof(double) * m->base.NPAR, cudaMemcpyHostToDevice, m->stream);
    g3::KArgs call{};
    call.par = m->d_par; call.B = 1; call.nll = d_nll; call.report = d_rep;
    rc = launch<double>(m, call, 1, m->stream);
    if (!rc) {
        cudaMemcpyAsync(out, d_rep, sizeof(double) * m->report_len, cudaMemcpyDeviceToHost, m->stream);
        cudaError_t e = cudaStreamSynchronize(m->stream);
        if (e != cudaSuccess) rc = fail(G3CUDA_ECUDA, "report failed: %s", cudaGetErrorString(e));
        const int32_t* I = m->I.data();
        for (int c = 0; c < m->base.NCOMP && !rc; c++) {     // regression parameters exist for survey indices only
            int f = I[I[G3H_COMP_OFF] + c * G3C_LEN + G3C_FUNC];
            if (f != G3F_SI_LOG && f != G3F_SI_LINEAR) out[m->base.rep_params + 2 * c] = out[m->base.rep_params + 2 * c + 1] = std::nan("");
        }
        // the detail_* / suit_*__report section is written by the thread kernel only
        if (!rc && !(m->thread_ok && (m->force_kernel == 0 || m->force_kernel == 3)))
            for (int64_t i = m->rep_detail; i < m->report_len; i++) out[i] = std::nan("");
    }
    cudaFree(d_rep); cudaFree(d_nll);
    return rc;
}

}  // extern "C"
