// g3cuda.cu -- C ABI (include/g3cuda.h) over the evaluation kernels: spec validation, the planners' device
// tables, workspaces, launches, the host <-> device copies of the host-buffer entry point.
// No CPU evaluation path exists in this library: without a CUDA device every compute entry point returns
// G3CUDA_ENODEVICE. The kernels themselves are compiled in the instantiation units g3_inst_*.cu (g3_launch.h).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "../../include/g3cuda.h"
#include "g3_kargs.cuh"
#include "g3_launch.h"
#include "g3_ntan.h"
#include "g3_tile_plan.h"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}
#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(G3CUDA_ECUDA, "%s: %s", #x, cudaGetErrorString(e_)); } while (0)

constexpr int NTAN = G3_NTAN;           // tangent directions per gradient pass
using DualT = g3::Dual<NTAN>;

using g3::TPB; using g3::TPB_BIG; using g3::WS_STRIDE;     // thread-kernel launch geometry (g3_kargs.cuh)

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
    int NPAR = 0, NNLL = 0, NT = 0, NCOMP = 0;
    std::vector<void*> owned;           // device allocations that live as long as the handle
    int num_sms = 0;
    size_t smem_optin = 0;
    int64_t report_len = 0, rep_params = 0;
    int64_t launches = 0;
    // per-call scratch (grow-only)
    double* d_par = nullptr; double* d_comp = nullptr; double* d_grad = nullptr; double* d_out = nullptr; double* d_nll2 = nullptr;
    int32_t* d_tidx = nullptr;
    size_t cap_par = 0, cap_comp = 0, cap_grad = 0, cap_tidx = 0, cap_out = 0, cap_nll2 = 0;
    double* h_stage = nullptr; size_t cap_stage = 0;       // pinned staging of the results
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_copy[2] = {nullptr, nullptr};
    bool timed = false;
    cudaEvent_t ws_free = nullptr;      // recorded after every launch that uses a workspace: the next one waits for it,
                                        // whichever stream it is on (the workspaces are per handle, not per stream)
    int force_kernel = 0;               // 0 = automatic, 3 = thread kernel, 4 = tile kernel
    int max_tiles = 0;                  // tuning hook: at most this many tiles in flight (0 = every slot of the device)
    int no_lean = 0;                    // set_kernel(5): the tile kernel's full instantiation even where a leaner one covers the model
    // ---- tile kernel (g3_tile.cuh): plan, device tables, per-CTA slabs
    g3t::TilePlan tplan;
    g3t::TArgs targs{};
    bool tile_ok = false;
    int32_t* d_tb = nullptr;            // the plan's tables in one block; re-uploaded when the warp count or the partition changes
    const unsigned char* d_pred_active = nullptr;
    double* t_gws = nullptr; size_t t_gws_bytes = 0;
    // ---- thread kernel (g3_thread_kernel.cuh)
    g3::KArgs base{};
    bool thread_ok = false;
    std::string thread_why;
    int t_entries = 0, t_maxn = 0;      // workspace entries per thread; widest growth band
    void* t_ws = nullptr; size_t t_ws_bytes = 0;
};

namespace {

int launch_ptr(g3cuda_model* m, const void* fn, unsigned grid, unsigned block, size_t smem, cudaStream_t st, void* args) {
    void* params[] = {args};
    if (!m->ws_free) CUDA_TRY(cudaEventCreateWithFlags(&m->ws_free, cudaEventDisableTiming));
    else CUDA_TRY(cudaStreamWaitEvent(st, m->ws_free, 0));
    CUDA_TRY(cudaLaunchKernel(fn, dim3(grid), dim3(block), params, smem, st));
    CUDA_TRY(cudaEventRecord(m->ws_free, st));
    m->launches++;
    return G3CUDA_OK;
}

template <class T> int ensure(T** p, size_t* cap, size_t n) {
    if (n <= *cap) return G3CUDA_OK;
    if (*p) { cudaDeviceSynchronize(); cudaFree(*p); }
    *p = nullptr; *cap = 0;
    if (cudaMalloc(p, n * sizeof(T)) != cudaSuccess) { cudaGetLastError(); return fail(G3CUDA_ENOMEM, "cudaMalloc of %zu bytes failed", n * sizeof(T)); }
    *cap = n;
    return G3CUDA_OK;
}

// ============================================================ tile kernel
int tile_upload_tb(g3cuda_model* m) {
    // the plan's record tables and index lists travel as one block (g3t::plan_pack_meta); re-uploaded when the warp count or
    // the partition changes
    if (m->d_tb) { CUDA_TRY(cudaDeviceSynchronize()); cudaFree(m->d_tb); m->d_tb = nullptr; }
    m->d_tb = dev_copy(m->tplan.meta);
    if (!m->d_tb) return fail(G3CUDA_ENOMEM, "device allocation failed");
    m->tplan.a.meta = m->d_tb;
    m->targs = m->tplan.a;              // scalars (W, n_g, offsets) change with the warp count
    return G3CUDA_OK;
}

// tiles in flight and lanes per tile for `work` items
struct TileGeom { const void* fn; size_t smem; int per_sm; long long slots; bool in_smem, meta_in_smem, use_h; };

template <class T>
int tile_geom(g3cuda_model* m, TileGeom* g) {
    constexpr bool dual = !std::is_same<T, double>::value;
    constexpr size_t cpe = (size_t)(g3::Tan<T>::n + 1) * g3t::NL;
    const size_t state = (size_t)m->tplan.n_s * cpe * sizeof(double);
    g->in_smem = state <= m->smem_optin;
    g->smem = g->in_smem ? state : 0;
    // the plan's tables beside the state, if they fit (else the kernel reads them from global memory)
    // ... and, for plain values, the hot per-evaluation tables the planner chose (g3t::plan_place_hot)
    const size_t meta_bytes = (size_t)m->targs.meta_n * sizeof(int32_t) + 16;
    const size_t hot_bytes = dual ? 0 : (size_t)m->targs.n_h * cpe * sizeof(double);
    g->use_h = hot_bytes > 0 && g->smem + hot_bytes + meta_bytes + 4096 <= m->smem_optin;
    if (g->use_h) g->smem += hot_bytes;
    g->meta_in_smem = g->smem + meta_bytes + 4096 <= m->smem_optin;
    if (g->meta_in_smem) g->smem += meta_bytes;
    const bool wide = m->tplan.maxn > 4;
    const int threads = (m->targs.W + 1) * 32;
    if (threads > 512 && (dual || wide)) return fail(G3CUDA_EUNSUPP, "more than 16 warps per tile exist for plain values and narrow growth bands only");
    const bool xf = m->targs.xf != 0;
    if (xf && threads > 512) return fail(G3CUDA_EUNSUPP, "more than 16 warps per tile: not for models with g3a_otherfood, g3a_spawn or prey_preferences");
    const int need = m->targs.need;     // features the model uses: the leanest instantiation that covers them (plain values, 16 warps)
    const bool lean_ok = !dual && !xf && threads <= 512 && !m->no_lean;
    if (xf) g->fn = dual ? (wide ? g3i::tile_x_g16(g->in_smem) : g3i::tile_x_g5(g->in_smem)) : (wide ? g3i::tile_x_d16(g->in_smem) : g3i::tile_x_d5(g->in_smem));
    else if (lean_ok && !wide && (need & 15) == 0) g->fn = g3i::tile_l15_d5(g->in_smem);
    else if (lean_ok && !wide && (need & 12) == 0) g->fn = g3i::tile_l12_d5(g->in_smem);
    else if (lean_ok && wide && (need & 15) == 0) g->fn = g3i::tile_l15_d16(g->in_smem);
    else if (lean_ok && wide && (need & 11) == 0) g->fn = g3i::tile_l11_d16(g->in_smem);
    else
    g->fn = dual ? (wide ? g3i::tile_g16(g->in_smem) : g3i::tile_g5(g->in_smem))
                 : (wide ? g3i::tile_d16(g->in_smem) : threads > 768 ? g3i::tile_d5_w32(g->in_smem) : threads > 512 ? g3i::tile_d5_w24(g->in_smem) : g3i::tile_d5(g->in_smem));
    CUDA_TRY(cudaFuncSetAttribute(g->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g->per_sm, g->fn, (m->targs.W + 1) * 32, g->smem));
    if (g->per_sm < 1) return fail(G3CUDA_EUNSUPP, "tile kernel does not fit one SM (%d threads, %zu bytes of shared memory)", (m->targs.W + 1) * 32, g->smem);
    g->slots = (long long)g->per_sm * m->num_sms;
    if (m->max_tiles > 0) g->slots = std::min<long long>(g->slots, m->max_tiles);
    return G3CUDA_OK;
}

template <class T>
int launch_tile(g3cuda_model* m, const g3::KArgs& call, long long work, cudaStream_t st) {
    TileGeom g;
    int rc = tile_geom<T>(m, &g);
    if (rc) return rc;
    g3t::TArgs A = m->targs;
    A.par = call.par; A.B = call.B; A.tangent_idx = call.tangent_idx; A.K = call.K;
    A.nll = call.nll; A.comp_out = call.comp; A.grad = call.grad;
    A.meta_in_smem = g.meta_in_smem; A.use_h = g.use_h;
    // few evaluations: spread them over the SMs with partly filled tiles (latency); many: full 32-lane tiles
    const int lpt = (int)std::min<long long>(g3t::NL, std::max<long long>(1, (work + g.slots - 1) / g.slots));
    const long long ntile = (work + lpt - 1) / lpt;
    const long long grid = std::max<long long>(1, std::min(ntile, g.slots));
    constexpr size_t cpe = (size_t)(g3::Tan<T>::n + 1) * g3t::NL;
    A.lanes_per_tile = lpt;
    A.slab_doubles = (long long)((size_t)(A.n_g + (g.in_smem ? 0 : m->tplan.n_s)) * cpe);
    const size_t need = (size_t)A.slab_doubles * sizeof(double) * (size_t)grid;
    if (need > m->t_gws_bytes) {
        CUDA_TRY(cudaDeviceSynchronize());      // an earlier asynchronous launch may still use the slabs
        if (m->t_gws) cudaFree(m->t_gws);
        m->t_gws = nullptr; m->t_gws_bytes = 0;
        const size_t want = (size_t)A.slab_doubles * sizeof(double) * (size_t)g.slots;     // all tiles that can be resident
        if (cudaMalloc(&m->t_gws, want) != cudaSuccess) { cudaGetLastError(); return fail(G3CUDA_ENOMEM, "cudaMalloc of the %zu-byte tile slabs failed", want); }
        m->t_gws_bytes = want;
    }
    A.gws = m->t_gws;
    return launch_ptr(m, g.fn, (unsigned)grid, (unsigned)((A.W + 1) * 32), g.smem, st, &A);       // column warps + the keeper warp
}

// ============================================================ thread kernel
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
        A.w_suit_nset[q] = 1;
        if (Q[G3PP_SUIT_TMAP] >= 0) for (int t = 0; t < I[G3H_NT]; t++) A.w_suit_nset[q] = std::max(A.w_suit_nset[q], I[Q[G3PP_SUIT_TMAP] + t] + 1);
        A.t_suitq[q] = take(A.w_suit_nset[q] * PL * I[I[G3H_STOCK_OFF] + Q[G3PP_PREY] * G3S_LEN + G3S_L]);
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
        A.w_mort_nset[s] = 1;
        if (St[G3S_MORT_OFF] >= 0 && St[G3S_MORT_TMAP_OFF] >= 0)
            for (int t = 0; t < I[G3H_NT]; t++) A.w_mort_nset[s] = std::max(A.w_mort_nset[s], I[St[G3S_MORT_TMAP_OFF] + t] + 1);
        A.t_mort[s] = take(St[G3S_MORT_OFF] >= 0 ? A.w_mort_nset[s] * St[G3S_NAGE] * St[G3S_NAREA] * A.w_ndt : 0);
        A.t_rd[s] = take(St[G3S_N_RENEW] * L);
        A.t_rw[s] = take(St[G3S_N_RENEW] * L);
        A.w_grow_nset[s] = 1;
        if (n >= 0 && St[G3S_GROW_TMAP_OFF] >= 0)
            for (int t = 0; t < I[G3H_NT]; t++) A.w_grow_nset[s] = std::max(A.w_grow_nset[s], I[St[G3S_GROW_TMAP_OFF] + t] + 1);
        const int gn = A.w_grow_nset[s];
        A.t_pw[s] = take(n >= 0 ? gn * L : 0);
        if (n < 0) { A.t_g[s] = A.t_gr[s] = A.t_gn[s] = 0; continue; }
        int g = (n + 1) * L; maxg = std::max(maxg, g);
        A.t_g[s] = take(gn * g * A.w_ndt);
        A.w_ncset[s] = 1;
        if (St[G3S_MAT_KIND] == G3MAT_CONTINUOUS && St[G3S_N_OUT] > 0) {
            // an age term (beta, a50: R/action_mature.R:18,44-74) makes the maturing share differ from column to column
            if (I[St[G3S_MAT_OFF] + 2] >= 0) A.w_ncset[s] = St[G3S_NAGE] * St[G3S_NAREA];
            A.t_gr[s] = take(gn * A.w_ncset[s] * g * A.w_ndt); A.t_gn[s] = take(gn * A.w_ncset[s] * g * A.w_ndt);
        } else { A.t_gr[s] = A.t_gn[s] = A.t_g[s]; }
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
    A.t_sps = take(A.NS);
    A.t_cblk = 2;       // wide bands: columns per growth sweep block (C4 on B200: 1: 48.9 k, 2: 51.9 k, 3: 50.4 k, all columns: 46.3 k evals/s)
    return off;
}

int thread_workspace(g3cuda_model* m, size_t need) {
    if (need <= m->t_ws_bytes) return G3CUDA_OK;
    CUDA_TRY(cudaDeviceSynchronize());      // the workspace may still be in use by an earlier asynchronous launch
    if (m->t_ws) cudaFree(m->t_ws);
    m->t_ws = nullptr; m->t_ws_bytes = 0;
    if (cudaMalloc(&m->t_ws, need) != cudaSuccess) { cudaGetLastError(); return fail(G3CUDA_ENOMEM, "cudaMalloc of the %zu-byte evaluation workspace failed", need); }
    m->t_ws_bytes = need;
    return G3CUDA_OK;
}

template <class T>
int launch_thread(g3cuda_model* m, const g3::KArgs& call, long long work, cudaStream_t st) {
    constexpr bool is_double = std::is_same<T, double>::value;
    if (!m->thread_ok) return fail(G3CUDA_EUNSUPP, "model does not fit the thread-per-evaluation kernel: %s", m->thread_why.c_str());
    g3::KArgs A = m->base;
    A.par = call.par; A.B = call.B; A.tangent_idx = call.tangent_idx; A.K = call.K;
    A.nll = call.nll; A.comp = call.comp; A.grad = call.grad; A.report = call.report;
    const bool cm = A.any_const_mat != 0, wide = m->t_maxn > 4;
    int rc;
    if (A.report) {     // single-vector report run: one warp (lane 0 is the evaluation), workspace of 32 threads
        if (!is_double) return fail(G3CUDA_EINVAL, "report runs are plain-double");
        if ((rc = thread_workspace(m, (size_t)m->t_entries * sizeof(double) * g3i::REPORT_STRIDE))) return rc;
        A.t_ws = m->t_ws;
        return launch_ptr(m, g3i::thread_report(wide, cm), 1, 32, sizeof(double) * g3::MAXTS * 32, st, &A);
    }
    // plain doubles and more work than one 512-thread wave: the 1024-thread (64-register) variant
    const bool big = is_double && !wide && work > (long long)TPB * m->num_sms;
    const int CB = big ? TPB_BIG : TPB;
    const void* fn = is_double ? (wide ? g3i::thread_d16(cm) : g3i::thread_d5(big, cm)) : (wide ? g3i::thread_g16(cm) : g3i::thread_g5(cm));
    const size_t smem = sizeof(T) * g3::MAXTS * CB;
    CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // One block per SM, sized so that the batch splits into equally full waves: an evaluation is
    // one thread for its whole life, so a ragged last wave would idle most of the machine.
    const int sms = std::min<int>(m->num_sms, (int)(WS_STRIDE / CB));
    const long long slots = (long long)CB * sms;
    const long long waves = (work + slots - 1) / slots;
    const long long per_wave = (work + waves - 1) / waves;
    long long grid = std::max<long long>(1, std::min<long long>(sms, (per_wave + 31) / 32));
    int block = (int)(((per_wave + grid - 1) / grid + 31) / 32 * 32);
    block = std::max(32, std::min(block, CB));
    if ((rc = thread_workspace(m, (size_t)m->t_entries * sizeof(T) * (size_t)WS_STRIDE))) return rc;
    A.t_ws = m->t_ws;
    return launch_ptr(m, fn, (unsigned)grid, (unsigned)block, smem, st, &A);
}

// Which kernel evaluates a launch of `work` items of scalar type T: the tile kernel (state in shared memory, or in its
// L2-resident slab when it does not fit or carries tangents) whenever the model fits it -- values and gradients, few vectors
// (a CTA shares an evaluation: a single fn(par) is a ~ms call) or many. Measured on gradients (tools/gpu_grad_sweep.py,
// (value + gradient)/s, tile vs thread kernel): C1 36.9 k vs 38.6 k, C2 8.0 k vs 5.7 k, C4 444 vs 151, C5 2.1 k vs 1.4 k.
// The thread-per-evaluation kernel writes g3cuda_report()'s arrays and takes over models the tile planner refuses.
template <class T>
bool use_tile(const g3cuda_model* m, long long work) {
    (void)work;
    if (m->force_kernel == 4) return true;
    if (m->force_kernel == 3) return false;
    return m->tile_ok;
}

template <class T>
int launch(g3cuda_model* m, const g3::KArgs& call, long long work, cudaStream_t st) {
    if (call.report) return launch_thread<T>(m, call, work, st);
    if (use_tile<T>(m, work)) {
        if (!m->tile_ok) return fail(G3CUDA_EUNSUPP, "model does not fit the tile kernel: %s", m->tplan.why.c_str());
        return launch_tile<T>(m, call, work, st);
    }
    return launch_thread<T>(m, call, work, st);
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
    if (I[G3H_NSTOCK] < 1) return fail(G3CUDA_EINVAL, "no stocks");
    if (I[G3H_NSTEPS] > 31) return fail(G3CUDA_EUNSUPP, "more than 31 steps per year");
    return G3CUDA_OK;
}

// The thread kernel's tables (fixed-capacity arrays inside KArgs). Returns "" or why the model does not fit it.
std::string plan_thread(g3cuda_model* m, std::vector<int32_t>& rem_off, std::vector<int32_t>& rem_slots,
                        std::vector<int32_t>& tsid, std::vector<int4>& agedst, std::vector<int4>& colinc) {
    using namespace g3;
    const int32_t* I = m->I.data();
    KArgs& A = m->base;
    A.NC = I[G3H_NCELL]; A.NS = I[G3H_NSTOCK]; A.NPP = I[G3H_NPP]; A.NCOMP = I[G3H_NCOMP]; A.NX = I[G3H_NXBUF];
    A.NSLOT = I[G3H_NSLOT]; A.NPAR = I[G3H_NPAR]; A.NT = I[G3H_NT]; A.NSTEPS = I[G3H_NSTEPS]; A.NNLL = I[G3H_NNLL];
    A.NBOUND = I[G3H_NBOUND];
    if (A.NS > MAXS) return "more than " + std::to_string(MAXS) + " stocks";
    if (A.NCOMP > MAXC) return "more than " + std::to_string(MAXC) + " likelihood components";
    if (A.NX > MAXX) return "more than " + std::to_string(MAXX) + " collect vectors";
    if (A.NPP > 32) return "more than 32 predator-prey pairs";
    if (I[G3H_NPRED] > MAXS) return "more than " + std::to_string(MAXS) + " predators";
    auto stock = [&](int s) { return I + I[G3H_STOCK_OFF] + s * G3S_LEN; };

    // ---- per-cell tables
    int maxnarea = 1;
    std::vector<int4> cell(A.NC);
    std::vector<int32_t> inc_src(A.NC, -1), inc_slot(A.NC, -1), inc_mask(A.NC, 0), age_src(A.NC, -1), age_slot(A.NC, -1);
    rem_off.assign(A.NC, 0); rem_slots.clear();
    A.any_const_mat = 0;
    for (int s = 0; s < A.NS; s++) {
        const int32_t* St = stock(s);
        int L = St[G3S_L], nage = St[G3S_NAGE], narea = St[G3S_NAREA], c0 = St[G3S_CELL_OFF];
        maxnarea = std::max(maxnarea, narea);
        if (St[G3S_N_RENEW] > MAXRN) return "more than " + std::to_string(MAXRN) + " renewal actions on one stock";
        if (St[G3S_GROW_N] >= 0 && St[G3S_MAT_KIND] == G3MAT_CONSTANT && St[G3S_N_OUT] > 0) A.any_const_mat = 1;
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
                if (inc_src[d] >= 0) return "a cell receives maturing fish from more than one source";
                inc_src[d] = c0 + i; inc_slot[d] = slot; inc_mask[d] = St[G3S_TRANS_MASK];
                rem_slots.push_back(slot);
            }
            rem_slots.push_back(-1);
        }
        if (St[G3S_AGE_ENABLE]) for (int o = 0; o < St[G3S_AGE_N_OUT]; o++) for (int r = 0; r < narea; r++) for (int l = 0; l < L; l++) {
            int d = I[St[G3S_AGE_MAP_OFF] + (o * narea + r) * L + l];
            if (d < 0) continue;
            if (age_src[d] >= 0) return "a cell receives ageing fish from more than one source";
            age_src[d] = c0 + (r * nage + nage - 1) * L + l;
            age_slot[d] = I[St[G3S_AGE_OUT_OFF] + 2 * o + 1];
        }
    }
    A.maxnarea = maxnarea;
    // ---- predation tables
    tsid.assign((size_t)std::max(A.NPP, 1) * maxnarea, -1);
    int nts = 0;
    std::vector<int> pred_ts_base(I[G3H_NPRED], 0);
    for (int p = 0; p < I[G3H_NPRED]; p++) {
        const int32_t* P = I + I[G3H_PRED_OFF] + p * G3P_LEN;
        if (P[G3P_KIND] == G3PRED_STOCK) { pred_ts_base[p] = -1; continue; }
        if (P[G3P_KIND] != G3PRED_TOTALFLEET && P[G3P_KIND] != G3PRED_NUMBERFLEET && P[G3P_KIND] != G3PRED_LINEARFLEET && P[G3P_KIND] != G3PRED_EFFORTFLEET)
            return "predator kind " + std::to_string(P[G3P_KIND]) + " is not supported";
        pred_ts_base[p] = nts; nts += P[G3P_NAREA];
    }
    if (nts > MAXTS) return "more than " + std::to_string(MAXTS) + " (fleet, area) pairs";
    A.NTS = nts;
    std::vector<std::vector<int>> pp_of_stock(A.NS);
    for (int q = 0; q < A.NPP; q++) {
        const int32_t* Q = I + I[G3H_PP_OFF] + q * G3PP_LEN;
        const int32_t* P = I + I[G3H_PRED_OFF] + Q[G3PP_PRED] * G3P_LEN;
        const int32_t* St = stock(Q[G3PP_PREY]);
        const int sk = Q[G3PP_SUIT_KIND];
        if (sk == G3SUIT_ANDERSEN || sk == G3SUIT_EXPONENTIAL || sk == G3SUIT_RICHARDS) {
            if (P[G3P_KIND] != G3PRED_STOCK) return "suitability kind " + std::to_string(sk) + " reads predator_length: it needs a stock predator";
        } else if (sk < 0 || sk > G3SUIT_RICHARDS) return "suitability kind " + std::to_string(sk) + " is not supported";
        pp_of_stock[Q[G3PP_PREY]].push_back(q);
        for (int r = 0; r < St[G3S_NAREA]; r++) {
            int area = I[St[G3S_AREA_OFF] + r];
            for (int k = 0; k < P[G3P_NAREA]; k++) if (I[P[G3P_AREA_OFF] + k] == area)
                tsid[(size_t)q * maxnarea + r] = P[G3P_KIND] == G3PRED_STOCK ? k : pred_ts_base[Q[G3PP_PRED]] + k;
        }
    }
    int ppo = 0;
    for (int s = 0; s < A.NS; s++) {
        A.stock_pp_off[s] = ppo;
        if ((int)pp_of_stock[s].size() > MAXQ) return "more than " + std::to_string(MAXQ) + " predators on one stock";
        for (int q : pp_of_stock[s]) A.stock_pp[ppo++] = q;
        A.us_flag[s] = 0;
    }
    A.stock_pp_off[A.NS] = ppo;
    for (int i = 0; i < I[G3H_US_N]; i++) A.us_flag[I[I[G3H_US_OFF] + i]] = 1;

    // ---- distinct step lengths, compact hand-over tables
    A.w_ndt = 0;
    for (int st = 0; st < A.NSTEPS; st++) {
        int len = I[I[G3H_STEPLEN_OFF] + st], k = 0;
        while (k < A.w_ndt && A.w_dt_len[k] != len) k++;
        if (k == A.w_ndt) { if (A.w_ndt == 4) return "more than 4 distinct step lengths"; A.w_dt_len[A.w_ndt++] = len; }
        A.w_step_idt[st] = k;
    }
    std::vector<int> mv_off(A.NS, 0);
    {
        int nmv = 0;
        for (int s = 0; s < A.NS; s++) {
            const int32_t* St = stock(s);
            mv_off[s] = nmv;
            if (St[G3S_AGE_ENABLE] && St[G3S_AGE_N_OUT] > 0) nmv += 2 * St[G3S_NAREA] * St[G3S_L];
            A.w_inc_steps[s] = 0;
        }
    }
    for (int c = 0; c < A.NC; c++) if (inc_src[c] >= 0) A.w_inc_steps[cell[c].x] |= inc_mask[c];
    agedst.clear();
    for (int c = 0; c < A.NC; c++) if (age_src[c] >= 0) {
        int src = age_src[c], ss = cell[src].x;
        const int32_t* St = stock(ss);
        int L = St[G3S_L], i = cell[src].w * L + cell[src].y;       // (area idx, l) of the oldest column
        agedst.push_back(make_int4(c, mv_off[ss] + i, mv_off[ss] + St[G3S_NAREA] * L + i, age_slot[c]));
    }
    A.w_n_agedst = (int)agedst.size();
    for (int p = 0; p < I[G3H_NPRED]; p++) {        // accumulators of (fleet, area): where their landings live
        const int32_t* P = I + I[G3H_PRED_OFF] + p * G3P_LEN;
        if (P[G3P_KIND] == G3PRED_STOCK) continue;
        for (int k = 0; k < P[G3P_NAREA]; k++) {
            A.w_ts_eoff[pred_ts_base[p] + k] = P[G3P_E_ROFF] + k;
            A.w_ts_estr[pred_ts_base[p] + k] = P[G3P_NAREA];
            A.w_ts_kind[pred_ts_base[p] + k] = P[G3P_KIND];
        }
    }
    for (int x = 0; x < A.NX; x++) {
        const int32_t* X = I + I[G3H_XBUF_OFF] + x * G3X_LEN;
        A.w_x_kind[x] = X[G3X_KIND]; A.w_x_unit[x] = X[G3X_UNIT]; A.w_x_ppmask[x] = 0;
        for (int j = 0; j < X[G3X_N_PP]; j++) { int q = I[X[G3X_PP_OFF] + j]; if (q < 32) A.w_x_ppmask[x] |= 1u << q; }
    }

    // ---- workspace layout, ordering checks
    for (int s = 0; s < A.NS; s++) {
        const int32_t* St = stock(s);
        A.t_mig_steps[s] = 0;
        if (St[G3S_MIGRATE_OFF] < 0) continue;
        int NR = St[G3S_NAREA];
        if (NR > MAXMIG) return "a migrating stock on more than " + std::to_string(MAXMIG) + " areas";
        for (int st = 0; st < A.NSTEPS; st++) for (int i = 0; i < NR * NR; i++)
            if (I[St[G3S_MIGRATE_OFF] + st * NR * NR + i] >= 0) A.t_mig_steps[s] |= 1 << st;
    }
    m->t_maxn = 0;
    for (int s = 0; s < A.NS; s++) m->t_maxn = std::max(m->t_maxn, stock(s)[G3S_GROW_N]);
    if (m->t_maxn > 15) return "maxlengthgroupgrowth above 15";
    // the fused pass hands maturing fish over within one sweep: sources must come before destinations
    for (int c = 0; c < A.NC; c++) if (inc_src[c] >= 0 && cell[inc_src[c]].x >= cell[c].x) return "a maturation source stock is ordered after its destination";
    m->t_entries = make_tlayout(m, A);
    // hand-over per COLUMN: every cell of a destination column must be fed by the same length of one source column
    colinc.clear();
    std::vector<int32_t> tr_of_cell(A.NC, -1);
    {
        int ntr = 0;
        for (int s = 0; s < A.NS; s++) {
            const int32_t* St = stock(s);
            int ncell = St[G3S_L] * St[G3S_NAGE] * St[G3S_NAREA];
            if (St[G3S_GROW_N] >= 0 && St[G3S_N_OUT] > 0) { for (int i = 0; i < ncell; i++) tr_of_cell[St[G3S_CELL_OFF] + i] = ntr + i; ntr += ncell; }
        }
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
                if (!ok) return "maturation hand-over is not column to column";
                if (rem_off[c] != rem_off[c00] && l > 0) {       // remainder ratios must be the same down a column
                    int a = rem_off[c], b = rem_off[c00];
                    while (rem_slots[a] >= 0 && rem_slots[a] == rem_slots[b]) { a++; b++; }
                    if (rem_slots[a] != rem_slots[b]) return "maturation outputs differ between the lengths of one column";
                }
            }
            colinc.push_back(e);
        }
    }
    // one renewal per (column, step) is what the fused pass applies
    for (int s = 0; s < A.NS; s++) {
        const int32_t* St = stock(s);
        for (int k = 0; k < St[G3S_N_RENEW]; k++) for (int k2 = k + 1; k2 < St[G3S_N_RENEW]; k2++) {
            const int32_t* Ra = I + St[G3S_RENEW_OFF] + k * G3R_LEN; const int32_t* Rb = I + St[G3S_RENEW_OFF] + k2 * G3R_LEN;
            if (Ra[G3R_AGE_IDX] == Rb[G3R_AGE_IDX] && (Ra[G3R_RUN_STEP] == 0 || Rb[G3R_RUN_STEP] == 0 || Ra[G3R_RUN_STEP] == Rb[G3R_RUN_STEP])) {
                // ... unless they never run in the same year (a renewal whose shape varies by year, one record per set of years)
                bool together = false;
                const int nya = (I[G3H_END_YEAR] - I[G3H_START_YEAR] + 1) * St[G3S_NAREA];
                for (int i = 0; i < nya; i++) together = together || (I[Ra[G3R_FACTOR_OFF] + i] >= 0 && I[Rb[G3R_FACTOR_OFF] + i] >= 0);
                if (together) return "two renewals into the same age at the same step";
            }
        }
    }
    for (int c = 0; c < A.NCOMP; c++) {
        const int32_t* C = I + I[G3H_COMP_OFF] + c * G3C_LEN;
        if (C[G3C_FUNC] < 0 || C[G3C_FUNC] > G3F_SUMSQLOGRATIOS) return "likelihood function " + std::to_string(C[G3C_FUNC]) + " is not supported";
    }
    // 32-bit element indexing of the workspace (in units of T): entries x stride must stay below 2^32
    if ((double)m->t_entries * (double)WS_STRIDE >= 4294967295.0) return "evaluation workspace too large for 32-bit element indexing";
    return "";
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
    else if (which == 2) out[i] = g3::lsa(x[i] * -1e3, y[i] * -1e3) / -1e3;
    else if (which == 3) out[i] = g3::avoid_zero_x(x[i]);
    else out[i] = g3::dif_pmax_x(x[i], y[i], -1e3);
}

extern "C" {

int g3cuda_math_probe(int device, int which, const double* x, const double* y, int64_t n, double* out) {
    if (!x || !out || n < 0 || which < 0 || which > 4 || (which != 1 && which != 3 && !y)) return fail(G3CUDA_EINVAL, "bad argument");
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
int32_t g3cuda_n_par(const g3cuda_model* m) { return m ? m->NPAR : 0; }
int32_t g3cuda_n_nll(const g3cuda_model* m) { return m ? m->NNLL : 0; }
int32_t g3cuda_n_steps(const g3cuda_model* m) { return m ? m->NT : 0; }
int64_t g3cuda_report_len(const g3cuda_model* m) { return m ? m->report_len : 0; }
int64_t g3cuda_launch_count(const g3cuda_model* m) { return m ? m->launches : 0; }

void g3cuda_model_free(g3cuda_model* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    cudaDeviceSynchronize();
    for (void* p : m->owned) cudaFree(p);
    for (void* p : {(void*)m->d_par, (void*)m->d_comp, (void*)m->d_grad, (void*)m->d_tidx, (void*)m->d_out, (void*)m->d_nll2,
                    m->t_ws, (void*)m->t_gws, (void*)m->d_tb})
        if (p) cudaFree(p);
    if (m->h_stage) cudaFreeHost(m->h_stage);
    for (cudaEvent_t e : {m->ev0, m->ev1, m->ws_free, m->ev_copy[0], m->ev_copy[1]}) if (e) cudaEventDestroy(e);
    if (m->stream) cudaStreamDestroy(m->stream);
    if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
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
    m->NPAR = I[G3H_NPAR]; m->NNLL = I[G3H_NNLL]; m->NT = I[G3H_NT]; m->NCOMP = I[G3H_NCOMP];
    auto bail = [&](int code) { g3cuda_model_free(m); return code; };
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(fail(G3CUDA_ECUDA, "cudaGetDeviceProperties failed"));
    m->num_sms = prop.multiProcessorCount;
    m->smem_optin = prop.sharedMemPerBlockOptin;
    auto up = [&](auto& vec) { auto* d = dev_copy(vec); if (d) m->owned.push_back((void*)d); return d; };
    const int32_t* d_I = up(m->I); const double* d_R = up(m->R);
    if (!d_I || !d_R) return bail(fail(G3CUDA_ENOMEM, "device allocation failed"));

    // ---- tile kernel: the planner's tables
    m->tplan = g3t::make_tile_plan(I, R, 0, 0, 15, m->smem_optin);
    if (m->tplan.ok) {
        g3t::TilePlan& TP = m->tplan;
        auto* d_op = up(TP.obsprop); auto* d_oc = up(TP.obsconst); auto* d_pa = up(TP.pred_active);
        m->d_pred_active = d_pa;
        if (!d_op || !d_oc || !d_pa)
            return bail(fail(G3CUDA_ENOMEM, "device allocation failed"));
        TP.a.I = d_I; TP.a.R = d_R; TP.a.obsprop = d_op; TP.a.obsconst = d_oc;
        if ((rc = tile_upload_tb(m))) return bail(rc);
        m->tile_ok = true;
    }

    // ---- thread kernel: fixed-capacity tables inside KArgs
    {
        std::vector<int32_t> rem_off, rem_slots, tsid;
        std::vector<int4> agedst, colinc;
        m->thread_why = plan_thread(m, rem_off, rem_slots, tsid, agedst, colinc);
        if (m->thread_why.empty() && !m->tplan.ok) m->thread_why = m->tplan.why;       // (it shares the planner's observation tables)
        m->thread_ok = m->thread_why.empty();
        if (m->thread_ok) {
            KArgs& A = m->base;
            A.I = d_I; A.R = d_R;
            A.rem_off = up(rem_off); A.rem_slots = up(rem_slots); A.pp_tsid = up(tsid);
            A.w_agedst = up(agedst); A.t_colinc = up(colinc);
            if (!A.rem_off || !A.rem_slots || !A.pp_tsid || !A.w_agedst || !A.t_colinc) return bail(fail(G3CUDA_ENOMEM, "device allocation failed"));
            A.obsprop = m->tplan.a.obsprop; A.obsconst = m->tplan.a.obsconst; A.w_pred_active = m->d_pred_active;
            for (int c = 0; c < A.NCOMP; c++) { A.obs_off[c] = m->tplan.comp[c].obs_off; A.obsconst_off[c] = m->tplan.comp[c].obsconst_off; }
        }
    }
    if (!m->tile_ok && !m->thread_ok)
        return bail(fail(G3CUDA_EUNSUPP, "model outside the supported subset: %s", m->tplan.ok ? m->thread_why.c_str() : m->tplan.why.c_str()));

    // ---- report layout (include/g3cuda.h)
    {
        auto stock = [&](int s) { return I + I[G3H_STOCK_OFF] + s * G3S_LEN; };
        const int NS = I[G3H_NSTOCK], NPP = I[G3H_NPP], NT = m->NT;
        KArgs& A = m->base;
        int64_t pos = 1 + (int64_t)m->NNLL * NT;
        for (int s = 0; s < NS; s++) {
            const int32_t* St = stock(s);
            if (s < MAXS) A.rep_dstart[s] = pos;
            pos += 2LL * NT * St[G3S_L] * St[G3S_NAGE] * St[G3S_NAREA];
        }
        for (int c = 0; c < m->NCOMP; c++) {
            const int32_t* C = I + I[G3H_COMP_OFF] + c * G3C_LEN;
            if (c < MAXC) A.rep_model[c] = pos;
            pos += (int64_t)C[G3C_N_TIMES] * C[G3C_N_DEST];
        }
        A.rep_params = pos; m->rep_params = pos; pos += 2LL * m->NCOMP;
        for (int s = 0; s < NS; s++) {
            const int32_t* St = stock(s);
            if (s < MAXS) A.rep_renew[s] = pos;
            pos += (int64_t)NT * St[G3S_L] * St[G3S_NAGE] * St[G3S_NAREA];
        }
        for (int q = 0; q < NPP; q++) {
            const int32_t* Q = I + I[G3H_PP_OFF] + q * G3PP_LEN;
            const int32_t* P = I + I[G3H_PRED_OFF] + Q[G3PP_PRED] * G3P_LEN;
            const int32_t* St = stock(Q[G3PP_PREY]);
            const int64_t PL = P[G3P_KIND] == G3PRED_STOCK ? stock(P[G3P_STOCK])[G3S_L] : 1;
            if (q < 32) A.rep_pp[q] = pos;
            pos += PL * St[G3S_L] + 2LL * NT * St[G3S_L] * St[G3S_NAGE] * St[G3S_NAREA];
        }
        m->report_len = pos;
    }
    if (cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&m->ev0) != cudaSuccess || cudaEventCreate(&m->ev1) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->ev_copy[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->ev_copy[1], cudaEventDisableTiming) != cudaSuccess)
        return bail(fail(G3CUDA_ECUDA, "stream/event creation failed"));
    *out = m;
    return G3CUDA_OK;
}

// launches on `st`; no timing events
static int eval_device(g3cuda_model* m, const double* d_par, int64_t B, const int32_t* d_tangent_idx, int32_t K,
                       double* d_nll, double* d_comp, double* d_grad, cudaStream_t st) {
    g3::KArgs call{};
    call.par = d_par; call.B = B; call.nll = d_nll; call.comp = d_comp; call.report = nullptr;
    call.tangent_idx = nullptr; call.K = 0; call.grad = nullptr;
    if (d_grad && K > 0) {
        if (!d_tangent_idx) return fail(G3CUDA_EINVAL, "grad requested without tangent_idx");
        // Forward-mode tangents: the work items are (vector, tile of NTAN directions). The first direction tile of a
        // vector also writes its nll and component sums, so no separate value launch is needed.
        const long long ntiles = (K + NTAN - 1) / NTAN;
        call.tangent_idx = d_tangent_idx; call.K = K; call.grad = d_grad;
        return launch<DualT>(m, call, B * ntiles, st);
    }
    return launch<double>(m, call, B, st);
}

int g3cuda_eval_batch_device(g3cuda_model* m, const double* d_par, int64_t B, const int32_t* d_tangent_idx, int32_t K,
                             double* d_nll, double* d_comp, double* d_grad, void* stream) {
    if (!m || !d_par || !d_nll || B < 0) return fail(G3CUDA_EINVAL, "bad argument");
    if (B == 0) return G3CUDA_OK;
    CUDA_TRY(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    m->timed = true;
    CUDA_TRY(cudaEventRecord(m->ev0, st));
    int rc = eval_device(m, d_par, B, d_tangent_idx, K, d_nll, d_comp, d_grad, st);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(m->ev1, st));
    return G3CUDA_OK;
}

int g3cuda_eval_batch(g3cuda_model* m, const double* par, int64_t B, const int32_t* tangent_idx, int32_t K,
                      double* nll, double* nll_comp, double* grad) {
    if (!m || !par || !nll || B < 0) return fail(G3CUDA_EINVAL, "bad argument");
    if (B == 0) return G3CUDA_OK;
    CUDA_TRY(cudaSetDevice(m->device));
    const int P = m->NPAR, NN = m->NNLL;
    const bool want_grad = grad && K > 0;
    if (want_grad && !tangent_idx) return fail(G3CUDA_EINVAL, "grad requested without tangent_idx");
    // Projections: the reference runs project_years more years (R/action_time.R:21-106); the kernels hold tables for the years
    // the model was lowered for. Beyond them, say so instead of returning a NaN that looks like a value (NaN stays a VALUE for
    // the reference's own NaN case, retro_years < 0).
    if (const int pj = m->I[G3H_PAR_PROJECT]; pj >= 0) {
        // the per-step tables reach max_project_years past end_year (g3_to_cuda(max_project_years = ...))
        const int room = m->I[G3H_NT] / m->I[G3H_NSTEPS] - (m->I[G3H_END_YEAR] - m->I[G3H_START_YEAR] + 1);
        for (int64_t b = 0; b < B; b++)
            if (par[b * P + pj] >= (double)(room + 1))
                return fail(G3CUDA_EUNSUPP, "vector %lld sets project_years = %g: the model was lowered with room for %d projection year(s) (max_project_years)", (long long)b, par[b * P + pj], room);
    }
    int rc;
    if ((rc = ensure(&m->d_par, &m->cap_par, (size_t)B * P))) return rc;
    if ((rc = ensure(&m->d_out, &m->cap_out, (size_t)B))) return rc;
    if (nll_comp && (rc = ensure(&m->d_comp, &m->cap_comp, (size_t)B * NN))) return rc;
    if (want_grad) {
        if ((rc = ensure(&m->d_grad, &m->cap_grad, (size_t)B * K)) || (rc = ensure(&m->d_tidx, &m->cap_tidx, (size_t)K))) return rc;
        CUDA_TRY(cudaMemcpyAsync(m->d_tidx, tangent_idx, sizeof(int32_t) * K, cudaMemcpyHostToDevice, m->stream));
    }
    // results come back through a pinned staging buffer, so that the device -> host copies are asynchronous
    const size_t per = 1 + (nll_comp ? NN : 0) + (want_grad ? K : 0);
    if ((size_t)B * per > m->cap_stage) {
        if (m->h_stage) { cudaStreamSynchronize(m->stream); cudaFreeHost(m->h_stage); m->h_stage = nullptr; m->cap_stage = 0; }
        if (cudaHostAlloc(&m->h_stage, (size_t)B * per * sizeof(double), cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return fail(G3CUDA_ENOMEM, "pinned staging allocation failed"); }
        m->cap_stage = (size_t)B * per;
    }
    double* h_nll = m->h_stage; double* h_comp = h_nll + B; double* h_grad = h_comp + (nll_comp ? (size_t)B * NN : 0);
    // Two chunks: a first small one (one wave of tiles), so that the evaluation starts as soon as its parameters
    // have landed, and the rest, whose host -> device copy runs on the copy stream underneath the first launch.
    int64_t first = B;
    if (!want_grad && m->tile_ok && use_tile<double>(m, B)) {
        TileGeom g;
        if (tile_geom<double>(m, &g) == G3CUDA_OK) {
            const int64_t wave = g.slots * g3t::NL;
            if (B >= 4 * wave) first = wave;
        }
    }
    const int64_t cb[3] = {0, first, B};
    for (int c = 0; c < 2; c++) {
        const int64_t lo = cb[c], n = cb[c + 1] - cb[c];
        if (n <= 0) continue;
        CUDA_TRY(cudaMemcpyAsync(m->d_par + lo * P, par + lo * P, sizeof(double) * n * P, cudaMemcpyHostToDevice, m->copy_stream));
        CUDA_TRY(cudaEventRecord(m->ev_copy[c], m->copy_stream));
    }
    for (int c = 0; c < 2; c++) {
        const int64_t lo = cb[c], n = cb[c + 1] - cb[c];
        if (n <= 0) continue;
        CUDA_TRY(cudaStreamWaitEvent(m->stream, m->ev_copy[c], 0));
        rc = eval_device(m, m->d_par + lo * P, n, want_grad ? m->d_tidx : nullptr, want_grad ? K : 0, m->d_out + lo,
                         nll_comp ? m->d_comp + lo * NN : nullptr, want_grad ? m->d_grad + lo * K : nullptr, m->stream);
        if (rc) { cudaStreamSynchronize(m->copy_stream); return rc; }
        CUDA_TRY(cudaMemcpyAsync(h_nll + lo, m->d_out + lo, sizeof(double) * n, cudaMemcpyDeviceToHost, m->stream));
        if (nll_comp) CUDA_TRY(cudaMemcpyAsync(h_comp + lo * NN, m->d_comp + lo * NN, sizeof(double) * n * NN, cudaMemcpyDeviceToHost, m->stream));
        if (want_grad) CUDA_TRY(cudaMemcpyAsync(h_grad + lo * K, m->d_grad + lo * K, sizeof(double) * n * K, cudaMemcpyDeviceToHost, m->stream));
    }
    cudaError_t e = cudaStreamSynchronize(m->stream);
    if (e != cudaSuccess) return fail(G3CUDA_ECUDA, "evaluation failed: %s", cudaGetErrorString(e));
    std::memcpy(nll, h_nll, sizeof(double) * B);
    if (nll_comp) std::memcpy(nll_comp, h_comp, sizeof(double) * B * NN);
    if (want_grad) std::memcpy(grad, h_grad, sizeof(double) * B * K);
    return G3CUDA_OK;
}

// ---- one handle over several devices ("copies all spec/obs data to every selected device"): the rows
// of a batch are cut into contiguous blocks, one per device, evaluated concurrently by one host thread per device; no
// device talks to another (the path has no exchange step), every block lands in the caller's buffers directly.
struct g3cuda_multi {
    std::vector<g3cuda_model*> m;
};

int g3cuda_multi_create(const g3cuda_spec* spec, const int32_t* devices, int32_t n_devices, g3cuda_multi** out) {
    if (!out || !devices || n_devices < 1) return fail(G3CUDA_EINVAL, "bad argument");
    *out = nullptr;
    auto* mm = new g3cuda_multi;
    for (int i = 0; i < n_devices; i++) {
        g3cuda_model* h = nullptr;
        const int rc = g3cuda_model_create(spec, devices[i], &h);
        if (rc != G3CUDA_OK) {          // (the error text is the failing device's)
            for (auto* x : mm->m) g3cuda_model_free(x);
            delete mm;
            return rc;
        }
        mm->m.push_back(h);
    }
    *out = mm;
    return G3CUDA_OK;
}

void g3cuda_multi_free(g3cuda_multi* mm) {
    if (!mm) return;
    for (auto* x : mm->m) g3cuda_model_free(x);
    delete mm;
}

int32_t g3cuda_multi_n_devices(const g3cuda_multi* mm) { return mm ? (int32_t)mm->m.size() : 0; }
g3cuda_model* g3cuda_multi_model(g3cuda_multi* mm, int32_t i) { return (mm && i >= 0 && i < (int32_t)mm->m.size()) ? mm->m[i] : nullptr; }

int g3cuda_multi_eval_batch(g3cuda_multi* mm, const double* par, int64_t B, const int32_t* tangent_idx, int32_t K,
                            double* nll, double* nll_comp, double* grad) {
    if (!mm || mm->m.empty() || !par || !nll || B < 0) return fail(G3CUDA_EINVAL, "bad argument");
    const int n = (int)mm->m.size();
    const int64_t P = mm->m[0]->NPAR, NN = mm->m[0]->NNLL;
    std::vector<int> rcs(n, G3CUDA_OK);
    std::vector<std::string> errs(n);
    std::vector<std::thread> th;
    const int64_t base = B / n, extra = B % n;      // the first B % n devices take one row more (gadget3_b200/shard.py)
    int64_t lo = 0;
    for (int i = 0; i < n; i++) {
        const int64_t rows = base + (i < extra ? 1 : 0), at = lo;
        lo += rows;
        if (rows == 0) continue;
        th.emplace_back([&, i, rows, at]() {
            rcs[i] = g3cuda_eval_batch(mm->m[i], par + at * P, rows, tangent_idx, K, nll + at,
                                       nll_comp ? nll_comp + at * NN : nullptr, (grad && K > 0) ? grad + at * (int64_t)K : nullptr);
            if (rcs[i] != G3CUDA_OK) errs[i] = g_err;      // (the error text is thread-local)
        });
    }
    for (auto& t : th) t.join();
    for (int i = 0; i < n; i++)
        if (rcs[i] != G3CUDA_OK) return fail(rcs[i], "device %d: %s", mm->m[i]->device, errs[i].c_str());
    return G3CUDA_OK;
}

int g3cuda_set_kernel(g3cuda_model* m, int which) {
    if (!m || (which != 0 && which != 3 && which != 4 && which != 5)) return fail(G3CUDA_EINVAL, "bad argument: kernel must be 0 (automatic), 3 (thread), 4 (tile) or 5 (tile, full instantiation)");
    m->no_lean = which == 5;
    if (which == 5) which = 4;
    if (which == 3 && !m->thread_ok) return fail(G3CUDA_EUNSUPP, "model does not fit the thread-per-evaluation kernel: %s", m->thread_why.c_str());
    if (which == 4 && !m->tile_ok) return fail(G3CUDA_EUNSUPP, "model does not fit the tile kernel: %s", m->tplan.why.c_str());
    m->force_kernel = which;
    return G3CUDA_OK;
}

int g3cuda_set_tile_warps(g3cuda_model* m, int warps) {
    if (!m || warps < 0 || warps > 15) return fail(G3CUDA_EINVAL, "bad argument");
    if (!m->tile_ok) return fail(G3CUDA_EUNSUPP, "model does not fit the tile kernel: %s", m->tplan.why.c_str());
    CUDA_TRY(cudaSetDevice(m->device));
    int w = warps;
    if (w == 0) w = 15;
    g3t::plan_set_warps(m->tplan, w);        // (the partition into units stays: only their assignment to warps changes)
    g3t::plan_place_hot(m->tplan, m->smem_optin);
    return tile_upload_tb(m);
}

int g3cuda_set_tile_slots(g3cuda_model* m, int tiles) {
    if (!m || tiles < 0) return fail(G3CUDA_EINVAL, "bad argument");
    m->max_tiles = tiles;
    return G3CUDA_OK;
}

int g3cuda_set_tile_partition(g3cuda_model* m, int warps, int seg_rows) {
    if (!m || warps < 0 || warps > g3t::SHW - 1 || seg_rows < -1) return fail(G3CUDA_EINVAL, "bad argument");
    if (!m->tile_ok) return fail(G3CUDA_EUNSUPP, "model does not fit the tile kernel: %s", m->tplan.why.c_str());
    CUDA_TRY(cudaSetDevice(m->device));
    CUDA_TRY(cudaDeviceSynchronize());
    g3t::TilePlan np = g3t::make_tile_plan(m->I.data(), m->R.data(), warps, seg_rows, std::max(warps, 15), m->smem_optin);
    if (!np.ok) return fail(G3CUDA_EUNSUPP, "model does not fit the tile kernel: %s", np.why.c_str());
    const g3t::TArgs old = m->tplan.a;
    np.a.I = old.I; np.a.R = old.R; np.a.obsprop = old.obsprop; np.a.obsconst = old.obsconst;
    m->tplan = std::move(np);
    return tile_upload_tb(m);
}

int32_t g3cuda_tile_info(const g3cuda_model* m, int32_t* out /*[6]*/) {
    if (!m || !out) return fail(G3CUDA_EINVAL, "bad argument");
    out[0] = m->tile_ok; out[1] = m->tplan.a.W; out[2] = m->tplan.a.n_g; out[3] = m->tplan.n_s;
    out[4] = (size_t)m->tplan.n_s * g3t::NL * sizeof(double) <= m->smem_optin; out[5] = m->thread_ok;
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
    if (!m->thread_ok) return fail(G3CUDA_EUNSUPP, "g3cuda_report runs on the thread-per-evaluation kernel, which does not fit this model: %s", m->thread_why.c_str());
    int rc;
    double* d_rep = nullptr;
    if ((rc = ensure(&m->d_par, &m->cap_par, (size_t)m->NPAR))) return rc;
    if ((rc = ensure(&m->d_nll2, &m->cap_nll2, 1))) return rc;
    if (cudaMalloc(&d_rep, sizeof(double) * m->report_len) != cudaSuccess) { cudaGetLastError(); return fail(G3CUDA_ENOMEM, "cudaMalloc of the report failed"); }
    cudaMemsetAsync(d_rep, 0, sizeof(double) * m->report_len, m->stream);
    cudaMemcpyAsync(m->d_par, par, sizeof(double) * m->NPAR, cudaMemcpyHostToDevice, m->stream);
    g3::KArgs call{};
    call.par = m->d_par; call.B = 1; call.nll = m->d_nll2; call.report = d_rep;
    rc = launch<double>(m, call, 1, m->stream);
    if (!rc) {
        cudaMemcpyAsync(out, d_rep, sizeof(double) * m->report_len, cudaMemcpyDeviceToHost, m->stream);
        cudaError_t e = cudaStreamSynchronize(m->stream);
        if (e != cudaSuccess) rc = fail(G3CUDA_ECUDA, "report failed: %s", cudaGetErrorString(e));
        const int32_t* I = m->I.data();
        for (int c = 0; c < m->NCOMP && !rc; c++) {     // regression parameters exist for survey indices only
            int f = I[I[G3H_COMP_OFF] + c * G3C_LEN + G3C_FUNC];
            if (f != G3F_SI_LOG && f != G3F_SI_LINEAR) out[m->rep_params + 2 * c] = out[m->rep_params + 2 * c + 1] = std::nan("");
        }
    } else cudaStreamSynchronize(m->stream);
    cudaFree(d_rep);
    return rc;
}

}  // extern "C"
