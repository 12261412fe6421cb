// g3emu.cpp -- kernel-logic test harness: runs the BODY of the tile kernel (gadget3_b200/csrc/g3_tile.cuh) on the
// CPU, one OS thread per warp with a real barrier, on the tables the product's planner builds.
// TEST INFRASTRUCTURE ONLY: built by tests/test_tile_emu.py into tests/emu/_build/, never loaded by the product
// library (which has no CPU path), never timed by bench.py. It exists so that barrier placement, table layouts
// and index arithmetic of the CUDA kernel are checked against the oracle on machines without a GPU.
#include <barrier>
#include <cstring>
#include <thread>
#include <type_traits>
#include <vector>

#include "../../gadget3_b200/csrc/g3_tile_plan.h"

namespace {
using namespace g3t;
bool g_no_lean = false;     // g3emu_set_no_lean: the full instantiation even where a leaner one covers the model
struct Bar { std::barrier<>* b; };
void bar_wait(void* p) { static_cast<Bar*>(p)->b->arrive_and_wait(); }

template <class T, int NW, bool SMEM, bool XF, int LEAN = 0>
void run_xf(const TArgs& A, int W, int ncta, size_t smem_doubles) {
    for (int cta = 0; cta < ncta; cta++)
        for (int lane = 0; lane < A.lanes_per_tile; lane++) {
            std::vector<double> smem(smem_doubles + 64, 0.0);
            int ctr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            std::barrier<> sync(W + 1);        // W column warps + the keeper warp
            Bar bar{&sync};
            EmuCtx ctx{bar_wait, &bar};
            std::vector<std::thread> th;
            for (int w = 0; w <= W; w++)
                th.emplace_back([&, w]() { tile_body<T, NW, SMEM, 0, XF, LEAN>(A, smem.data(), ctr, w, lane, cta, ncta, ctx); });
            for (auto& t : th) t.join();
        }
}
// the instantiation the product's dispatch picks (TArgs::xf)
template <class T, int NW, bool SMEM>
void run(const TArgs& A, int W, int ncta, size_t smem_doubles) {
    if (A.xf) { run_xf<T, NW, SMEM, true>(A, W, ncta, smem_doubles); return; }
    // plain values: the leanest instantiation the product compiles that covers the model (g3cuda.cu: tile_geom)
    if constexpr (std::is_same<T, double>::value) {
        if (!g_no_lean) {
            if ((A.need & 15) == 0) { run_xf<T, NW, SMEM, false, 15>(A, W, ncta, smem_doubles); return; }
            if (NW == 5 && (A.need & 12) == 0) { run_xf<T, NW, SMEM, false, 12>(A, W, ncta, smem_doubles); return; }
            if (NW == 16 && (A.need & 11) == 0) { run_xf<T, NW, SMEM, false, 11>(A, W, ncta, smem_doubles); return; }
        }
    }
    run_xf<T, NW, SMEM, false>(A, W, ncta, smem_doubles);
}
}  // namespace

extern "C" void g3emu_set_no_lean(int32_t v) { g_no_lean = v != 0; }

extern "C" int g3emu_eval(const int32_t* I, const double* R, const double* par, int64_t B, const int32_t* tidx, int32_t K,
                          double* nll, double* comp, double* grad, int32_t W, int32_t seg, int32_t lanes_per_tile, int32_t smem_state,
                          int32_t ncta, char* why, int32_t why_len, int32_t meta_smem, int32_t use_h) {
    TilePlan P = make_tile_plan(I, R, W, seg);
    if (!P.ok) { if (why) { std::strncpy(why, P.why.c_str(), why_len - 1); why[why_len - 1] = 0; } return -4; }
    TArgs A = P.a;
    A.I = I; A.R = R; A.meta = P.meta.data(); A.meta_in_smem = meta_smem;
    A.obsprop = P.obsprop.data(); A.obsconst = P.obsconst.data();
    A.par = par; A.B = B; A.tangent_idx = tidx; A.K = K; A.nll = nll; A.comp_out = comp; A.grad = grad;
    A.lanes_per_tile = lanes_per_tile;
    const bool dual = grad && K > 0;
    const size_t cpe = (dual ? 5 : 1) * NL;          // doubles per entry
    A.slab_doubles = (long long)((size_t)(A.n_g + (smem_state ? 0 : P.n_s)) * cpe);
    std::vector<double> gws((size_t)A.slab_doubles * ncta + 64, 0.0);
    A.gws = gws.data();
    A.use_h = !dual && use_h;
    const size_t sd = (size_t)P.n_s * cpe + (size_t)A.n_h * cpe + (size_t)(A.meta_n + 1) / 2 + 8;       // state, hot tables, the plan's tables
    const bool wide = P.maxn > 4;
    if (!dual) {
        if (wide) { if (smem_state) run<double, 16, true>(A, A.W, ncta, sd); else run<double, 16, false>(A, A.W, ncta, sd); }
        else { if (smem_state) run<double, 5, true>(A, A.W, ncta, sd); else run<double, 5, false>(A, A.W, ncta, sd); }
    } else {
        if (wide) { if (smem_state) run<Dual<4>, 16, true>(A, A.W, ncta, sd); else run<Dual<4>, 16, false>(A, A.W, ncta, sd); }
        else { if (smem_state) run<Dual<4>, 5, true>(A, A.W, ncta, sd); else run<Dual<4>, 5, false>(A, A.W, ncta, sd); }
    }
    return 0;
}

extern "C" int g3emu_plan_info(const int32_t* I, const double* R, int32_t W, int32_t seg, int32_t* out /*[8 + stocks]*/) {
    TilePlan P = make_tile_plan(I, R, W, seg);
    if (!P.ok) return -4;
    out[0] = P.a.W; out[1] = P.a.n_g; out[2] = P.n_s; out[3] = P.a.NCOLS; out[4] = P.maxn; out[5] = (int)P.tb.size();
    out[6] = P.a.NU; out[7] = P.a.NTS;
    for (int s = 0; s < P.a.NS; s++) out[8 + s] = P.seglen[s];
    return 0;
}
