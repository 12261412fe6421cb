// g3_tile_plan.h -- host-side planner of the tile kernel (g3_tile.cuh): turns the packed model spec
// (include/g3cuda_spec.h) into the kernel's record tables, entry layouts (shared-memory state, per-CTA global
// slab) and the column -> warp assignment. Pure C++ (no CUDA API): g3cuda.cu uploads the result, tests/emu
// runs the kernel body on it on the CPU.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "g3_tile.cuh"

namespace g3t {

struct TilePlan {
    bool ok = false;
    std::string why;            // why the model does not fit the tile kernel
    TArgs a{};                  // scalars and entry offsets; pointers are filled by whoever places the tables
    std::vector<StockRec> st; std::vector<ColRec> col; std::vector<UnitRec> unit; std::vector<PairRec> pp; std::vector<PredRec> pred;
    std::vector<TsRec> ts; std::vector<CompRec> comp; std::vector<XRec> xb; std::vector<int4> agedst;
    std::vector<int32_t> tb;
    std::vector<int32_t> meta;  // all of the above in one block (plan_pack_meta): what the kernel reads, from shared memory if it fits
    std::vector<double> obsprop, obsconst;
    std::vector<unsigned char> pred_active;
    int n_s = 0;                // state entries per lane (shared memory, or after the G entries in the slab)
    int maxn = 0;               // widest growth band (maxlengthgroupgrowth)
    std::vector<int> seglen;    // rows per unit of each stock's columns (L: one unit per column)
    std::vector<double> row_cost;       // relative cost of one row of each column per step
    std::vector<double> unit_cost;
    int g_fixed = 0, tb_fixed = 0;      // G entries / tb words that do not depend on the partition into units
    int tb_units = 0;                   // tb words up to and including the partition's lists (the collect task lists follow)
    int max_cx = 0, maxnarea = 1;
    int ctas_per_sm = 1;        // CTAs the plan wants resident per SM (small models: 4): the shared-memory budget of one is 1 / that
    int max_W = 15;             // most unit warps (15 + keeper = 512 threads at 128 registers; 23 + keeper = 768 at 80)
    std::vector<std::vector<int>> pp_of_stock;
    // collect terms of all components (cell index, coefficient) and the tasks they are cut into: (dest, first term, end term, slot, group)
    std::vector<int32_t> cterm_idx; std::vector<double> cterm_coef;
    std::vector<std::vector<int>> ctask;
};

constexpr int TASK_TERMS = 24;          // longest run of collect terms one warp sums in one go

constexpr double UNIT_OVERHEAD = 1.5;   // per-unit cost (pointer setup, ghost rows) in rows of a plain column

inline double plan_avoid_zero(double a) {       // R/aab_env.R:24-31, for parameter-independent terms
    const double p = a * 1000.0;
    return (std::fmax(p, 0.0) + std::log1p(std::exp(-std::fabs(p)))) / 1000.0;
}

// column -> warp assignment for W warps: longest-processing-time greedy; returns the efficiency sum / (W * max load)
inline double assign_columns(const std::vector<double>& cost, int W, std::vector<std::vector<int>>* out) {
    std::vector<int> order(cost.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cost[x] > cost[y]; });
    std::vector<double> load(W, 0.0);
    std::vector<std::vector<int>> cols(W);
    double sum = 0.0;
    for (int c : order) {
        int w = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        load[w] += cost[c]; cols[w].push_back(c); sum += cost[c];
    }
    for (auto& v : cols) std::sort(v.begin(), v.end());
    if (out) *out = cols;
    const double mx = *std::max_element(load.begin(), load.end());
    return mx > 0 ? sum / (W * mx) : 1.0;
}

// Split the columns of every stock into units of P.seglen[s] rows and lay out what depends on that partition: ghost rows,
// understocking sums per unit, partial suitable biomass per (pair, unit) with the summation lists of the totals.
inline void plan_build_units(TilePlan& P) {
    TArgs& A = P.a;
    P.tb.resize(P.tb_fixed);
    int ng = P.g_fixed;
    auto takeg = [&](int n) { int o = ng; ng += n; return o; };
    const int NWc = A.NWc;
    P.unit.clear(); P.unit_cost.clear();
    A.any_ghost = 0;
    int max_len = 1;
    for (int s = 0; s < A.NS; s++) {
        StockRec& r = P.st[s];
        r.u0 = (int)P.unit.size();
        for (int col = 0; col < r.ncol; col++) {
            const int gc = r.gcol0 + col;
            for (int lo = 0; lo < r.L; lo += P.seglen[s]) {
                UnitRec u{};
                u.gc = gc; u.lo = lo; u.len = std::min(P.seglen[s], r.L - lo); u.lu = (int)P.unit.size() - r.u0;
                u.g_ghost = -1; u.up_ghost = -1;
                if (lo > 0) {
                    u.g_ghost = takeg(2 * (NWc - 1));
                    P.unit.back().up_ghost = u.g_ghost;
                    A.any_ghost = 1;
                }
                u.g_usp = takeg(2);
                u.g_sps = r.sp_ioff >= 0 ? takeg(1) : -1;
                max_len = std::max(max_len, u.len);
                P.unit.push_back(u);
                P.unit_cost.push_back(P.row_cost[gc] * ((u.len + NBX - 1) / NBX * NBX) + UNIT_OVERHEAD);
            }
        }
        r.nu = (int)P.unit.size() - r.u0;
    }
    A.NU = (int)P.unit.size();
    // partial suitable biomass per (pair, prey unit); summation lists of the totals in the reference's order: prey stock, pair, cell
    for (int q = 0; q < A.NPP; q++) P.pp[q].g_part = takeg(P.st[P.pp[q].prey].nu * P.pp[q].PL);
    for (int k = 0; k < A.NTS; k++) {
        TsRec& T = P.ts[k];
        T.list_off = (int)P.tb.size(); T.list_n = 0;
        for (int s = 0; s < A.NS; s++)
            for (int q : P.pp_of_stock[s]) {
                const PairRec& r = P.pp[q];
                if (r.pred_kind != G3PRED_TOTALFLEET && r.pred_kind != G3PRED_NUMBERFLEET) continue;
                for (int u = 0; u < P.st[s].nu; u++) {
                    const int col = P.col[P.unit[P.st[s].u0 + u].gc].col;
                    if (P.tb[r.tsid_off + col / P.st[s].nage] == k) { P.tb.push_back(r.g_part + u); T.list_n++; }
                }
            }
    }
    for (int p = 0; p < A.NPRED; p++) {
        PredRec& Pr = P.pred[p];
        if (Pr.kind != G3PRED_STOCK) continue;
        const StockRec& Pp = P.st[Pr.stock];
        std::vector<std::vector<int>> lists(Pp.narea);
        for (int s = 0; s < A.NS; s++)
            for (int q : P.pp_of_stock[s]) {
                const PairRec& r = P.pp[q];
                if (r.pred != p) continue;
                for (int u = 0; u < P.st[s].nu; u++) {
                    const int col = P.col[P.unit[P.st[s].u0 + u].gc].col;
                    const int pr = P.tb[r.tsid_off + col / P.st[s].nage];
                    if (pr >= 0) lists[pr].push_back(r.g_part + u * r.PL);
                }
            }
        Pr.list_off = (int)P.tb.size();
        P.tb.resize(P.tb.size() + 2 * Pp.narea);
        for (int pr = 0; pr < Pp.narea; pr++) {
            P.tb[Pr.list_off + 2 * pr] = (int)P.tb.size();
            P.tb[Pr.list_off + 2 * pr + 1] = (int)lists[pr].size();
            for (int e : lists[pr]) P.tb.push_back(e);
        }
    }
    // per-warp scratch: the sources of a block of migrating rows
    A.cf_rows = (max_len + NBX - 1) / NBX * NBX;
    A.wscr_n = std::max(2 * P.maxnarea * NBX, 1);
    A.g_wscr = ng;
    // the work queues hand the units out most expensive first
    std::vector<int> order(P.unit.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return P.unit_cost[x] > P.unit_cost[y]; });
    A.uorder_off = (int)P.tb.size();
    for (int u : order) P.tb.push_back(u);
    P.tb_units = (int)P.tb.size();
}

// every record table and tb in one block of ints, sections aligned to 16 bytes
inline void plan_pack_meta(TilePlan& P) {
    TArgs& A = P.a;
    P.meta.clear();
    auto put = [&](const void* data, size_t bytes) {
        while (P.meta.size() % 4) P.meta.push_back(0);
        const int off = (int)P.meta.size();
        const size_t n = (bytes + 3) / 4;
        P.meta.resize(off + std::max<size_t>(n, 1), 0);
        if (bytes) std::memcpy(P.meta.data() + off, data, bytes);
        return off;
    };
    A.m_st = put(P.st.data(), P.st.size() * sizeof(StockRec));
    A.m_col = put(P.col.data(), P.col.size() * sizeof(ColRec));
    A.m_unit = put(P.unit.data(), P.unit.size() * sizeof(UnitRec));
    A.m_pp = put(P.pp.data(), P.pp.size() * sizeof(PairRec));
    A.m_pred = put(P.pred.data(), P.pred.size() * sizeof(PredRec));
    A.m_ts = put(P.ts.data(), P.ts.size() * sizeof(TsRec));
    A.m_comp = put(P.comp.data(), P.comp.size() * sizeof(CompRec));
    A.m_xb = put(P.xb.data(), P.xb.size() * sizeof(XRec));
    A.m_agedst = put(P.agedst.data(), P.agedst.size() * sizeof(int4));
    A.m_tb = put(P.tb.data(), P.tb.size() * sizeof(int32_t));
    A.m_cidx = put(P.cterm_idx.data(), P.cterm_idx.size() * sizeof(int32_t));       // collect terms: cell, lgmatrix coefficient
    A.m_ccoef = put(P.cterm_coef.data(), P.cterm_coef.size() * sizeof(double));
    while (P.meta.size() % 4) P.meta.push_back(0);
    A.meta_n = (int)P.meta.size();
    A.n_s = P.n_s;
}

// Hot per-evaluation tables in shared memory: what is left of `smem_bytes` beside the state (if it lives there) and the plan's
// own tables takes walpha l^wbeta of every growing stock first, then growth bands, the stocks with most columns first. (The L1
// cache does not help here: the tables of the 15 warps stream through it -- measured: the same speed with 28 KB of L1 as with 60.)
inline void plan_place_hot(TilePlan& P, size_t smem_bytes) {
    TArgs& A = P.a;
    const size_t entry = (size_t)NL * sizeof(double);
    const size_t state = (size_t)P.n_s * entry;
    size_t used = 4096 + (size_t)A.meta_n * sizeof(int32_t) + (state + 4096 <= smem_bytes ? state : 0);
    smem_bytes /= (size_t)std::max(1, P.ctas_per_sm);
    int nh = 0;
    auto room = [&](int n) { return used + (size_t)(nh + n) * entry <= smem_bytes; };
    for (auto& r : P.st) r.h_wq = r.h_band = r.h_bandr = -1;
    // Only where the state itself does not fit shared memory (which then stands empty): C4 + 7 %, C5 + 5 %. Beside a resident
    // state the tables bring nothing (C2: 806 k with, 830 k without): its growth pass waits on the fp64 pipe, not on these loads.
    const bool state_resident = state + 4096 <= smem_bytes * (size_t)std::max(1, P.ctas_per_sm);
    for (auto& r : P.st) {
        if (r.grow_n < 0 || state_resident || r.grow_nset > 1) continue;
        const int n = A.NWc - 1 + r.L + NBX - 1;
        if (room(n)) { r.h_wq = nh + A.NWc - 1; nh += n; }
    }
    std::vector<int> order(P.st.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return P.st[x].ncol > P.st[y].ncol; });
    for (int s : order) {
        StockRec& r = P.st[s];
        if (r.grow_n < 0 || state_resident || r.grow_nset > 1) continue;
        const int n = (r.band_percol ? r.ncol : 1) * A.ndt * r.bsz;
        if (room(n)) { r.h_band = nh; nh += n; }
        if (r.g_bandr != r.g_band && r.h_band >= 0 && room(n)) { r.h_bandr = nh; nh += n; }
    }
    A.n_h = nh;
    plan_pack_meta(P);          // (the stock records changed)
}

// set the warp count of a plan: W unit warps take units from the work queues; the collect tasks are dealt to all W + 1 warps
// (lists at the end of tb)
inline void plan_set_warps(TilePlan& P, int W) {
    W = std::max(1, std::min({W, P.a.NU, P.max_W}));
    P.tb.resize(P.tb_units);
    P.a.W = W;
    // per-warp scratch follows the other G entries (W unit warps + the keeper warp)
    P.a.n_g = P.a.g_wscr + (W + 1) * P.a.wscr_n;
    // collect tasks of every component, dealt by size to all W + 1 warps (the keeper warp takes part)
    for (int c = 0; c < P.a.NCOMP; c++) {
        const std::vector<int>& T = P.ctask[c];
        std::vector<double> cost;
        for (size_t k = 0; k < T.size(); k += 5) cost.push_back(3.0 + (T[k + 2] - T[k + 1]));
        std::vector<std::vector<int>> mine;
        assign_columns(cost, W + 1, &mine);
        const int off = (int)P.tb.size();
        P.comp[c].task_off = off;
        P.tb.resize(off + W + 2);
        for (int w = 0; w <= W; w++) {
            P.tb[off + w] = (int)P.tb.size();
            std::stable_sort(mine[w].begin(), mine[w].end(), [&](int x, int y) { return T[5 * x + 4] < T[5 * y + 4]; });   // by group
            for (int k : mine[w]) for (int i = 0; i < 5; i++) P.tb.push_back(T[5 * k + i]);
        }
        P.tb[off + W + 1] = (int)P.tb.size();
    }
    plan_pack_meta(P);
    // per-warp scratch follows the other G entries (W unit warps + the keeper warp)
    P.a.n_g = P.a.g_wscr + (W + 1) * P.a.wscr_n;
}

// rows per unit for every stock: the partition with the smallest largest warp load at `W` warps. Candidates per stock:
// whole columns, or segments of a multiple of NBX rows that hold the NW - 1 rows the segment above reads below itself.
inline void plan_choose_partition(TilePlan& P, int W) {
    const TArgs& A = P.a;
    std::vector<std::vector<int>> cand(A.NS);
    for (int s = 0; s < A.NS; s++) {
        const StockRec& r = P.st[s];
        cand[s].push_back(r.L);
        if (r.grow_n < 0) continue;
        for (int m = NBX; m < r.L; m += NBX)
            if (m >= A.NWc - 1 && (r.L + m - 1) / m <= 6) cand[s].push_back(m);
    }
    auto maxload = [&](const std::vector<int>& seg) {
        std::vector<double> cost;
        for (int s = 0; s < A.NS; s++) {
            const StockRec& r = P.st[s];
            for (int col = 0; col < r.ncol; col++)
                for (int lo = 0; lo < r.L; lo += seg[s]) {
                    const int len = std::min(seg[s], r.L - lo);
                    cost.push_back(P.row_cost[r.gcol0 + col] * ((len + NBX - 1) / NBX * NBX) + UNIT_OVERHEAD);
                }
        }
        const int w = std::max(1, std::min<int>(W, (int)cost.size()));
        double sum = 0; for (double c : cost) sum += c;
        const double eff = assign_columns(cost, w, nullptr);
        return sum / (w * eff) * ((double)W / w > 1.0 ? 1.0 + 0.25 * ((double)W / w - 1.0) : 1.0);      // fewer warps hide less latency
    };
    std::vector<int> seg(A.NS);
    for (int s = 0; s < A.NS; s++) seg[s] = P.st[s].L;
    double best = maxload(seg);
    for (int pass = 0; pass < 3; pass++) {           // coordinate descent over the stocks
        bool changed = false;
        for (int s = 0; s < A.NS; s++) {
            int keep = seg[s];
            for (int c : cand[s]) {
                std::vector<int> t = seg; t[s] = c;
                const double v = maxload(t);
                if (v < best - 1e-9) { best = v; keep = c; changed = true; }
            }
            seg[s] = keep;
        }
        if (!changed) break;
    }
    P.seglen = seg;
}

// want_W: unit warps (0 = as many as useful, at most 15). want_seg: rows per unit, 0 = the planner's choice (made for the
// automatic warp count whatever want_W is, so that changing the warp count alone never changes a sum), -1 = whole columns.
inline TilePlan make_tile_plan(const int32_t* I, const double* R, int want_W = 0, int want_seg = 0, int max_W = 15,
                               size_t smem_bytes = 232448) {
    TilePlan P;
    P.max_W = std::max(1, std::min(max_W, SHW - 1));
    TArgs& A = P.a;
    auto fail = [&](const std::string& w) { P.ok = false; P.why = w; return P; };
    A.NC = I[G3H_NCELL]; A.NS = I[G3H_NSTOCK]; A.NPP = I[G3H_NPP]; A.NPRED = I[G3H_NPRED]; A.NCOMP = I[G3H_NCOMP];
    A.NX = I[G3H_NXBUF]; A.NSLOT = I[G3H_NSLOT]; A.NPAR = I[G3H_NPAR]; A.NT = I[G3H_NT]; A.NSTEPS = I[G3H_NSTEPS];
    A.NNLL = I[G3H_NNLL]; A.NBOUND = I[G3H_NBOUND]; A.maxL = I[G3H_MAXL];
    A.spawn_mask = 0; A.sp_n = 0; A.sp_off = 0; A.xf = 0; A.need = I[G3H_NRAND] > 0 ? 8 : 0;      // random terms: with the rarer likelihood code
    if (A.NCOMP > 32) return fail("more than 32 likelihood components");
    if (A.NX > 32) return fail("more than 32 collect vectors");
    if (A.NSTEPS > 31) return fail("more than 31 steps per year");
    auto stock = [&](int s) { return I + I[G3H_STOCK_OFF] + s * G3S_LEN; };

    // ---- distinct step lengths
    A.ndt = 0;
    for (int sy = 0; sy < A.NSTEPS; sy++) {
        const int len = I[I[G3H_STEPLEN_OFF] + sy];
        int k = 0;
        while (k < A.ndt && A.dt_len[k] != len) k++;
        if (k == A.ndt) { if (A.ndt == 4) return fail("more than 4 distinct step lengths"); A.dt_len[A.ndt++] = len; }
        A.step_idt[sy] = k;
    }

    for (int s = 0; s < A.NS; s++) P.maxn = std::max(P.maxn, stock(s)[G3S_GROW_N]);
    if (P.maxn > 15) return fail("maxlengthgroupgrowth above 15");
    const int NWc = A.NWc = P.maxn > 4 ? 16 : 5;        // band width of the kernel instantiation that runs this model
    int ng = 0, ns = 0;
    auto takeg = [&](int n) { int o = ng; ng += n; return o; };
    auto takes = [&](int n) { int o = ns; ns += n; return o; };
    // entries zeroed once per tile: (space 0 = G / 1 = S, first entry, count)
    std::vector<int> zr;
    auto zero_g = [&](int e0, int n) { if (n > 0) { zr.push_back(0); zr.push_back(e0); zr.push_back(n); } };
    auto zero_s = [&](int e0, int n) { if (n > 0) { zr.push_back(1); zr.push_back(e0); zr.push_back(n); } };
    // a table read by source row: NW - 1 zero rows below row 0 (the growth gather reads them against zero band entries)
    // and NBX - 1 above the last row (the dropped rows of a partial top block)
    auto take_padded = [&](int n) { const int o = takeg(NWc - 1 + n + NBX - 1); zero_g(o, NWc - 1); zero_g(o + NWc - 1 + n, NBX - 1); return o + NWc - 1; };
    A.g_slot = takeg(A.NSLOT);
    zero_s(takes(NWc - 1), NWc - 1);
    A.s_num = takes(A.NC); A.s_wgt = takes(A.NC);

    // ---- stocks and columns
    P.st.resize(A.NS);
    int maxnarea = 1, ncols = 0;
    std::vector<int> stock_of_cell(A.NC), col_of_cell(A.NC), l_of_cell(A.NC);
    for (int s = 0; s < A.NS; s++) {
        const int32_t* St = stock(s);
        StockRec& r = P.st[s];
        r.L = St[G3S_L]; r.nage = St[G3S_NAGE]; r.narea = St[G3S_NAREA]; r.ncol = r.nage * r.narea; r.c0 = St[G3S_CELL_OFF];
        r.gcol0 = ncols; ncols += r.ncol; r.minage = St[G3S_MINAGE];
        r.grow_n = St[G3S_GROW_N]; r.mat_kind = St[G3S_MAT_KIND]; r.has_out = St[G3S_N_OUT] > 0; r.trans_mask = St[G3S_TRANS_MASK];
        r.inc_steps = 0;
        r.mort_ioff = St[G3S_MORT_OFF]; r.midlen_roff = St[G3S_MIDLEN_ROFF]; r.plusdl_roff = St[G3S_PLUSDL_ROFF];
        r.grow_ioff = St[G3S_GROW_OFF]; r.mat_ioff = St[G3S_MAT_OFF]; r.init_ioff = St[G3S_INIT_OFF];
        r.n_renew = St[G3S_N_RENEW]; r.renew_ioff = St[G3S_RENEW_OFF];
        r.ren_off = (int)P.tb.size();
        for (int k = 0; k < r.n_renew; k++) {
            const int32_t* Rn = I + r.renew_ioff + k * G3R_LEN;
            P.tb.push_back(Rn[G3R_RUN_STEP]); P.tb.push_back(Rn[G3R_AGE_IDX]); P.tb.push_back(Rn[G3R_FACTOR_OFF]);
        }
        if (r.n_renew > 30) return fail("more than 30 renewal actions on one stock");
        maxnarea = std::max(maxnarea, r.narea);
        // time-varying M: one table per (slot set, step length)
        r.mort_tmap = St[G3S_MORT_TMAP_OFF]; r.mort_nset = 1;
        if (r.mort_ioff >= 0 && r.mort_tmap >= 0) for (int t = 0; t < A.NT; t++) r.mort_nset = std::max(r.mort_nset, I[r.mort_tmap + t] + 1);
        r.g_mort = takeg(r.mort_ioff >= 0 ? r.mort_nset * r.ncol * A.ndt : 0);
        if (r.mort_ioff >= 0 && r.mort_tmap >= 0) A.xf = 1;
        r.g_rd = takeg(r.n_renew * r.L); r.g_rw = takeg(r.n_renew * r.L);
        // time-varying growth parameters: one copy of walpha l^wbeta and of the bands per slot set
        r.grow_tmap = St[G3S_GROW_TMAP_OFF]; r.grow_nset = 1;
        if (r.grow_n >= 0 && r.grow_tmap >= 0) { A.xf = 1; for (int t = 0; t < A.NT; t++) r.grow_nset = std::max(r.grow_nset, I[r.grow_tmap + t] + 1); }
        r.wq_stride = NWc - 1 + r.L + NBX - 1;
        r.g_wq = 0;
        if (r.grow_n >= 0) for (int gsn = 0; gsn < r.grow_nset; gsn++) { const int o = take_padded(r.L); if (gsn == 0) r.g_wq = o; }
        const bool trans = r.grow_n >= 0 && r.has_out;
        const bool cont = trans && r.mat_kind == G3MAT_CONTINUOUS;
        r.band_percol = cont && I[r.mat_ioff + 2] >= 0;         // age term: M0 differs per column (R/action_mature.R:18,44-74)
        r.bsz = r.grow_n >= 0 ? (r.L + NBX - 1) * NWc : 0;
        const int nset = (r.band_percol ? r.ncol : 1) * r.grow_nset;
        r.g_band = takeg(nset * A.ndt * r.bsz);
        zero_g(r.g_band, nset * A.ndt * r.bsz);
        r.g_bandr = cont ? takeg(nset * A.ndt * r.bsz) : r.g_band;
        if (cont) zero_g(r.g_bandr, nset * A.ndt * r.bsz);
        r.g_cmat = (trans && r.mat_kind == G3MAT_CONSTANT) ? take_padded(r.ncol * r.L) : -1;
        r.mig_ioff = St[G3S_MIGRATE_OFF]; r.mig_steps = 0;
        if (r.mig_ioff >= 0) A.need |= 1;
        if (r.grow_n >= 0 && r.has_out && r.mat_kind == G3MAT_CONSTANT) A.need |= 4;
        r.of_ioff = St[G3S_OTHERFOOD_OFF];
        // g3a_spawn (R/action_spawn.R:86-233): proportion spawning and spawning mortality by length, the fecundity weights
        r.sp_ioff = St[G3S_SPAWN_OFF]; r.g_sprop = r.g_smort = r.g_sfec = -1; r.sp_out_off = -1;
        if (r.sp_ioff >= 0 || r.of_ioff >= 0) A.xf = 1;
        if (r.sp_ioff >= 0) {
            const int32_t* Sp = I + r.sp_ioff;
            r.g_sprop = takeg(r.L);
            if (Sp[G3SP_MORT_KIND] >= 0) r.g_smort = takeg(r.L);
            if (Sp[G3SP_RECR_KIND] == G3SPAWN_FECUNDITY) r.g_sfec = takeg(r.nage * r.L);
            if (Sp[G3SP_RECR_KIND] < 0 || Sp[G3SP_RECR_KIND] > G3SPAWN_HOCKEYSTICK) return fail("unknown recruitment function");
            A.spawn_mask |= Sp[G3SP_RUN_STEP] == 0 ? (1 << A.NSTEPS) - 1 : 1 << (Sp[G3SP_RUN_STEP] - 1);
        }
        r.g_mig = r.mig_ioff >= 0 ? takeg(A.NSTEPS * r.narea * r.narea) : 0;
        if (r.mig_ioff >= 0)
            for (int sy = 0; sy < A.NSTEPS; sy++) for (int i = 0; i < r.narea * r.narea; i++)
                if (I[r.mig_ioff + sy * r.narea * r.narea + i] >= 0) r.mig_steps |= 1 << sy;
        r.age_enable = St[G3S_AGE_ENABLE]; r.age_n_out = St[G3S_AGE_N_OUT]; r.s_mv = -1;
        r.us_flag = 0;
        for (int col = 0; col < r.ncol; col++) for (int l = 0; l < r.L; l++) {
            const int c = r.c0 + col * r.L + l;
            stock_of_cell[c] = s; col_of_cell[c] = col; l_of_cell[c] = l;
        }
    }
    for (int i = 0; i < I[G3H_US_N]; i++) P.st[I[I[G3H_US_OFF] + i]].us_flag = 1;
    A.NCOLS = ncols;

    // ---- maturation hand-over (R/action_mature.R:78-134) per cell, then per column
    std::vector<int> inc_src(A.NC, -1), inc_slot(A.NC, -1), inc_mask(A.NC, 0);
    std::vector<std::vector<int>> rem(A.NC);
    for (int s = 0; s < A.NS; s++) {
        const int32_t* St = stock(s);
        const StockRec& r = P.st[s];
        if (!(r.grow_n >= 0 && r.has_out)) continue;
        const int ncell = r.L * r.ncol;
        for (int i = 0; i < ncell; i++)
            for (int o = 0; o < St[G3S_N_OUT]; o++) {
                const int d = I[St[G3S_OUT_MAP_OFF] + o * ncell + i];
                if (d < 0) continue;
                const int slot = I[St[G3S_OUT_OFF] + 2 * o + 1];
                if (inc_src[d] >= 0) return fail("a cell receives maturing fish from more than one source");
                inc_src[d] = r.c0 + i; inc_slot[d] = slot; inc_mask[d] = r.trans_mask;
                rem[r.c0 + i].push_back(slot);
            }
    }
    P.col.resize(ncols);
    A.inc_any_mask = 0;
    int ntr = 0;        // S entries of the maturing fish in flight (tn then tw per stashing column)
    const int s_x0 = ns;        // start of the transient region (tn/tw; reused by the ageing movement stash)
    for (int s = 0; s < A.NS; s++) {
        const StockRec& r = P.st[s];
        for (int col = 0; col < r.ncol; col++) {
            ColRec& c = P.col[r.gcol0 + col];
            c.s = s; c.col = col; c.aidx = col % r.nage; c.ridx = col / r.nage; c.cell0 = r.c0 + col * r.L;
            c.s_tn = c.s_tw = -1; c.inc_tn = c.inc_tw = c.inc_slot = -1; c.inc_mask = 0;
            c.g_remf = takeg(1);
            // remainder ratios must be the same down a column
            for (int l = 1; l < r.L; l++)
                if (rem[c.cell0 + l] != rem[c.cell0]) return fail("maturation outputs differ between the lengths of one column");
            c.rem_off = (int)P.tb.size();
            for (int sl : rem[c.cell0]) P.tb.push_back(sl);
            P.tb.push_back(-1);
            if (!rem[c.cell0].empty()) { c.s_tn = s_x0 + ntr; c.s_tw = s_x0 + ntr + r.L; ntr += 2 * r.L; }
        }
    }
    for (int gc = 0; gc < ncols; gc++) {        // arrivals: every cell of a destination column is fed by the same length of ONE source column
        ColRec& c = P.col[gc];
        const StockRec& r = P.st[c.s];
        const int c00 = c.cell0;
        if (inc_src[c00] >= 0) {
            const int src = inc_src[c00];
            if (l_of_cell[src] != 0) return fail("maturing fish change length group on hand-over");
            const ColRec& sc = P.col[P.st[stock_of_cell[src]].gcol0 + col_of_cell[src]];
            if (P.st[sc.s].L != r.L) return fail("maturation between stocks of different length grids");
            c.inc_tn = sc.s_tn; c.inc_tw = sc.s_tw; c.inc_slot = inc_slot[c00]; c.inc_mask = inc_mask[c00];
            A.inc_any_mask |= c.inc_mask;
        }
        for (int l = 0; l < r.L; l++) {
            const int cc = c00 + l;
            const bool ok = inc_src[c00] < 0 ? inc_src[cc] < 0
                                             : (inc_src[cc] == inc_src[c00] + l && inc_slot[cc] == inc_slot[c00] && inc_mask[cc] == inc_mask[c00]);
            if (!ok) return fail("maturation hand-over is not column to column");
        }
    }
    A.inc_off = (int)P.tb.size(); A.inc_n = 0;
    for (int gc = 0; gc < ncols; gc++) if (P.col[gc].inc_tn >= 0) { P.tb.push_back(gc); A.inc_n++; }
    // ---- spawning: the parents, and per output stock (stock, normalised length distribution, weights, ratio slot)
    A.sp_off = (int)P.tb.size(); A.sp_n = 0;
    for (int s = 0; s < A.NS; s++) if (P.st[s].sp_ioff >= 0) { P.tb.push_back(s); A.sp_n++; }
    for (int s = 0; s < A.NS; s++) {
        StockRec& r = P.st[s];
        if (r.sp_ioff < 0) continue;
        const int32_t* Sp = I + r.sp_ioff;
        r.sp_out_off = (int)P.tb.size();
        for (int o = 0; o < Sp[G3SP_N_OUT]; o++) {
            const int32_t* Or = I + Sp[G3SP_OUT_OFF] + 6 * o;
            const StockRec& So = P.st[Or[0]];
            // the offspring land after the barrier that follows pass 2, beside the blocks that receive maturing fish: not the same cells
            for (int ar = 0; ar < So.narea; ar++)
                if (P.col[So.gcol0 + ar * So.nage].inc_tn >= 0) return fail("spawning into a column that also receives maturing fish");
            P.tb.push_back(Or[0]); P.tb.push_back(takeg(So.L)); P.tb.push_back(takeg(So.L)); P.tb.push_back(Or[5]);
        }
    }
    // one renewal per (column, step) is what finish_cell applies in a defined order
    for (int s = 0; s < A.NS; s++) {
        const StockRec& r = P.st[s];
        for (int k = 0; k < r.n_renew; k++) for (int k2 = k + 1; k2 < r.n_renew; k2++) {
            const int32_t* Ra = I + r.renew_ioff + k * G3R_LEN; const int32_t* Rb = I + r.renew_ioff + k2 * G3R_LEN;
            if (Ra[G3R_AGE_IDX] == Rb[G3R_AGE_IDX] && (Ra[G3R_RUN_STEP] == 0 || Rb[G3R_RUN_STEP] == 0 || Ra[G3R_RUN_STEP] == Rb[G3R_RUN_STEP])) {
                // ... unless they never run in the same year (one renewal whose shape parameters vary by year, split into a
                // record per set of years: gadget3_b200/run_cuda.py)
                bool together = false;
                const int nya = (I[G3H_END_YEAR] - I[G3H_START_YEAR] + 1) * r.narea;
                for (int i = 0; i < nya; i++) together = together || (I[Ra[G3R_FACTOR_OFF] + i] >= 0 && I[Rb[G3R_FACTOR_OFF] + i] >= 0);
                if (together) return fail("two renewals into the same age at the same step");
            }
        }
    }
    // ---- ageing movement (R/action_age.R:16-20,57-63): stash of the oldest column per (area, l), destinations
    int nmv = 0;
    for (int s = 0; s < A.NS; s++) {
        StockRec& r = P.st[s];
        if (r.age_enable && r.age_n_out > 0) { r.s_mv = s_x0 + nmv; nmv += 2 * r.narea * r.L; }
    }
    takes(std::max(ntr, nmv));
    zero_s(takes(NBX - 1), NBX - 1);        // (rows the dropped part of the last column's top block reads)
    {
        std::vector<int> age_src(A.NC, -1);
        std::vector<int32_t> agedst_l;
        for (int s = 0; s < A.NS; s++) {
            const int32_t* St = stock(s);
            const StockRec& r = P.st[s];
            if (!r.age_enable) continue;
            for (int o = 0; o < r.age_n_out; o++) for (int ar = 0; ar < r.narea; ar++) for (int l = 0; l < r.L; l++) {
                const int d = I[St[G3S_AGE_MAP_OFF] + (o * r.narea + ar) * r.L + l];
                if (d < 0) continue;
                if (age_src[d] >= 0) return fail("a cell receives ageing fish from more than one source");
                if (l_of_cell[d] != l) return fail("ageing fish change length group on hand-over");
                age_src[d] = 1;
                const int i = ar * r.L + l;
                P.agedst.push_back(make_int4(d, r.s_mv + i, r.s_mv + r.narea * r.L + i, I[St[G3S_AGE_OUT_OFF] + 2 * o + 1]));
                agedst_l.push_back(l);
            }
        }
        // the kernel walks destinations in cell order within a length group (the thread kernel's order)
        std::vector<int> ord(P.agedst.size());
        for (size_t i = 0; i < ord.size(); i++) ord[i] = (int)i;
        std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) {
            return agedst_l[x] != agedst_l[y] ? agedst_l[x] < agedst_l[y] : P.agedst[x].x < P.agedst[y].x; });
        std::vector<int4> ad2; std::vector<int32_t> al2;
        for (int i : ord) { ad2.push_back(P.agedst[i]); al2.push_back(agedst_l[i]); }
        P.agedst = ad2;
        A.n_agedst = (int)P.agedst.size();
        A.agedst_l_off = (int)P.tb.size();      // first entry of each length group
        for (int l = 0; l <= A.maxL; l++) {
            int k = 0;
            while (k < (int)al2.size() && al2[k] < l) k++;
            P.tb.push_back(k);
        }
        if (P.agedst.empty()) P.agedst.push_back(make_int4(-1, 0, 0, 0));
    }

    // ---- collect vectors
    P.xb.resize(std::max(A.NX, 1));
    std::vector<unsigned> x_ppmask(std::max(A.NX, 1), 0);
    for (int x = 0; x < A.NX; x++) {
        const int32_t* X = I + I[G3H_XBUF_OFF] + x * G3X_LEN;
        P.xb[x].kind = X[G3X_KIND]; P.xb[x].unit = X[G3X_UNIT]; P.xb[x].g_x = takeg(A.NC);
        for (int j = 0; j < X[G3X_N_PP]; j++) {
            const int q = I[X[G3X_PP_OFF] + j];
            if (q >= 32) return fail("more than 32 predator-prey pairs feed collect vectors");
            x_ppmask[x] |= 1u << q;
        }
    }
    A.ax_off = (int)P.tb.size(); A.ax_n = 0;
    for (int x = 0; x < A.NX; x++) if (P.xb[x].kind != 0) { P.tb.push_back(x); A.ax_n++; }

    // ---- predators, pairs, total-suitability accumulators
    P.pred.resize(std::max(A.NPRED, 1));
    std::vector<int> pred_ts_base(std::max(A.NPRED, 1), -1);
    int nts = 0;
    for (int p = 0; p < A.NPRED; p++) {
        const int32_t* Pr = I + I[G3H_PRED_OFF] + p * G3P_LEN;
        PredRec& r = P.pred[p];
        r.kind = Pr[G3P_KIND]; r.stock = Pr[G3P_STOCK]; r.slot_ioff = Pr[G3P_SLOT_OFF]; r.narea = Pr[G3P_NAREA];
        r.g_pts = r.g_pmc = 0; r.list_off = 0;
        if (r.kind == G3PRED_STOCK) {
            const StockRec& Pp = P.st[r.stock];
            r.g_pts = takeg(Pp.narea * Pp.L); r.g_pmc = takeg(Pp.L);
            continue;
        }
        if (r.kind != G3PRED_TOTALFLEET && r.kind != G3PRED_NUMBERFLEET && r.kind != G3PRED_LINEARFLEET && r.kind != G3PRED_EFFORTFLEET)
            return fail("predator kind " + std::to_string(r.kind) + " is not supported");
        pred_ts_base[p] = nts;
        for (int k = 0; k < r.narea; k++) P.ts.push_back(TsRec{r.kind, Pr[G3P_E_ROFF] + k, r.narea, 0, 0});
        nts += r.narea;
    }
    A.NTS = nts;
    A.g_eot = takeg(nts);
    P.pp.resize(std::max(A.NPP, 1));
    std::vector<std::vector<int>> pp_of_stock(A.NS);
    for (int q = 0; q < A.NPP; q++) {
        const int32_t* Q = I + I[G3H_PP_OFF] + q * G3PP_LEN;
        const PredRec& Pr = P.pred[Q[G3PP_PRED]];
        const StockRec& Ps = P.st[Q[G3PP_PREY]];
        PairRec& r = P.pp[q];
        r.pred = Q[G3PP_PRED]; r.prey = Q[G3PP_PREY]; r.suit_kind = Q[G3PP_SUIT_KIND]; r.suit_ioff = Q[G3PP_SUIT_OFF];
        r.energy_slot = Q[G3PP_ENERGY]; r.pref_slot = Q[G3PP_PREF]; if (r.pref_slot >= 0) A.xf = 1;
        if (Pr.kind != G3PRED_TOTALFLEET) A.need |= 2; r.pred_kind = Pr.kind; r.pred_stock = Pr.kind == G3PRED_STOCK ? Pr.stock : -1;
        r.PL = Pr.kind == G3PRED_STOCK ? P.st[Pr.stock].L : 1;
        const int sk = r.suit_kind;
        if (sk == G3SUIT_ANDERSEN || sk == G3SUIT_EXPONENTIAL || sk == G3SUIT_RICHARDS) {
            if (Pr.kind != G3PRED_STOCK) return fail("suitability kind " + std::to_string(sk) + " reads predator_length: it needs a stock predator");
        } else if (sk < 0 || sk > G3SUIT_RICHARDS) return fail("suitability kind " + std::to_string(sk) + " is not supported");
        r.suit_tmap = Q[G3PP_SUIT_TMAP]; r.suit_nset = 1;
        if (r.suit_tmap >= 0) { A.xf = 1; for (int t = 0; t < A.NT; t++) r.suit_nset = std::max(r.suit_nset, I[r.suit_tmap + t] + 1); }
        r.g_suit = takeg(r.suit_nset * r.PL * Ps.L);
        r.g_part = 0;
        r.cxmask = 0;
        r.tsid_off = (int)P.tb.size();
        for (int ar = 0; ar < Ps.narea; ar++) {
            const int area = I[stock(r.prey)[G3S_AREA_OFF] + ar];
            int id = -1;
            const int32_t* Prr = I + I[G3H_PRED_OFF] + r.pred * G3P_LEN;
            for (int k = 0; k < Pr.narea; k++)
                if (I[Prr[G3P_AREA_OFF] + k] == area) id = Pr.kind == G3PRED_STOCK ? k : pred_ts_base[r.pred] + k;
            P.tb.push_back(id);
        }
        pp_of_stock[r.prey].push_back(q);
    }
    for (int s = 0; s < A.NS; s++) {
        StockRec& r = P.st[s];
        r.pp_off = (int)P.tb.size(); r.pp_n = (int)pp_of_stock[s].size();
        for (int q : pp_of_stock[s]) P.tb.push_back(q);
        // catch collect vectors fed by this stock's pairs
        r.cx_off = (int)P.tb.size(); r.cx_n = 0;
        for (int x = 0; x < A.NX; x++) {
            if (P.xb[x].kind != 0) continue;
            bool fed = false;
            for (int q : pp_of_stock[s]) if ((x_ppmask[x] >> q) & 1) fed = true;
            if (!fed) continue;
            if (r.cx_n == 32) return fail("more than 32 catch collect vectors on one stock");
            for (int q : pp_of_stock[s]) if ((x_ppmask[x] >> q) & 1) P.pp[q].cxmask |= 1 << r.cx_n;
            P.tb.push_back(x); r.cx_n++;
        }
    }
    if (P.ts.empty()) P.ts.push_back(TsRec{0, 0, 0, 0, 0});
    // steps with predation: some landing is non-zero, or a stock predator exists (it eats every step)
    P.pred_active.assign(A.NT, 0);
    for (int p = 0; p < A.NPRED; p++) {
        const int32_t* Pr = I + I[G3H_PRED_OFF] + p * G3P_LEN;
        if (Pr[G3P_KIND] == G3PRED_STOCK) { std::fill(P.pred_active.begin(), P.pred_active.end(), 1); continue; }
        for (int k = 0; k < Pr[G3P_NAREA]; k++)
            for (int t = 0; t < A.NT; t++)
                if (R[Pr[G3P_E_ROFF] + (size_t)t * Pr[G3P_NAREA] + k] != 0.0) P.pred_active[t] = 1;
    }

    // ---- likelihood components and their parameter-independent observation terms
    P.comp.resize(std::max(A.NCOMP, 1));
    for (int c = 0; c < A.NCOMP; c++) {
        const int32_t* Cm = I + I[G3H_COMP_OFF] + c * G3C_LEN;
        CompRec& r = P.comp[c];
        if (Cm[G3C_FUNC] != G3F_SUMOFSQUARES && Cm[G3C_FUNC] != G3F_SI_LOG) A.need |= 8;
        r.func = Cm[G3C_FUNC]; r.weight_par = Cm[G3C_WEIGHT_PAR]; r.xbuf = Cm[G3C_XBUF]; r.mode = Cm[G3C_MODE];
        r.n_dest = Cm[G3C_N_DEST]; r.n_bins = Cm[G3C_N_BINS]; r.n_slices = Cm[G3C_N_SLICES]; r.n_areag = Cm[G3C_N_AREAG];
        r.csr_ptr_ioff = Cm[G3C_CSR_PTR_OFF]; r.csr_idx_ioff = Cm[G3C_CSR_IDX_OFF]; r.csr_coef_roff = Cm[G3C_CSR_COEF_ROFF];
        r.cell_dest_ioff = Cm[G3C_CELL_DEST_OFF]; r.cell_coef_roff = Cm[G3C_CELL_COEF_ROFF];
        r.tidx_ioff = Cm[G3C_TIDX_OFF]; r.cmp_ioff = Cm[G3C_CMP_OFF]; r.n_times = Cm[G3C_N_TIMES]; r.param_roff = Cm[G3C_PARAM_ROFF];
        if (r.func < 0 || r.func > G3F_SUMSQLOGRATIOS) return fail("likelihood function " + std::to_string(r.func) + " is not supported");
        const bool si = r.func == G3F_SI_LOG || r.func == G3F_SI_LINEAR;
        r.g_ms = takeg(r.n_dest * (si ? r.n_times : 1));
        // groups of dests whose total the comparison divides by
        // groups of dests that share a total: sumofsquares the area group's whole slab, or -- stratified -- one length group
        // across the slices; multinomial every slice
        const bool strat = r.func == G3F_SUMOFSQUARES && R[r.param_roff] != 0.0;
        std::vector<int> grp(r.n_dest, 0);
        r.n_groups = 0;
        if (strat) {
            r.n_groups = r.n_areag * r.n_bins;
            for (int e = 0; e < r.n_dest; e++) grp[e] = (e / (r.n_slices * r.n_bins)) * r.n_bins + e % r.n_bins;
        } else if (r.func == G3F_SUMOFSQUARES) {
            r.n_groups = r.n_areag;
            for (int e = 0; e < r.n_dest; e++) grp[e] = e / (r.n_slices * r.n_bins);
        } else if (r.func == G3F_MULTINOMIAL) {
            r.n_groups = r.n_areag * r.n_slices;
            for (int e = 0; e < r.n_dest; e++) grp[e] = e / r.n_bins;
        }
        r.grp_off = (int)P.tb.size();
        for (int v : grp) P.tb.push_back(v);
        r.g_pt = takeg(r.n_groups * SHW); r.g_pv = takeg(SHW);
        // collect terms per dest, cut into tasks
        {
            std::vector<std::vector<std::pair<int, double>>> rows(r.n_dest);
            if (r.mode == 0) {
                const int32_t* ptr = I + r.csr_ptr_ioff; const int32_t* idx = I + r.csr_idx_ioff; const double* cf = R + r.csr_coef_roff;
                for (int e = 0; e < r.n_dest; e++) for (int j = ptr[e]; j < ptr[e + 1]; j++) rows[e].push_back({idx[j], cf[j]});
            } else {
                const int32_t* cd = I + r.cell_dest_ioff; const double* cf = R + r.cell_coef_roff;
                for (int i = 0; i < A.NC; i++) if (cd[i] >= 0 && cd[i] < r.n_dest) rows[cd[i]].push_back({i, cf[i]});
            }
            std::vector<int> tasks, split;
            int nslot = 0, nsplit = 0;
            for (int e = 0; e < r.n_dest; e++) {
                const int j0 = (int)P.cterm_idx.size();
                for (auto& tc : rows[e]) { P.cterm_idx.push_back(tc.first); P.cterm_coef.push_back(tc.second); }
                const int n = (int)rows[e].size(), ge = grp[e];
                if (n <= TASK_TERMS) { tasks.insert(tasks.end(), {e, j0, j0 + n, -1, ge}); continue; }
                const int parts = (n + TASK_TERMS - 1) / TASK_TERMS;
                split.insert(split.end(), {e, nslot, parts}); nsplit++;
                for (int k = 0; k < parts; k++) {
                    const int a = j0 + (int)((long long)n * k / parts), b = j0 + (int)((long long)n * (k + 1) / parts);
                    tasks.insert(tasks.end(), {e, a, b, k == 0 ? -2 - nslot : nslot, ge});
                    nslot++;
                }
            }
            r.g_cpart = takeg(nslot);
            r.split_off = (int)P.tb.size();
            P.tb.push_back(nsplit);
            for (int v : split) P.tb.push_back(v);
            P.ctask.push_back(tasks);
        }
        r.obs_off = (int)P.obsprop.size(); r.obsconst_off = (int)P.obsconst.size();
        const double* obs = R + Cm[G3C_OBS_ROFF];
        const int nd = r.n_dest, nb = r.n_bins, nsl = r.n_slices, nag = r.n_areag;
        for (int t = 0; t < r.n_times; t++) {
            double cst = 0.0;
            for (int ag = 0; ag < nag; ag++) {
                double tot = 0.0;
                for (int i = 0; i < nsl * nb; i++) tot += obs[(size_t)t * nd + ag * nsl * nb + i];
                std::vector<double> btot(nb, 0.0);          // stratified: totals per length group
                for (int i = 0; i < nsl * nb; i++) btot[i % nb] += obs[(size_t)t * nd + ag * nsl * nb + i];
                for (int i = 0; i < nsl * nb; i++) {
                    const double o = obs[(size_t)t * nd + ag * nsl * nb + i];
                    const double az = plan_avoid_zero(strat ? btot[i % nb] : tot);
                    double v = o;
                    if (r.func == G3F_SUMOFSQUARES) v = o / az;                          // R/likelihood_distribution.R:12-27
                    else if (r.func == G3F_SI_LOG) v = std::log(plan_avoid_zero(o));     // R/likelihood_distribution.R:86
                    else if (r.func == G3F_SUMSQLOGRATIOS) v = std::log(o + R[r.param_roff]);    // R/likelihood_distribution.R:131
                    P.obsprop.push_back(v);
                }
                if (r.func == G3F_MULTINOMIAL)
                    for (int sl = 0; sl < nsl; sl++) {      // sum(lgamma(1 + data)) - lgamma(1 + sum(data))
                        double so = 0, slg = 0;
                        for (int b = 0; b < nb; b++) { const double o = obs[(size_t)t * nd + (ag * nsl + sl) * nb + b]; so += o; slg += std::lgamma(1 + o); }
                        cst += slg - std::lgamma(1 + so);
                    }
            }
            P.obsconst.push_back(cst);
        }
    }
    // per step: components with a time index (bit c), components that compare (bit c), predation on
    A.tmask_off = (int)P.tb.size();
    for (int which = 0; which < 2; which++)
        for (int t = 0; t < A.NT; t++) {
            unsigned mask = 0;
            for (int c = 0; c < A.NCOMP; c++) {
                const int v = which == 0 ? I[P.comp[c].tidx_ioff + t] : I[P.comp[c].cmp_ioff + t];
                if (which == 0 ? v >= 0 : v != 0) mask |= 1u << c;
            }
            P.tb.push_back((int32_t)mask);
        }
    for (int t = 0; t < A.NT; t++) P.tb.push_back(P.pred_active[t]);
    if (P.obsprop.empty()) P.obsprop.push_back(0.0);
    if (P.obsconst.empty()) P.obsconst.push_back(0.0);
    A.g_cn = takeg(std::max(A.NCOMP, 1));
    A.g_ctot = takeg(A.NNLL);
    A.g_bterm = takeg(A.NBOUND);
    P.n_s = ns;
    A.zr_off = (int)P.tb.size(); A.zr_n = (int)zr.size() / 3;
    for (int v : zr) P.tb.push_back(v);
    P.g_fixed = ng; P.tb_fixed = (int)P.tb.size();
    P.maxnarea = maxnarea; P.pp_of_stock = pp_of_stock;
    for (int s = 0; s < A.NS; s++) P.max_cx = std::max(P.max_cx, P.st[s].cx_n);

    // ---- cost of one row of each column per step, in avoid_zero chains (they dominate): growth 1 (+ the gather), the
    // maturing split another one when the maturing fish have a destination, predation 3 on the steps that have any
    int npred = 0;
    for (int t = 0; t < A.NT; t++) npred += P.pred_active[t] != 0;
    const double pred_frac = A.NT > 0 ? (double)npred / A.NT : 0.0;
    P.row_cost.resize(ncols);
    for (int gc = 0; gc < ncols; gc++) {
        const StockRec& r = P.st[P.col[gc].s];
        const bool split = r.grow_n >= 0 && r.has_out && r.mat_kind != G3MAT_NONE;
        double c = 0.1;
        if (r.grow_n >= 0) c = 1.0 + 0.05 * NWc * (split ? 2 : 1) + (split && P.col[gc].s_tn >= 0 ? 1.0 : 0.0);
        if (r.pp_n > 0) c += 3.3 * pred_frac;
        P.row_cost[gc] = c;
    }
    // ---- partition into units and the warp count. A small model (its state fits shared memory several times over) runs as
    // several small CTAs per SM instead of one large one: their barriers and serial phases overlap (measured on the
    // introduction-single-stock model: 4 CTAs of 3 unit warps on whole columns 4.09 M evals/s, one CTA of 15 on row
    // segments 2.95 M).
    P.seglen.resize(A.NS);
    for (int s = 0; s < A.NS; s++) P.seglen[s] = P.st[s].L;
    const size_t state_bytes = (size_t)ns * NL * sizeof(double) + 1024;
    const int fit = (int)(smem_bytes / state_bytes);
    if (want_W == 0 && want_seg == 0 && max_W <= 15 && fit >= 4) { want_W = 3; want_seg = -1; P.ctas_per_sm = 4; }
    if (want_seg == 0) plan_choose_partition(P, P.max_W);
    else if (want_seg > 0)
        for (int s = 0; s < A.NS; s++) {
            const int m = (want_seg + NBX - 1) / NBX * NBX;
            if (P.st[s].grow_n >= 0 && m >= NWc - 1 && m < P.st[s].L) P.seglen[s] = m;
        }
    plan_build_units(P);
    plan_set_warps(P, want_W > 0 ? want_W : P.max_W);
    plan_place_hot(P, smem_bytes);
    P.ok = true;
    return P;
}

}  // namespace g3t
