// g3_tile.cuh -- the production evaluation kernel: one CTA owns a TILE of 32 parameter vectors for the
// whole time loop, with the stock state resident in shared memory.
//
// Mapping (north_star item 1; baseline/multiple-substocks.cpp:692-2161 is the path it evaluates):
//   lane  = evaluation. The 32 lanes of every warp are the same 32 parameter vectors, so the per-cell
//           arithmetic is the straight-line scalar code of the reference and control flow is warp-uniform
//           (branches depend on model structure and the step, never on the lane).
//   warp  = a list of UNITS: row segments [lo, lo + len) of age/area columns of a stock, dealt to the warps by
//           cost (g3_tile_plan.h). Growth only looks DOWN the length axis of its own column, so a unit needs its own
//           rows plus the NW - 1 rows below it: the unit below copies those into the unit's GHOST rows before the
//           growth pass (one block barrier), after which every unit sweeps its rows in place, top block first.
//           What crosses columns goes through memory and one block barrier: fleet totals, the maturation hand-over,
//           likelihood sums, ageing, migration.
//   state = num/wgt of every stock cell (+ the maturing fish in flight) in shared memory, [cell][lane]:
//           a warp access is one conflict-free 256-byte row. Models whose state does not fit 227 KB
//           (and Dual<N> gradient runs) keep the same layout in the CTA's global slab instead (SMEM = false).
//   tables (per evaluation, time-invariant: growth bands, suitability, mortality, renewal shape, likelihood
//           slabs, collect vectors) live in a per-CTA global slab [entry][lane]; only gridDim x 32
//           evaluations are in flight, so the slabs stay in L2/L1 instead of streaming from HBM.
// The hot loops (growth gather, predation) are straight-line code over blocks of NB rows: every load is a block
// pointer plus a compile-time offset, the band is stored destination-oriented ([l][d] = P(l - d -> l), plus-group
// tail sums folded into row L - 1, zero where l - d < 0), and the avoid_zero / reciprocal chains of the NB rows are
// interleaved stage by stage (g3_math.cuh).
// Nothing is exchanged between lanes: no shuffles, no atomics; every sum has a fixed order.
//
// The body is written against a few macros (G3T_BAR, G3T_ANY) so that the same source also compiles under
// g++ for tests/emu (kernel-logic tests on the CPU, one OS thread per warp) -- test infrastructure, never
// loaded by the product library.
#pragma once
#include "g3_common.cuh"

namespace g3t {
using namespace g3;

constexpr int NL = 32;      // lanes = evaluations per tile
constexpr int NBX = 4;      // rows per block of the plain-double instantiations: the planner aligns row segments to it
constexpr int SHW = 32;     // most warps of a tile (unit warps + keeper): stride of the per-warp likelihood shares
#ifndef __CUDACC__
using std::max; using std::min;
#endif

struct StockRec {
    int L, nage, narea, ncol, c0, gcol0, minage;
    int grow_n, mat_kind, has_out, trans_mask, inc_steps;
    int mort_ioff, midlen_roff, plusdl_roff, grow_ioff, mat_ioff, init_ioff;
    int n_renew, renew_ioff;
    int ren_off;            // tb: per renewal (run step or 0 = every step, age index it lands in, offset in I of its factor slots per year x area)
    int g_mort;             // [set][idt][col] exp(-M dt)
    int mort_tmap, mort_nset;   // I: the slot set of M each step uses (time-varying M), or -1; number of sets
    int g_rd, g_rw;         // [k][l] renewal numbers shape (x 1e4) and weight
    int g_wq;               // [gset][l] walpha * midlen^wbeta; NW - 1 zero rows before row 0, NBX - 1 after row L - 1 (wq_stride apart)
    int grow_tmap, grow_nset, wq_stride;    // I: the set of growth parameters each step uses (time-varying growth), or -1; sets
    int g_band, g_bandr;    // destination-oriented growth bands, staying (or whole) / maturing part: [gset][set][idt][(L + NBX - 1) * NW]
    int h_wq, h_band, h_bandr;      // the same tables in the H entries (shared memory beside the state), or -1: the planner puts
                                    // the hottest tables there as far as room allows (plain-double runs only)
    int band_percol, bsz;   // one set per column (continuous maturity with an age term) or per stock
    int g_cmat;             // [col][l] g3a_mature_constant ratio (NW - 1 zero rows before, NBX - 1 after), or -1
    int mig_ioff, g_mig, mig_steps;
    int of_ioff;            // I: g3a_otherfood (num slot, wgt slot) per (step, area), or -1
    int sp_ioff;            // I: g3a_spawn record of this (parent) stock, or -1
    int g_sprop, g_smort, g_sfec;   // [l] proportion spawning; [l] 1 - exp(-spawning mortality) or -1; [age][l] midlen^p1 age^p2 or -1
    int sp_out_off;         // tb: per output stock (stock, G [l] share of the offspring, G [l] their weight, ratio slot)
    int age_enable, age_n_out, s_mv;
    int us_flag;
    int pp_off, pp_n;       // tb: predator-prey pairs acting on this stock, run order
    int cx_off, cx_n;       // tb: catch collect vectors those pairs feed
    int u0, nu;             // units of this stock: unit[u0 .. u0 + nu), column by column, bottom segment first (= cell order)
};
struct ColRec {
    int s, col, aidx, ridx, cell0;
    int s_tn, s_tw;         // S entries of this column's maturing fish, or -1 (no destination: only the remainder matters)
    int inc_tn, inc_tw, inc_slot, inc_mask;     // maturing fish arriving INTO this column: source entries, ratio slot, step mask
    int g_remf;             // share of the maturing fish no output stock claims
    int rem_off;            // tb: -1 terminated list of ratio slots with a destination
};
struct UnitRec {
    int gc, lo, len;        // rows [lo, lo + len) of global column gc
    int lu;                 // index among the units of its stock
    int g_ghost;            // G entries [2][NW - 1]: num, then wgt, of rows lo - (NW - 1) .. lo - 1; -1 for a bottom segment
    int up_ghost;           // g_ghost of the unit above in the same column (this unit fills it), or -1
    int g_usp;              // (sum tp, sum tp') of this unit at this step's predation
    int g_sps;              // this unit's part of the spawning stock's recruitment sum s at this step, or -1
};
struct PairRec {
    int pred, prey, suit_kind, suit_ioff, energy_slot, pref_slot, pred_kind, pred_stock, PL;
    int g_suit;             // [set][PL][L]
    int suit_tmap, suit_nset;   // I: the slot set of the suitability each step uses (time-varying selectivity), or -1; number of sets
    int g_part;             // [prey unit][PL] partial suitable biomass
    int cxmask;             // bit j: feeds the prey stock's j-th catch collect vector
    int tsid_off;           // tb: [prey narea] accumulator (fleet) / predator area index (stock predator), or -1
};
struct PredRec { int kind, stock, slot_ioff, narea, g_pts, g_pmc, list_off; };
struct TsRec { int kind, eoff, estr, list_off, list_n; };
struct CompRec {
    int func, weight_par, xbuf, mode, n_dest, n_bins, n_slices, n_areag;
    int csr_ptr_ioff, csr_idx_ioff, csr_coef_roff, cell_dest_ioff, cell_coef_roff, tidx_ioff, cmp_ioff;
    int n_times, param_roff, g_ms, obs_off, obsconst_off;
    int n_groups, grp_off;      // dests that share a total (sumofsquares: the area group's slab, or -- stratified -- a length
                                // group across the slices; multinomial: every slice); tb: [n_dest] group of each dest
    int g_pt, g_pv;             // [n_groups][SHW] the warps' shares of the group totals; [SHW] of the component's value
    int g_cpart;                // partial sums of the split dests
    int task_off;               // tb: [W + 2] offsets into tb, then each warp's tasks (dest, first term, end term, slot, group)
    int split_off;              // tb: count, then (dest, first slot, slots) of every split dest
};
struct XRec { int kind, unit, g_x; };

struct TArgs {
    const int32_t* I; const double* R;
    // every record table and index list of the plan in ONE block of ints (m_* = offsets of its sections), copied to shared
    // memory when the kernel starts if it fits beside the state: the per-step control code of every warp walks these tables,
    // and from global memory every hop is an L2 round trip (the bands and collect vectors stream through L1)
    const int32_t* meta; int meta_n, meta_in_smem, n_s;
    int n_h, use_h;         // H entries: hot per-evaluation tables in shared memory (after the state, before the plan's tables)
    int m_st, m_col, m_unit, m_pp, m_pred, m_ts, m_comp, m_xb, m_agedst, m_tb, m_cidx, m_ccoef;
    int tmask_off;          // tb: [3][NT] per step: components whose time index is set (bit c), that compare (bit c), predation on
    const double* obsprop; const double* obsconst;
    const double* par; long long B; const int32_t* tangent_idx; int K;
    double* nll; double* comp_out; double* grad;
    double* gws; long long slab_doubles; int lanes_per_tile;
    int NC, NS, NPP, NPRED, NTS, NCOMP, NX, NSLOT, NPAR, NT, NSTEPS, NNLL, NBOUND, NCOLS, NU, maxL, n_agedst;
    int ndt, dt_len[4], step_idt[32];
    int n_g;                // G entries per lane; the state entries follow them when SMEM = false
    int s_num, s_wgt;       // S entry bases (NW - 1 zero entries lie below s_num)
    int g_slot, g_eot, g_cn, g_ctot, g_bterm, g_wscr, wscr_n, cf_rows;
    int ax_off, ax_n;       // tb: abundance collect vectors
    int zr_off, zr_n;       // tb: (space 0 = G / 1 = S, first entry, count) triples zeroed once per tile: pads and the band tables
    int agedst_l_off;       // tb: [maxL + 1] first agedst entry of each length group (they are sorted by length group, then cell)
    int uorder_off;         // tb: the units, most expensive first (the work queues' order)
    int inc_off, inc_n;     // tb: global columns that receive maturing fish
    int inc_any_mask;       // bit (cur_step-1): some column receives maturing fish on that step
    int sp_off, sp_n;       // tb: the spawning stocks
    int xf;                 // the model needs the XF instantiation (g3a_otherfood, g3a_spawn, prey_preferences)
    int need;               // features the model uses, in the bits of LEAN (an instantiation with LEAN & need == 0 can run it)
    int spawn_mask;         // bit (cur_step-1): some stock spawns on that step
    int any_ghost;          // some column is split into several units
    int NWc;                // band width the tables are laid out for (the kernel's NW)
    int W;
};

#ifdef __CUDACC__
#define G3T_BAR() __syncthreads()
#define G3T_ANY(p) __any_sync(0xffffffffu, (p))
#define G3T_CTX_DECL
#define G3T_CTX_ARG
// the next ticket of a work queue (a counter in shared memory), the same for every lane of the warp
__device__ __forceinline__ int g3t_grab(int* ctr) {
    int v = 0;
    if ((threadIdx.x & 31) == 0) v = atomicAdd(ctr, 1);
    return __shfl_sync(0xffffffffu, v, 0);
}
#define G3T_GRAB(p) g3t_grab(p)
__device__ __forceinline__ void g3t_prefetch(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#define G3T_TID(warp, lane) ((warp) * 32 + (lane))
#define G3T_NTHREADS(warps) ((warps) * 32)
#else
#define G3T_TID(warp, lane) (warp)              /* one emulated lane at a time: its warps share the copy */
#define G3T_NTHREADS(warps) (warps)
#define G3T_GRAB(p) __atomic_fetch_add((p), 1, __ATOMIC_SEQ_CST)
inline void g3t_prefetch(const void*) {}
struct EmuCtx { void (*bar)(void*); void* arg; };
#define G3T_BAR() ctx.bar(ctx.arg)
#define G3T_ANY(p) (p)
#define G3T_CTX_DECL , const EmuCtx& ctx
#define G3T_CTX_ARG , ctx
#endif

#if defined(__CUDACC__) && defined(G3T_PHASE_CLOCKS)
// profiling build only (tools/gpu_phase_clocks.py): cycles between phase boundaries as warp 0 sees them, summed per phase
__device__ unsigned long long g3t_phase_cycles[16];
#define G3T_PH_DECL long long ph_t0 = clock64(); unsigned long long ph_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define G3T_PH(k) { const long long ph_now = clock64(); ph_acc[k] += (unsigned long long)(ph_now - ph_t0); ph_t0 = ph_now; }
#define G3T_PH_FLUSH() if (warp == 0 && lane == 0) { for (int k = 0; k < 12; k++) atomicAdd(&g3t_phase_cycles[k], ph_acc[k]); }
#else
#define G3T_PH_DECL
#define G3T_PH(k)
#define G3T_PH_FLUSH()
#endif
// element access in the [entry][component][lane] layout (component 0 = value, 1.. = tangents)
__device__ __forceinline__ void ldv(const double* p, double& r) { r = *p; }
__device__ __forceinline__ void stv(double* p, const double& v) { *p = v; }
template <int N> __device__ __forceinline__ void ldv(const double* p, Dual<N>& r) {
    r.v = p[0];
#pragma unroll
    for (int k = 0; k < N; k++) r.d[k] = p[(k + 1) * NL];
}
template <int N> __device__ __forceinline__ void stv(double* p, const Dual<N>& v) {
    p[0] = v.v;
#pragma unroll
    for (int k = 0; k < N; k++) p[(k + 1) * NL] = v.d[k];
}

template <class T> __device__ __forceinline__ T ldT(const double* p) { T r; ldv(p, r); return r; }
template <class T> __device__ __forceinline__ void stT(double* p, const T& v) { stv(p, v); }

// ---- g3a_growmature (R/action_grow.R:340-497), the banded gather of ONE BLOCK of NB destination rows r0 .. r0 + NB - 1 of a
// unit, as straight-line code: source rows r0 - (NW - 1) .. r0 + NB - 1 are loaded once each, every (row, jump) pair is a
// compile-time offset from the block pointers. Pointers address row 0 of the unit (pN, pW: state; pQ: walpha l^wbeta;
// pCm: constant-maturity ratios) resp. band row 0 of the unit (pB, pBr).
//   MODE 0: no maturation now            b = sum_d G[l][d] n[l-d]
//   MODE 1: g3a_mature_continuous        a (maturing) uses the second band (R/action_grow.R:420-437)
//   MODE 2: g3a_mature_constant          a / b split by the ratio of the SOURCE cell (R/action_grow.R:438-457)
//   LOW:    the block reaches below the unit's first row and the unit has ghost rows (pGh: [2][NW - 1]) for them
// Weights: a fish of row s that ends in row l gains wq[l] - wq[s] with wq = walpha l^wbeta (g3a_grow_weightsimple,
// R/action_grow.R:58-71). The difference is formed FIRST, as in the reference: for the fish that stay (s = l) it is exactly
// zero whatever the size of wq against the weight -- regrouping it as wq[l] * sum m + sum m (w[s] - wq[s]) saves two adds per
// pair but loses digits of w when walpha l^wbeta is far above the actual weights (measured: 4e-10 on nll for such vectors).
template <class T, int NW, int NB, int MODE, bool LOW>
__device__ __forceinline__ void grow_gather(const double* __restrict__ pN, const double* __restrict__ pW, const double* __restrict__ pQ,
                                            const double* __restrict__ pB, const double* __restrict__ pBr,
                                            const double* __restrict__ pCm, const double* __restrict__ pGh, const int r0,
                                            const T& mortf, T (&an)[NB], T (&aw)[NB], T (&bn)[NB], T (&bw)[NB]) {
    constexpr int C = (Tan<T>::n + 1) * NL;
    T wq[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) { an[b] = T(0.0); aw[b] = T(0.0); bn[b] = T(0.0); bw[b] = T(0.0); wq[b] = ldT<T>(pQ + (r0 + b) * C); }
    const int rb = r0 - (NW - 1);
    const double* const sN = pN + rb * C;
    const double* const sW = pW + rb * C;
    const double* const sQ = pQ + rb * C;
    const double* const sC = pCm + rb * C;
    const double* const bB = pB + r0 * (NW * C);
    const double* const bR = pBr + r0 * (NW * C);
#pragma unroll
    for (int i = 0; i < NW + NB - 1; i++) {
        const bool ghost = LOW && rb + i < 0;       // (block-uniform)
        const double* const pn = ghost ? pGh + (rb + i + (NW - 1)) * C : sN + i * C;
        const double* const pw = ghost ? pGh + (rb + i + 2 * (NW - 1)) * C : sW + i * C;
        const T ns = ldT<T>(pn) * mortf;
        const T ws = ldT<T>(pw), wqs = ldT<T>(sQ + i * C);
        T nsr(0.0), nss(0.0);
        if (MODE == 2) { const T ratio = ldT<T>(sC + i * C); nsr = ns * ratio; nss = ns * (1.0 - ratio); }
#pragma unroll
        for (int b = 0; b < NB; b++) {
            const int d = NW - 1 + b - i;           // jump that takes source row i of the window to destination row b
            if (d >= 0 && d < NW) {
                const T g = ldT<T>(bB + (b * NW + d) * C);
                if (MODE == 0) {
                    const T m = ns * g;
                    bn[b] = bn[b] + m; bw[b] = bw[b] + m * ((wq[b] - wqs) + ws);
                } else if (MODE == 1) {
                    const T gr = ldT<T>(bR + (b * NW + d) * C);
                    const T ma = ns * gr, m = ns * g;
                    an[b] = an[b] + ma; aw[b] = aw[b] + ma * ((wq[b] - wqs) + ws);
                    bn[b] = bn[b] + m; bw[b] = bw[b] + m * ((wq[b] - wqs) + ws);
                } else {
                    const T ma = nsr * g, m = nss * g;
                    an[b] = an[b] + ma; aw[b] = aw[b] + ma * ((wq[b] - wqs) + ws);
                    bn[b] = bn[b] + m; bw[b] = bw[b] + m * ((wq[b] - wqs) + ws);
                }
            }
        }
    }
}

// T: double or Dual<N>; NW: compiled width of the growth band (maxlengthgroupgrowth + 1 <= NW);
// SMEM: stock state in shared memory (smem = its base) or in the global slab
// XF: the model uses g3a_otherfood, g3a_spawn or prey_preferences (TArgs::xf, set by the planner). The kernels of the models
// that do not are compiled without that code: beside the hot loops it cost 5-9 % (registers, instruction footprint).
// LEAN: features the model does NOT use, compiled out (TArgs::need is what it does use; the dispatch picks the leanest instantiation
// that covers it): bit 0 migration, bit 1 anything but totalfleet predation (stock predators, number- / linear- / effort-fleets),
// bit 2 constant maturity, bit 3 likelihood functions other than sumofsquares and log survey indices. Never-taken branches
// beside the hot loops are not free here (registers, code layout): C2 +6 %, C1 +4 %, C4 / C5 +2 % (variant A/B on a B200).
template <class T, int NW, bool SMEM, int NBV = 0, bool XF = true, int LEAN = 0>
__device__ __forceinline__ void tile_body(const TArgs& A, double* smem, int* ctr, const int warp, const int lane, const int cta,
                                          const int ncta G3T_CTX_DECL) {
    constexpr bool F_MIG = !(LEAN & 1), F_PRED = !(LEAN & 2), F_CMAT = !(LEAN & 4), F_LIKE = !(LEAN & 8);
    constexpr int NTAN = Tan<T>::n;
    constexpr int NB = NBV > 0 ? NBV : (NTAN > 0 ? 1 : NBX);    // rows per block (interleaved avoid_zero / reciprocal chains)
    constexpr unsigned C = (unsigned)(NTAN + 1) * NL;       // doubles per entry
    const int32_t* __restrict__ I = A.I;
    const double* __restrict__ R = A.R;
    // A.W unit warps plus one KEEPER warp (the last): it owns no unit; it keeps the objective (nll, component sums,
    // understocking) in the reference's summation order while the unit warps are already on the next step
    const int W = A.W + 1;
    const bool keeper = warp == A.W;
    double* const gl = A.gws + (size_t)cta * (size_t)A.slab_doubles + lane;
    double* const sm = SMEM ? smem + lane : gl + (size_t)A.n_g * C;
    double* const hs = smem + (SMEM ? (size_t)A.n_s * C : 0) + lane;
#define HP(e) (hs + (unsigned)(e) * C)
    const int32_t* M = A.meta;
    if (A.meta_in_smem) {       // the plan's tables, once per CTA
        int32_t* const ms = reinterpret_cast<int32_t*>(smem + (SMEM ? (size_t)A.n_s * C : 0) + (A.use_h ? (size_t)A.n_h * C : 0));
        for (int i = G3T_TID(warp, lane); i < A.meta_n; i += G3T_NTHREADS(A.W + 1)) ms[i] = A.meta[i];
        G3T_BAR();
        M = ms;
    }
    const StockRec* const Tst = reinterpret_cast<const StockRec*>(M + A.m_st);
    const ColRec* const Tcol = reinterpret_cast<const ColRec*>(M + A.m_col);
    const UnitRec* const Tunit = reinterpret_cast<const UnitRec*>(M + A.m_unit);
    const PairRec* const Tpp = reinterpret_cast<const PairRec*>(M + A.m_pp);
    const PredRec* const Tpred = reinterpret_cast<const PredRec*>(M + A.m_pred);
    const TsRec* const Tts = reinterpret_cast<const TsRec*>(M + A.m_ts);
    const CompRec* const Tcomp = reinterpret_cast<const CompRec*>(M + A.m_comp);
    const XRec* const Txb = reinterpret_cast<const XRec*>(M + A.m_xb);
    const int4* const Tagedst = reinterpret_cast<const int4*>(M + A.m_agedst);
#define GP(e) (gl + (unsigned)(e) * C)
#define SP(e) (sm + (unsigned)(e) * C)
    auto G = [&](int e) { T r; ldv(GP(e), r); return r; };
    auto GS = [&](int e, const T& v) { stv(GP(e), v); };
    auto S = [&](int e) { T r; ldv(SP(e), r); return r; };
    auto SS = [&](int e, const T& v) { stv(SP(e), v); };
#define SLOT(i) G(A.g_slot + (i))

    const int32_t* __restrict__ tb = M + A.m_tb;
    const int NS = A.NS, NSTEPS = A.NSTEPS, NT = A.NT;
    const int ntan_tiles = NTAN > 0 ? (A.K + NTAN - 1) / NTAN : 1;
    const long long total = A.B * ntan_tiles;
    const int LPT = A.lanes_per_tile;
    const long long ntile = (total + LPT - 1) / LPT;
    const double us_power = R[I[G3H_US_ROFF]], us_weight = R[I[G3H_US_ROFF] + 1];
    const bool have_us = I[G3H_US_N] > 0;
    // The units of a phase are a work queue: a counter in shared memory (ctr: one per phase and step parity) hands the next unit
    // of the planner's list (most expensive first) to whichever unit warp is free -- the cost of a unit depends on the data
    // (rows with plenty of fish skip the avoid_zero chain), so no static assignment balances. Every per-unit result is stored
    // per unit and summed in unit order afterwards: which warp processed a unit changes nothing.
    const int NU = A.NU;
#define G3T_UNITS(q, k) for (int k = keeper ? NU : G3T_GRAB(ctr + (q)); k < NU; k = G3T_GRAB(ctr + (q)))

    G3T_PH_DECL
    for (long long tile = cta; tile < ntile; tile += ncta) {
        const long long work0 = tile * LPT + lane;
        const bool live = lane < LPT && work0 < total;
        const long long work = live ? work0 : tile * LPT;       // idle lanes redo the tile's first item without writing
        const long long b = work / ntan_tiles;
        const int tile_t = (int)(work - b * ntan_tiles);
        int seed[NTAN > 0 ? NTAN : 1];
#pragma unroll
        for (int k = 0; k < (NTAN > 0 ? NTAN : 1); k++) {
            const int g = tile_t * NTAN + k;
            seed[k] = (NTAN > 0 && g < A.K) ? A.tangent_idx[g] : -1;
        }
        const double* __restrict__ prow = A.par + b * A.NPAR;

        // ---- g3a_time (R/action_time.R:75-91)
        const double retro = I[G3H_PAR_RETRO] >= 0 ? prow[I[G3H_PAR_RETRO]] : 0.0;
        const double proj = I[G3H_PAR_PROJECT] >= 0 ? prow[I[G3H_PAR_PROJECT]] : 0.0;
        const int total_steps = NSTEPS * (I[G3H_END_YEAR] - (int)floor(retro) - I[G3H_START_YEAR] + (int)floor(proj)) + NSTEPS - 1;
        const bool bad_time = !(retro >= 0.0) || !(proj >= 0.0) || total_steps + 1 > NT;

        G3T_BAR();      // the previous tile is fully retired before the slab and the state are reused

        // ================= scalar formula slots, once per parameter vector; pads and band tables start from zero
        for (int i = warp; i < A.NSLOT; i += W) GS(A.g_slot + i, eval_slot_g<T>(I, R, prow, seed, i));
        if (keeper) for (int i = 0; i < 8; i++) ctr[i] = 0;
        if (A.use_h) for (int i = warp; i < A.n_h; i += W) stT<T>(HP(i), T(0.0));       // (pads and the band tables start from zero)
        for (int z = 0; z < A.zr_n; z++) {
            const int space = tb[A.zr_off + 3 * z], e0 = tb[A.zr_off + 3 * z + 1], n = tb[A.zr_off + 3 * z + 2];
            for (int i = warp; i < n; i += W) { if (space) SS(e0 + i, T(0.0)); else GS(e0 + i, T(0.0)); }
        }
        G3T_BAR();

        // ================= per-evaluation tables. Work units are dealt round-robin to the warps.
        int ucnt = 0;
        auto mine = [&]() { const bool m = ucnt == warp; ucnt = ucnt + 1 == W ? 0 : ucnt + 1; return m; };

        for (int q = 0; q < A.NPP; q++) {           // suitability (R/suitability.R:5-94): [predator length][prey length]
            const PairRec& Q = Tpp[q];
            const StockRec& Ps = Tst[Q.prey];
            const bool spred = Q.pred_kind == G3PRED_STOCK;
            const double* pml = R + Ps.midlen_roff;
            for (int sset = 0; sset < Q.suit_nset; sset++)        // (one set unless the selectivity varies with time)
            for (int pl = 0; pl < Q.PL; pl++) {
                if (!mine()) continue;
                const int32_t* ss = I + Q.suit_ioff + 6 * sset;
                const int gsu = Q.g_suit + sset * Q.PL * Ps.L;
                T sp[5];
#pragma unroll
                for (int k = 0; k < 5; k++) sp[k] = ss[k] >= 0 ? SLOT(ss[k]) : T(0.0);
                const double plen = spred && Q.suit_kind != G3SUIT_ANDERSENFLEET ? R[Tst[Q.pred_stock].midlen_roff + pl] : pml[Ps.L - 1];
                if (!XF || Q.pref_slot < 0)
                    for (int l = 0; l < Ps.L; l++) GS(gsu + pl * Ps.L + l, suitability_of<T>(Q.suit_kind, sp, pml[l], plen));
                else {
                    // prey_preference p (R/action_predate.R:220-225): (S E B)^p = S^p E^p B^p; the table holds S^p E^(p-1) -- x E x B^p
                    // in the predator's total, x B^p in the consumption (where one E cancels against cons' 1 / E)
                    const T p = SLOT(Q.pref_slot), ef = xpow(SLOT(Q.energy_slot), p - 1.0);
                    for (int l = 0; l < Ps.L; l++) GS(gsu + pl * Ps.L + l, xpow(suitability_of<T>(Q.suit_kind, sp, pml[l], plen), p) * ef);
                }
            }
        }
        // stock predators: maximum consumption per predator length without the step length,
        // m0 exp(m1 T - m2 T^3) l^m3 (R/action_predate.R:190-205)
        for (int pi = 0; pi < A.NPRED; pi++) {
            const PredRec& P = Tpred[pi];
            if (P.kind != G3PRED_STOCK) continue;
            const StockRec& Pp = Tst[P.stock];
            for (int pl = 0; pl < Pp.L; pl++) {
                if (!mine()) continue;
                const int32_t* ps = I + P.slot_ioff;
                const T m0 = SLOT(ps[0]), m1 = SLOT(ps[1]), m2 = SLOT(ps[2]), m3 = SLOT(ps[3]), temp = SLOT(ps[5]);
                const T tf = m0 * xexp(m1 * temp - m2 * temp * temp * temp);
                GS(P.g_pmc + pl, tf * xpow(T(R[Pp.midlen_roff + pl]), m3));
            }
        }
        for (int s = 0; s < NS; s++) {
            const StockRec& St = Tst[s];
            const int L = St.L;
            const double* midlen = R + St.midlen_roff;
            // migration matrices per step of the year: proportion moving src -> dest (R/action_migrate.R:2-15,49-56)
            if (St.mig_ioff >= 0) {
                const int NR = St.narea;
                for (int sy = 0; sy < NSTEPS; sy++) {
                    if (!mine()) continue;
                    const int32_t* mig = I + St.mig_ioff + sy * NR * NR;
                    for (int src = 0; src < NR; src++) {
                        T tot(0.0);
                        for (int d = 0; d < NR; d++) {
                            T v = (d != src && mig[d * NR + src] >= 0) ? SLOT(mig[d * NR + src]) : T(0.0);
                            v = v * v;
                            tot = tot + v;
                            GS(St.g_mig + (sy * NR + d) * NR + src, v / 1.0);
                        }
                        GS(St.g_mig + (sy * NR + src) * NR + src, (1.0 - tot) / 1.0);
                    }
                }
            }
            // g3a_spawn (R/action_spawn.R:86-233): proportion spawning and spawning mortality by length, fecundity weights
            // midlen^p1 age^p2; per output stock the offspring's share by length group -- exp(-((midlen - mean) / sd)^2 / 2) over
            // avoid_zero(its sum over the stock's areas and lengths) -- and their weight alpha midlen^beta
            if (XF && St.sp_ioff >= 0) {
                const int32_t* Sp = I + St.sp_ioff;
                for (int which = 0; which < 2; which++) {
                    const int kind = Sp[which ? G3SP_MORT_KIND : G3SP_PROP_KIND];
                    if (kind < 0 || !mine()) continue;
                    const int32_t* ss = Sp + (which ? G3SP_MORT_SLOTS : G3SP_PROP_SLOTS);
                    T sp[5];
#pragma unroll
                    for (int k = 0; k < 5; k++) sp[k] = ss[k] >= 0 ? SLOT(ss[k]) : T(0.0);
                    for (int l = 0; l < L; l++) {
                        const T v = suitability_of<T>(kind, sp, midlen[l], midlen[l]);
                        if (which) GS(St.g_smort + l, 1.0 - xexp(-v)); else GS(St.g_sprop + l, v);
                    }
                }
                if (St.g_sfec >= 0 && mine()) {
                    const T p1 = SLOT(Sp[G3SP_RECR_SLOTS + 1]), p2 = SLOT(Sp[G3SP_RECR_SLOTS + 2]);
                    for (int a = 0; a < St.nage; a++) {
                        const T ap = xpow(T((double)(St.minage + a)), p2);
                        for (int l = 0; l < L; l++) GS(St.g_sfec + a * L + l, xpow(T(midlen[l]), p1) * ap);
                    }
                }
                for (int o = 0; o < Sp[G3SP_N_OUT]; o++) {
                    if (!mine()) continue;
                    const int32_t* Or = I + Sp[G3SP_OUT_OFF] + 6 * o;
                    const StockRec& So = Tst[Or[0]];
                    const double* oml = R + So.midlen_roff;
                    const int gn = tb[St.sp_out_off + 4 * o + 1], gw = tb[St.sp_out_off + 4 * o + 2];
                    const T mean = SLOT(Or[1]), isd = 1.0 / SLOT(Or[2]), wa = SLOT(Or[3]), wb = SLOT(Or[4]);
                    T tot(0.0);
                    for (int l = 0; l < So.L; l++) { const T z = (oml[l] - mean) * isd; GS(gn + l, xexp(-(z * z) * 0.5)); }
                    for (int r = 0; r < So.narea; r++) for (int l = 0; l < So.L; l++) tot = tot + G(gn + l);
                    const T den = avoid_zero(tot);
                    for (int l = 0; l < So.L; l++) { GS(gn + l, G(gn + l) / den); GS(gw + l, wa * xpow(T(oml[l]), wb)); }
                }
            }
            // renewal length distribution, normalised, and weight (R/action_renewal.R:56-80)
            for (int k = 0; k < St.n_renew; k++) {
                if (!mine()) continue;
                const int32_t* Rn = I + St.renew_ioff + k * G3R_LEN;
                const T mean = SLOT(Rn[G3R_MEAN]), sd = avoid_zero(SLOT(Rn[G3R_SD]));
                const T wa = SLOT(Rn[G3R_WALPHA]), wb = SLOT(Rn[G3R_WBETA]);
                T tot(0.0);
                for (int l = 0; l < L; l++) { const T d = dnorm(midlen[l], mean, sd); GS(St.g_rd + k * L + l, d); tot = tot + d; }
                const T den = avoid_zero(tot);
                for (int l = 0; l < L; l++) {
                    GS(St.g_rd + k * L + l, G(St.g_rd + k * L + l) / den * 10000.0);
                    GS(St.g_rw + k * L + l, wa * xpow(midlen[l], wb));
                }
            }
            const int n2 = St.grow_n;
            if (n2 < 0) continue;
            // where these tables live: the H entries (shared memory) if the planner found room for them, else the slab
            const bool hwq = A.use_h && St.h_wq >= 0;
            double* const bandp = A.use_h && St.h_band >= 0 ? HP(St.h_band) : GP(St.g_band);
            double* const bandrp = A.use_h && St.h_bandr >= 0 ? HP(St.h_bandr) : GP(St.g_bandr);
            const int ncset = St.band_percol ? St.ncol : 1;
            for (int gset = 0; gset < (XF ? St.grow_nset : 1); gset++) {       // (one set unless the growth parameters vary with time)
            const int32_t* gs = I + St.grow_ioff + 5 * gset;
            for (int l = 0; l < L; l++)         // walpha l^wbeta (g3a_grow_weightsimple, R/action_grow.R:58-71)
                if (mine()) stT<T>(hwq ? HP(St.h_wq + l) : GP(St.g_wq + gset * St.wq_stride + l), SLOT(gs[3]) * xpow(midlen[l], SLOT(gs[4])));
            // ---- growth bands: P(l -> l + x), x = 0..n, of each source row l (growth_bbinom in product form,
            // R/action_grow.R:155-213), stored destination-oriented: entry [l + x][x]; everything that would leave the
            // length axis is summed into the plus group's row L - 1 (R/action_grow.R:233-271). Continuous maturity splits
            // each entry into the staying and the maturing part (R/action_mature.R:1-42, R/action_grow.R:420-437); with an
            // age term there is a set per column.
            const bool cont = St.mat_kind == G3MAT_CONTINUOUS && St.has_out;
            const double pdl = R[St.plusdl_roff];
            for (int set = 0; set < ncset; set++)
                for (int idt = 0; idt < A.ndt; idt++)
                    for (int l = 0; l < L; l++) {
                        if (!mine()) continue;
                        const double dt = A.dt_len[idt] / 12.0;
                        const T linf = SLOT(gs[0]), kk = 1.0 - xexp(-(SLOT(gs[1])) * dt), beta = avoid_zero(SLOT(gs[2]));
                        // g3a_grow_lengthvbsimple (R/action_grow.R:54) inside the bbinom wrapper (R/action_grow.R:220)
                        const T dl = avoid_zero((linf - midlen[l]) * kk);
                        const T delta = avoid_zero(dl / pdl);
                        const T alpha = (beta * delta) / ((double)n2 - delta);
                        T g = T(1.0);
                        for (int j = 0; j < n2; j++) g = g * ((beta + (double)j) / xabs(alpha + beta + (double)j));
                        T m0(0.0), malpha(0.0), mbeta(0.0);
                        if (cont) {         // g3a_mature_constant as M0 (R/action_mature.R:44-74)
                            const int32_t* ms = I + St.mat_ioff + 4 * gset;      // (the maturity parameters vary with the growth sets)
                            T inner(0.0);
                            if (ms[0] >= 0) { malpha = SLOT(ms[0]); inner = inner - malpha * (midlen[l] - SLOT(ms[1])); }
                            if (ms[2] >= 0) { mbeta = SLOT(ms[2]); inner = inner - mbeta * ((double)(St.minage + set % St.nage) - SLOT(ms[3])); }
                            m0 = 1.0 / (1.0 + xexp(inner));
                        }
                        const int base = ((gset * ncset + set) * A.ndt + idt) * St.bsz;
                        const int dtail = L - 1 - l;            // jumps of at least dtail end in the plus group
                        T ta(0.0), tr(0.0);
                        T gx[NW], rx[NW];
#pragma unroll
                        for (int x = 0; x < NW; x++) {
                            gx[x] = T(0.0); rx[x] = T(0.0);
                            if (x <= n2) {
                                if (cont) {
                                    const T ratio = dif_pminmax_c(m0 * (malpha * (pdl * x) + mbeta * dt), 0.0, 1.0, 1e5);
                                    rx[x] = g * ratio; gx[x] = g * (1.0 - ratio);
                                } else gx[x] = g;
                                if (x < dtail) {
                                    stT<T>(bandp + (unsigned)(base + (l + x) * NW + x) * C, gx[x]);
                                    if (cont) stT<T>(bandrp + (unsigned)(base + (l + x) * NW + x) * C, rx[x]);
                                }
                                if (x < n2)
                                    g = g * ((double)(n2 - x) / (double)(x + 1)) * (xabs(alpha + (double)x) / (beta + (double)(n2 - x - 1)));
                            }
                        }
                        if (dtail <= n2) {
#pragma unroll
                            for (int x = 0; x < NW; x++)
                                if (x >= dtail && x <= n2) { ta = ta + gx[x]; tr = tr + rx[x]; }
                            stT<T>(bandp + (unsigned)(base + (L - 1) * NW + dtail) * C, ta);
                            if (cont) stT<T>(bandrp + (unsigned)(base + (L - 1) * NW + dtail) * C, tr);
                        }
                    }
            }   // gset
        }
        // ---- per column: mortality factors, initial conditions, maturation remainder, constant-maturity ratios
        for (int gc = warp; gc < A.NCOLS; gc += W) {
            const ColRec& Cl = Tcol[gc];
            const StockRec& St = Tst[Cl.s];
            const int L = St.L, col = Cl.col;
            const double* midlen = R + St.midlen_roff;
            if (St.mort_ioff >= 0)      // natural mortality factor per distinct step length (R/action_naturalmortality.R:26)
                for (int ms = 0; ms < St.mort_nset; ms++)       // (one set unless M varies with time)
                    for (int idt = 0; idt < A.ndt; idt++)
                        GS(St.g_mort + (ms * A.ndt + idt) * St.ncol + col, xexp(-SLOT(I[St.mort_ioff + ms * St.ncol + col]) * (A.dt_len[idt] / 12.0)));
            // g3a_initialconditions (R/action_renewal.R:83-163)
            if (St.init_ioff < 0) {
                for (int l = 0; l < L; l++) { SS(A.s_num + Cl.cell0 + l, T(0.0)); SS(A.s_wgt + Cl.cell0 + l, T(1.0)); }
            } else {
                const int32_t* sl = I + St.init_ioff + col * 5;
                const T mean = SLOT(sl[0]), sd = avoid_zero(SLOT(sl[1])), factor = SLOT(sl[2]), wa = SLOT(sl[3]), wb = SLOT(sl[4]);
                T tot(0.0);
                for (int l = 0; l < L; l++) { const T d = dnorm(midlen[l], mean, sd); SS(A.s_num + Cl.cell0 + l, d); tot = tot + d; }
                const T den = avoid_zero(tot);                           // normalize_vec, R/aab_env.R:37-41
                for (int l = 0; l < L; l++) {
                    SS(A.s_num + Cl.cell0 + l, S(A.s_num + Cl.cell0 + l) / den * 10000.0 * factor);
                    SS(A.s_wgt + Cl.cell0 + l, wa * xpow(midlen[l], wb));
                }
            }
            if (St.grow_n >= 0 && St.has_out) {
                // share of the maturing fish of this column that no output stock claims (R/action_mature.R:126-131)
                T remf(1.0);
                for (int k = Cl.rem_off; tb[k] >= 0; k++) remf = remf - SLOT(tb[k]);
                GS(Cl.g_remf, remf);
                if (St.mat_kind == G3MAT_CONSTANT) {        // ratio per (column, length) (R/action_mature.R:44-74)
                    const int32_t* msl = I + St.mat_ioff;
                    for (int l = 0; l < L; l++) {
                        T inner(0.0);
                        if (msl[0] >= 0) inner = inner - SLOT(msl[0]) * (midlen[l] - SLOT(msl[1]));
                        if (msl[2] >= 0) inner = inner - SLOT(msl[2]) * ((double)(St.minage + Cl.aidx) - SLOT(msl[3]));
                        GS(St.g_cmat + col * L + l, 1.0 / (1.0 + xexp(inner)));
                    }
                }
            }
        }
        // ---- g3l_bounds_penalty terms (R/likelihood_bounds.R:13-16); the keeper warp adds them in run order below
        {
            const double bs = A.NBOUND > 0 ? R[I[G3H_BOUND_WS_ROFF] + 1] : 0.0;
            for (int i = warp; i < A.NBOUND; i += W) {
                const T p = mkpar_g<T>(prow, seed, I[I[G3H_BOUND_OFF] + i]);
                const double lo = R[I[G3H_BOUND_ROFF] + 2 * i], up = R[I[G3H_BOUND_ROFF] + 2 * i + 1];
                const T a = lsa_c(bs * (p - up) / (up - lo), 0.0) + lsa_c(bs * (lo - p) / (up - lo), 0.0);
                GS(A.g_bterm + i, a * a);
            }
        }
        for (int c = 0; c < A.NCOMP; c++) {         // zero the likelihood slabs
            const CompRec& Cm = Tcomp[c];
            const bool si = Cm.func == G3F_SI_LOG || Cm.func == G3F_SI_LINEAR;
            const int n = Cm.n_dest * (si ? Cm.n_times : 1);
            for (int i = warp; i < n; i += W) GS(Cm.g_ms + i, T(0.0));
        }
        G3T_BAR();

        // which components this tile evaluates: <name>_weight > 0 on some lane (R/likelihood_distribution.R:352)
        unsigned cw_any = 0, cw_lane = 0;
        for (int c = 0; c < A.NCOMP; c++) {
            const bool on = prow[Tcomp[c].weight_par] > 0.0;
            if (on) cw_lane |= 1u << c;
            if (G3T_ANY(on)) cw_any |= 1u << c;
        }

        T nll(0.0);         // the keeper warp holds the objective and the component sums
        if (keeper) {
            for (int i = 0; i < A.NNLL; i++) GS(A.g_ctot + i, T(0.0));
            if (A.NBOUND > 0) {
                T pen(0.0);
                for (int i = 0; i < A.NBOUND; i++) pen = pen + G(A.g_bterm + i);
                nll = nll + R[I[G3H_BOUND_WS_ROFF]] * pen;
                GS(A.g_ctot + A.NNLL - 1, pen);
            }
        }

        G3T_PH(0)       // tile set-up: slots, tables, initial conditions
        for (int t = 0; t < NT; t++) {
            const bool lane_on = !bad_time && t <= total_steps;
            if (!G3T_ANY(lane_on)) break;       // (every warp holds the same 32 evaluations: the decision is block-uniform)
            const int step = t % NSTEPS, yidx = t / NSTEPS;
            int* const qn = ctr + 4 * (t & 1);          // this step's queues; the keeper rewinds the next step's (idle since the step before)
            if (keeper) for (int i = 0; i < 4; i++) ctr[4 * ((t + 1) & 1) + i] = 0;
            const int idt = A.step_idt[step];
            const double dt = A.dt_len[idt] / 12.0;
            const bool final_step = step == NSTEPS - 1;
            const bool pred_now = tb[A.tmask_off + 2 * NT + t] != 0;
            // components collecting at this step (bit c) and the CATCH collect vectors they read (bit x; abundance components read
            // the state itself when they collect)
            const unsigned cact = (unsigned)tb[A.tmask_off + t] & cw_any;
            const unsigned cmp_now = (unsigned)tb[A.tmask_off + NT + t];      // components that compare at this step (bit c)
            unsigned xact = 0;
            for (int c = 0; c < A.NCOMP; c++)
                if ((cact >> c) & 1) {
                    const CompRec& Cm = Tcomp[c];
                    if (Txb[Cm.xbuf].kind == 0) xact |= 1u << Cm.xbuf;
                }
            const bool any_catch = xact != 0;

            // ================= g3a_otherfood (R/action_renewal.R:276-308): number and mean weight assigned from formulas
            bool of_any = false;
            for (int s = 0; XF && s < NS; s++) {
                const StockRec& St = Tst[s];
                if (St.of_ioff < 0) continue;
                of_any = true;
                const int per = St.nage * St.L;
                for (int i = warp; i < St.narea * per; i += W) {
                    const int32_t* sl = I + St.of_ioff + ((size_t)t * St.narea + i / per) * 2;
                    SS(A.s_num + St.c0 + i, SLOT(sl[0])); SS(A.s_wgt + St.c0 + i, SLOT(sl[1]));
                }
            }
            // ================= g3a_migrate (R/action_migrate.R:17-82): every (age, length) mixes its areas
            {
                bool mig_any = of_any;      // (one barrier for both)
                ucnt = 0;
                for (int s = 0; F_MIG && s < NS; s++) {
                    const StockRec& St = Tst[s];
                    if (!((St.mig_steps >> step) & 1)) continue;
                    mig_any = true;
                    const int L = St.L, nage = St.nage, NR = St.narea, c0 = St.c0;
                    const int mo = St.g_mig + step * NR * NR;
                    // work items: NB length groups of one age, dealt to the warps. Destination by destination; the sources are
                    // read before any destination is written (the warp's scratch holds them); the ratio_add_pop chains
                    // (stock_combine_subpop: one avoid_zero + reciprocal per source area) of the NB rows run interleaved
                    const int ws = A.g_wscr + warp * A.wscr_n;
                    for (int a = 0; a < nage; a++)
                        for (int l0 = 0; l0 < L; l0 += NB) {
                            if (!mine()) continue;
                            const int nb = min(NB, L - l0);
                            for (int r = 0; r < NR; r++)
#pragma unroll
                                for (int b = 0; b < NB; b++) {
                                    const int i = a * L + min(l0 + b, L - 1);
                                    GS(ws + (2 * r) * NB + b, S(A.s_num + c0 + r * nage * L + i));
                                    GS(ws + (2 * r + 1) * NB + b, S(A.s_wgt + c0 + r * nage * L + i));
                                }
                            for (int d = 0; d < NR; d++) {
                                T pn[NB], pw[NB];
#pragma unroll
                                for (int b = 0; b < NB; b++) { pn[b] = T(0.0); pw[b] = T(0.0); }
                                for (int r = 0; r < NR; r++) {
                                    const T prop = G(mo + d * NR + r);
                                    T nu[NB], de[NB], moved[NB];
#pragma unroll
                                    for (int b = 0; b < NB; b++) {
                                        moved[b] = G(ws + (2 * r) * NB + b) * prop;
                                        nu[b] = pw[b] * pn[b] + G(ws + (2 * r + 1) * NB + b) * moved[b];      // ratio_add_pop, R/aab_env.R:133-153
                                        de[b] = pn[b] + moved[b];
                                    }
                                    div_az_n(nu, de);
#pragma unroll
                                    for (int b = 0; b < NB; b++) { pw[b] = nu[b]; pn[b] = pn[b] + moved[b]; }
                                }
#pragma unroll
                                for (int b = 0; b < NB; b++)
                                    if (b < nb) { const int i = a * L + l0 + b; SS(A.s_num + c0 + d * nage * L + i, pn[b]); SS(A.s_wgt + c0 + d * nage * L + i, pw[b]); }
                            }
                        }
                }
                if (mig_any) G3T_BAR();
                G3T_PH(1)
            }

            // ================= g3a_predate step 1: suitable biomass per (predator, area[, predator length]) (R/action_predate.R:314-333)
            if (pred_now) {
                G3T_UNITS(qn - ctr + 0, uk) {              // partial sums per unit
                    const UnitRec& U = Tunit[tb[A.uorder_off + uk]];
                    const ColRec& Cl = Tcol[U.gc];
                    const StockRec& St = Tst[Cl.s];
                    const int L = St.L, c_lo = Cl.cell0 + U.lo;
                    if (!SMEM && St.pp_n > 0)       // state in the slab: the unit's rows on their way (this loop and the predation pass read them)
                        for (int r = 0; r < U.len; r++) { g3t_prefetch(SP(A.s_num + c_lo + r)); g3t_prefetch(SP(A.s_wgt + c_lo + r)); }
                    for (int j = 0; j < St.pp_n; j++) {
                        const int q = tb[St.pp_off + j];
                        const PairRec& Q = Tpp[q];
                        if (F_PRED && (Q.pred_kind == G3PRED_LINEARFLEET || Q.pred_kind == G3PRED_EFFORTFLEET)) continue;     // no total needed
                        if (tb[Q.tsid_off + Cl.ridx] < 0) continue;
                        const bool spred = F_PRED && Q.pred_kind == G3PRED_STOCK;
                        const bool by_number = F_PRED && Q.pred_kind == G3PRED_NUMBERFLEET;       // suit = S * num (R/action_predate.R:7-11)
                        const int gsu = Q.g_suit + (XF && Q.suit_tmap >= 0 ? I[Q.suit_tmap + t] * Q.PL * L : 0);      // this step's table
                        // four predator length groups at a time: the unit's biomass is read once per four, their loads in flight together
                        for (int pl0 = 0; pl0 < Q.PL; pl0 += 4) {
                            T cs[4];
#pragma unroll
                            for (int u = 0; u < 4; u++) cs[u] = T(0.0);
                            for (int r = 0; r < U.len; r++) {
                                const T n = S(A.s_num + c_lo + r);
                                T bm = by_number ? n : n * S(A.s_wgt + c_lo + r);
                                if (XF && Q.pref_slot >= 0) bm = xpow(bm, SLOT(Q.pref_slot));
#pragma unroll
                                for (int u = 0; u < 4; u++)
                                    cs[u] = cs[u] + G(gsu + min(pl0 + u, Q.PL - 1) * L + U.lo + r) * bm;
                            }
                            const T en = spred ? SLOT(Q.energy_slot) : T(1.0);
#pragma unroll
                            for (int u = 0; u < 4; u++)
                                if (pl0 + u < Q.PL) GS(Q.g_part + U.lu * Q.PL + pl0 + u, spred ? cs[u] * en : cs[u]);
                        }
                    }
                }
                G3T_BAR();
                G3T_PH(2)
                ucnt = 0;
                for (int k = 0; k < A.NTS; k++) {           // fleets: landings over total suitable biomass, cons = E * (suit / totalsuit)
                    if (!mine()) continue;
                    const TsRec& Ts = Tts[k];
                    const double E = R[Ts.eoff + (size_t)t * Ts.estr];
                    if (F_PRED && (Ts.kind == G3PRED_LINEARFLEET || Ts.kind == G3PRED_EFFORTFLEET)) {
                        GS(A.g_eot + k, T(E * dt));         // harvest rate x step length (R/action_predate.R:13-18,117-126)
                    } else {
                        T tot(0.0);
                        for (int i = 0; i < Ts.list_n; i++) tot = tot + G(tb[Ts.list_off + i]);
                        GS(A.g_eot + k, E / tot);
                    }
                }
                // stock predators: predN * max_consumption * feeding_level / total_predsuit per (area, predator length)
                // (R/action_predate.R:207-238; energycontent cancels between suit and cons)
                for (int pi = 0; F_PRED && pi < A.NPRED; pi++) {
                    const PredRec& P = Tpred[pi];
                    if (P.kind != G3PRED_STOCK) continue;
                    const StockRec& Pp = Tst[P.stock];
                    const int PL = Pp.L, PA = Pp.nage;
                    for (int pr = 0; pr < Pp.narea; pr++) {
                        const int lo = tb[P.list_off + 2 * pr], ln = tb[P.list_off + 2 * pr + 1];
                        for (int pl = 0; pl < PL; pl++) {
                            if (!mine()) continue;
                            T ts(0.0);
                            for (int i = 0; i < ln; i++) ts = ts + G(tb[lo + i] + pl);
                            T predn(0.0);
                            for (int pa = 0; pa < PA; pa++) predn = predn + S(A.s_num + Pp.c0 + (pr * PA + pa) * PL + pl);
                            const T hf = SLOT(I[P.slot_ioff + 4]);
                            const T feeding = ts / (hf * dt + ts);
                            GS(P.g_pts + pr * PL + pl, predn * (G(P.g_pmc + pl) * dt) * feeding / ts);
                        }
                    }
                }
                G3T_BAR();
            }

            // ---- what a block of rows does once its growth is through: hand-over, renewal, store, collect vectors.
            // Rows row0 .. row0 + NB - 1 of a column, the first nb of them real (the others are computed and dropped).
            auto finish_block = [&](const ColRec& Cl, const StockRec& St, const int row0, const int nb, T (&num)[NB], T (&wgt)[NB],
                                    const bool inc_here, const int ren_mask) {
                const int L = St.L;
                // g3a_step_transition (R/action_mature.R:78-134): maturing fish arriving from their source column
                if (inc_here) {
                    const T ratio = SLOT(Cl.inc_slot);
                    T nu[NB], de[NB], mv[NB];
#pragma unroll
                    for (int b = 0; b < NB; b++) {
                        const int l = min(row0 + b, L - 1);
                        mv[b] = S(Cl.inc_tn + l) * ratio;
                        nu[b] = wgt[b] * num[b] + S(Cl.inc_tw + l) * mv[b];     // ratio_add_pop, R/aab_env.R:133-153
                        de[b] = num[b] + mv[b];
                    }
                    div_az_n(nu, de);
#pragma unroll
                    for (int b = 0; b < NB; b++) { wgt[b] = nu[b]; num[b] = num[b] + mv[b]; }
                }
                // g3a_renewal (R/action_renewal.R:166-188)
                if (ren_mask)
                    for (int r = 0; r < St.n_renew; r++)
                        if (((ren_mask >> r) & 1) && Cl.aidx == tb[St.ren_off + 3 * r + 1]) {
                            const int fs = I[tb[St.ren_off + 3 * r + 2] + yidx * St.narea + Cl.ridx];
                            if (fs < 0) continue;
                            const T factor = SLOT(fs);
                            T nu[NB], de[NB], rn[NB];
#pragma unroll
                            for (int b = 0; b < NB; b++) {
                                const int l = min(row0 + b, L - 1);
                                rn[b] = G(St.g_rd + r * L + l) * factor;
                                nu[b] = wgt[b] * num[b] + G(St.g_rw + r * L + l) * rn[b];
                                de[b] = num[b] + rn[b];
                            }
                            div_az_n(nu, de);
#pragma unroll
                            for (int b = 0; b < NB; b++) { wgt[b] = nu[b]; num[b] = num[b] + rn[b]; }
                        }
                const int cell = Cl.cell0 + row0;
#pragma unroll
                for (int b = 0; b < NB; b++)
                    if (b < nb) { SS(A.s_num + cell + b, num[b]); SS(A.s_wgt + cell + b, wgt[b]); }
                // collect vectors for g3l_distribution (R/likelihood_distribution.R:275-287)
                if (xact) {
                    bool any_cn = false;        // catch in numbers: cons / avoid_zero(wgt)
                    for (int j = 0; j < St.cx_n; j++) { const int x = tb[St.cx_off + j]; any_cn = any_cn || (((xact >> x) & 1) && Txb[x].unit == 0); }
                    if (any_cn) {
                        // (the collect values were written by the predation pass: where the slabs do not fit the L2 every one of
                        // these loads is a DRAM round trip -- get them on their way before the reciprocal chain, and read a block's
                        // rows together instead of load - store - load)
                        for (int j = 0; j < St.cx_n; j++) {
                            const int x = tb[St.cx_off + j];
                            if (!(((xact >> x) & 1) && Txb[x].unit == 0)) continue;
#pragma unroll
                            for (int b = 0; b < NB; b++) if (b < nb) g3t_prefetch(GP(Txb[x].g_x + cell + b));
                        }
                        T iw[NB], de[NB];
#pragma unroll
                        for (int b = 0; b < NB; b++) { iw[b] = T(1.0); de[b] = wgt[b]; }
                        div_az_n(iw, de);
                        for (int j = 0; j < St.cx_n; j++) {
                            const int x = tb[St.cx_off + j];
                            if (!(((xact >> x) & 1) && Txb[x].unit == 0)) continue;
                            const int gx = Txb[x].g_x + cell;
                            T xv[NB];
#pragma unroll
                            for (int b = 0; b < NB; b++) xv[b] = G(gx + min(b, nb - 1));
#pragma unroll
                            for (int b = 0; b < NB; b++) if (b < nb) GS(gx + b, xv[b] * iw[b]);
                        }
                    }
                }
            };
            auto renewals_now = [&](const StockRec& St) {
                int ren_mask = 0;       // renewal records that run at this step
                for (int k = 0; k < St.n_renew; k++) {
                    const int rs = tb[St.ren_off + 3 * k];
                    if (rs == 0 || rs == step + 1) ren_mask |= 1 << k;
                }
                return ren_mask;
            };

            // ================= own units, pass 1: g3a_predate (R/action_predate.R:335-406) in blocks of NB rows, ascending
            // (the order of the sums); then the unit's top rows go to the ghost rows of the unit above
            G3T_UNITS(qn - ctr + 1, uk) {
                const UnitRec& U = Tunit[tb[A.uorder_off + uk]];
                const ColRec& Cl = Tcol[U.gc];
                const StockRec& St = Tst[Cl.s];
                const int L = St.L, nq = St.pp_n, c_lo = Cl.cell0 + U.lo;
                const bool prey_now = pred_now && nq > 0;
                const bool catch_now = any_catch && nq > 0;
                double* const pNum = SP(A.s_num + c_lo);
                double* const pWgt = SP(A.s_wgt + c_lo);
                if (prey_now) {
                    // consumption per unit of biomass by pair k at rows r0 .. r0 + NB - 1 of the unit (the plan's records are in
                    // shared memory; the suitability rows and the fleets' landings / total suitable biomass are table loads)
                    auto cfac = [&](const PairRec& Q, const int ts, const int r0, T (&cf)[NB]) {
                        const int gsu = Q.g_suit + (XF && Q.suit_tmap >= 0 ? I[Q.suit_tmap + t] * Q.PL * L : 0);      // this step's table
                        if (!F_PRED || Q.pred_kind != G3PRED_STOCK) {
                            T e = G(A.g_eot + ts);
                            if (F_PRED && Q.pred_kind == G3PRED_EFFORTFLEET) e = e * SLOT(Q.energy_slot);      // catchability_fs of this prey
#pragma unroll
                            for (int b = 0; b < NB; b++) cf[b] = G(gsu + U.lo + min(r0 + b, U.len - 1)) * e;
                        } else {        // sum over the predator's length groups of suitability x consumption factor
                            const int gp = Tpred[Q.pred].g_pts + ts * Q.PL;
#pragma unroll
                            for (int b = 0; b < NB; b++) cf[b] = T(0.0);
                            for (int pl = 0; pl < Q.PL; pl++) {
                                const T f = G(gp + pl);
#pragma unroll
                                for (int b = 0; b < NB; b++) cf[b] = cf[b] + G(gsu + pl * L + U.lo + min(r0 + b, U.len - 1)) * f;
                            }
                        }
                    };
                    T o1(0.0), o2(0.0);
                    for (int r0 = 0; r0 < U.len; r0 += NB) {
                        const int nb = min(NB, U.len - r0);
                        T num[NB], B0[NB], tp[NB], cr[NB], de[NB], cf[NB];
#pragma unroll
                        for (int b = 0; b < NB; b++) {
                            num[b] = ldT<T>(pNum + (unsigned)(r0 + b) * C);
                            B0[b] = num[b] * ldT<T>(pWgt + (unsigned)(r0 + b) * C);
                            tp[b] = T(0.0);
                        }
                        for (int k = 0; k < nq; k++) {
                            const PairRec& Q = Tpp[tb[St.pp_off + k]];
                            const int ts = tb[Q.tsid_off + Cl.ridx];
                            if (ts < 0) continue;
                            cfac(Q, ts, r0, cf);
                            if (XF && Q.pref_slot >= 0) {       // biomass^prey_preference for this predator
                                const T p = SLOT(Q.pref_slot);
#pragma unroll
                                for (int b = 0; b < NB; b++) tp[b] = tp[b] + cf[b] * xpow(B0[b], p);
                                continue;
                            }
#pragma unroll
                            for (int b = 0; b < NB; b++) tp[b] = tp[b] + cf[b] * B0[b];
                        }
#pragma unroll
                        for (int b = 0; b < NB; b++) { cr[b] = tp[b]; de[b] = B0[b]; }
                        div_azx_n(cr, de);                      // consratio = tp / avoid_zero(B0)   (exact-division forms: g3_math.cuh)
                        dif_pmax_x_n(cr, 0.95, -1e3);           // dif_pmin(consratio, 0.95, 1e3)
                        T tpn[NB];
#pragma unroll
                        for (int b = 0; b < NB; b++) {
                            tpn[b] = B0[b] * cr[b];
                            if (b < nb) {
                                stT<T>(pNum + (unsigned)(r0 + b) * C, num[b] * (1.0 - cr[b]));
                                o1 = o1 + tp[b]; o2 = o2 + tpn[b];
                            }
                        }
                        if (catch_now) {
                            T cc[NB];
#pragma unroll
                            for (int b = 0; b < NB; b++) { cc[b] = T(1.0); de[b] = tp[b]; }
                            div_az_n(cc, de);
#pragma unroll
                            for (int b = 0; b < NB; b++) cc[b] = cc[b] * tpn[b] * B0[b];        // consconv * B0
                            for (int j = 0; j < St.cx_n; j++) {
                                const int x = tb[St.cx_off + j];
                                if (!((xact >> x) & 1)) continue;
                                T fx[NB];
#pragma unroll
                                for (int b = 0; b < NB; b++) fx[b] = T(0.0);
                                for (int k = 0; k < nq; k++) {
                                    const PairRec& Q = Tpp[tb[St.pp_off + k]];
                                    const int ts = tb[Q.tsid_off + Cl.ridx];
                                    if (ts < 0 || !((Q.cxmask >> j) & 1)) continue;
                                    cfac(Q, ts, r0, cf);
#pragma unroll
                                    for (int b = 0; b < NB; b++) fx[b] = fx[b] + cf[b];
                                }
#pragma unroll
                                for (int b = 0; b < NB; b++) if (b < nb) GS(Txb[x].g_x + c_lo + r0 + b, fx[b] * cc[b]);
                            }
                        }
                    }
                    GS(U.g_usp, o1); GS(U.g_usp + 1, o2);
                } else if (catch_now) {
                    for (int j = 0; j < St.cx_n; j++) {
                        const int x = tb[St.cx_off + j];
                        if ((xact >> x) & 1)
                            for (int r = 0; r < U.len; r++) GS(Txb[x].g_x + c_lo + r, T(0.0));       // no landings: cons = 0
                    }
                }
                if (U.up_ghost >= 0) {
#pragma unroll
                    for (int i = 0; i < NW - 1; i++) {
                        GS(U.up_ghost + i, ldT<T>(pNum + (unsigned)(U.len - (NW - 1) + i) * C));
                        GS(U.up_ghost + (NW - 1) + i, ldT<T>(pWgt + (unsigned)(U.len - (NW - 1) + i) * C));
                    }
                }
            }
            G3T_PH(3)           // (totals of the predation steps, if any)
            G3T_BAR();          // (pass 2 of a unit may run on another warp than its pass 1; ghost rows are in place)
            G3T_PH(4)           // pass 1: predation, ghost rows

            // g3a_spawn, order 6 (R/action_spawn.R:121-168), for the rows of a block once they have grown: spawningnum = num x
            // proportion, the block's part of the recruitment sum s, spawning mortality of the parents
            auto spawn_rows = [&](const StockRec& St, const ColRec& Cl, const int row0, const int nb, T (&num)[NB], T (&wgt)[NB], T& acc) {
                const int32_t* Sp = I + St.sp_ioff;
                const bool fec = Sp[G3SP_RECR_KIND] == G3SPAWN_FECUNDITY;
                T p3(0.0), p4(0.0);
                if (fec) { p3 = SLOT(Sp[G3SP_RECR_SLOTS + 3]); p4 = SLOT(Sp[G3SP_RECR_SLOTS + 4]); }
#pragma unroll
                for (int b = 0; b < NB; b++) {
                    if (b >= nb) continue;
                    const int l = row0 + b;
                    const T spn = num[b] * G(St.g_sprop + l);
                    if (fec) acc = acc + G(St.g_sfec + Cl.aidx * St.L + l) * xpow(spn, p3) * xpow(wgt[b], p4);
                    else acc = acc + wgt[b] * spn;
                    if (St.g_smort >= 0) num[b] = num[b] - spn * G(St.g_smort + l);
                    if (Sp[G3SP_WLOSS] >= 0) {      // spawning weight loss (R/action_spawn.R:166-181, R/action_weightloss.R:19-44)
                        const T mw = SLOT(Sp[G3SP_WLOSS + 1]);
                        const T lost = ((wgt[b] - mw) * (1.0 - SLOT(Sp[G3SP_WLOSS]))) + mw;
                        wgt[b] = ratio_add_pop(wgt[b], num[b] - spn, lost, spn);
                    }
                }
            };
            const bool spawn_step = XF && ((A.spawn_mask >> step) & 1);

            // ================= own units, pass 2: natural mortality, growth (+ maturing split), and -- unless maturing fish are
            // about to arrive -- renewal and the collect vectors. Length growth only looks DOWN the length axis, so sweeping
            // the blocks from the top the update is in place: every source row is still the old value when a block is rebuilt.
            G3T_UNITS(qn - ctr + 2, uk) {
                const UnitRec& U = Tunit[tb[A.uorder_off + uk]];
                const ColRec& Cl = Tcol[U.gc];
                const StockRec& St = Tst[Cl.s];
                const int L = St.L, c_lo = Cl.cell0 + U.lo;
                const int grow_n = St.grow_n;
                const bool trans_now = grow_n >= 0 && St.has_out && ((St.trans_mask >> step) & 1);
                const bool cont_now = trans_now && St.mat_kind == G3MAT_CONTINUOUS;
                const bool const_now = F_CMAT && trans_now && St.mat_kind == G3MAT_CONSTANT;
                const bool has_mort = St.mort_ioff >= 0;
                const bool inc_here = Cl.inc_tn >= 0 && ((Cl.inc_mask >> step) & 1);
                const int ren_mask = renewals_now(St);
                double* const pNum = SP(A.s_num + c_lo);
                double* const pWgt = SP(A.s_wgt + c_lo);
                bool spawn_here = false;    // this stock spawns at this step
                if (XF && spawn_step && St.sp_ioff >= 0) { const int rs = I[St.sp_ioff + G3SP_RUN_STEP]; spawn_here = rs == 0 || rs == step + 1; }
                if (!has_mort && grow_n < 0 && !inc_here && !ren_mask && xact == 0 && !spawn_here) continue;
                const int mset = XF && St.mort_tmap >= 0 ? I[St.mort_tmap + t] : 0;       // time-varying M: this step's table
                const T mortf = has_mort ? G(St.g_mort + (mset * A.ndt + idt) * St.ncol + Cl.col) : T(1.0);
                const int nblk = (U.len + NB - 1) / NB;
                T sp_acc(0.0);
                if (grow_n < 0) {
                    for (int j = nblk - 1; j >= 0; j--) {
                        const int r0 = j * NB, nb = min(NB, U.len - r0);
                        T num[NB], wgt[NB];
#pragma unroll
                        for (int b = 0; b < NB; b++) {
                            num[b] = ldT<T>(pNum + (unsigned)(r0 + b) * C) * mortf; wgt[b] = ldT<T>(pWgt + (unsigned)(r0 + b) * C);
                        }
                        if (XF && spawn_here) spawn_rows(St, Cl, U.lo + r0, nb, num, wgt, sp_acc);
                        if (!inc_here) finish_block(Cl, St, U.lo + r0, nb, num, wgt, false, ren_mask);
                        else {
#pragma unroll
                            for (int b = 0; b < NB; b++) if (b < nb) stT<T>(pNum + (unsigned)(r0 + b) * C, num[b]);
                        }
                    }
                    if (XF && spawn_here) GS(U.g_sps, sp_acc);
                    continue;
                }
                const T remf = trans_now ? G(Cl.g_remf) : T(0.0);
                const int gset = XF && St.grow_tmap >= 0 ? I[St.grow_tmap + t] : 0;       // time-varying growth: this step's tables
                const int bset = (((XF ? gset * (St.band_percol ? St.ncol : 1) : 0) + (St.band_percol ? Cl.col : 0)) * A.ndt + idt) * St.bsz + U.lo * NW;
                const double* const pB = A.use_h && St.h_band >= 0 ? HP(St.h_band + bset) : GP(St.g_band + bset);
                const double* const pBr = A.use_h && St.h_bandr >= 0 ? HP(St.h_bandr + bset) : GP(St.g_bandr + bset);
                const double* const pQ = A.use_h && St.h_wq >= 0 ? HP(St.h_wq + U.lo) : GP(St.g_wq + (XF ? gset * St.wq_stride : 0) + U.lo);
                const double* const pCm = St.g_cmat >= 0 ? GP(St.g_cmat + Cl.col * L + U.lo) : pQ;
                const double* const pGh = U.g_ghost >= 0 ? GP(U.g_ghost) : pQ;
                const bool stash = trans_now && Cl.s_tn >= 0;
                double* const pTn = stash ? SP(Cl.s_tn + U.lo) : pNum;
                double* const pTw = stash ? SP(Cl.s_tw + U.lo) : pWgt;
                bool ren_here = false;      // a renewal into this column at this step
                for (int r = 0; r < St.n_renew; r++)
                    ren_here = ren_here || (((ren_mask >> r) & 1) && Cl.aidx == tb[St.ren_off + 3 * r + 1]);
                const bool finish_here = !inc_here && (ren_here || xact);
                for (int j = nblk - 1; j >= 0; j--) {
                    const int r0 = j * NB, nb = min(NB, U.len - r0);
                    T an[NB], aw[NB], num[NB], wgt[NB], de[NB];
                    const bool low = U.g_ghost >= 0 && r0 < NW - 1;
                    if (!SMEM && j > 0) {       // state in the slab: the rows the NEXT block adds to the window, on their way now
#pragma unroll
                        for (int b = 0; b < NB; b++) {
                            const int rr = r0 - NB - (NW - 1) + b;
                            if (rr >= 0) { g3t_prefetch(pNum + (unsigned)rr * C); g3t_prefetch(pWgt + (unsigned)rr * C); }
                        }
                    }
                    if (cont_now) {
                        if (low) grow_gather<T, NW, NB, 1, true>(pNum, pWgt, pQ, pB, pBr, pCm, pGh, r0, mortf, an, aw, num, wgt);
                        else grow_gather<T, NW, NB, 1, false>(pNum, pWgt, pQ, pB, pBr, pCm, pGh, r0, mortf, an, aw, num, wgt);
                    } else if (const_now) {
                        if (low) grow_gather<T, NW, NB, 2, true>(pNum, pWgt, pQ, pB, pBr, pCm, pGh, r0, mortf, an, aw, num, wgt);
                        else grow_gather<T, NW, NB, 2, false>(pNum, pWgt, pQ, pB, pBr, pCm, pGh, r0, mortf, an, aw, num, wgt);
                    } else {
                        if (low) grow_gather<T, NW, NB, 0, true>(pNum, pWgt, pQ, pB, pBr, pCm, pGh, r0, mortf, an, aw, num, wgt);
                        else grow_gather<T, NW, NB, 0, false>(pNum, pWgt, pQ, pB, pBr, pCm, pGh, r0, mortf, an, aw, num, wgt);
                    }
                    // ---- mean weights: one avoid_zero + reciprocal per row and part, the chains of the block interleaved; where
                    // every row of the block has more than 0.034 fish on every lane avoid_zero() is the identity and the chain is
                    // skipped (div_az_n, g3_math.cuh) -- rows of that kind come in runs, so the blocks catch nearly all of them
#pragma unroll
                    for (int b = 0; b < NB; b++) de[b] = num[b];
                    div_az_n(wgt, de);
                    if (XF && spawn_here) spawn_rows(St, Cl, U.lo + r0, nb, num, wgt, sp_acc);      // (before the maturing fish move: order 6 < 7)
                    if (trans_now) {
                        if (stash) {
#pragma unroll
                            for (int b = 0; b < NB; b++) de[b] = an[b];
                            div_az_n(aw, de);
#pragma unroll
                            for (int b = 0; b < NB; b++)
                                if (b < nb) { stT<T>(pTn + (unsigned)(r0 + b) * C, an[b]); stT<T>(pTw + (unsigned)(r0 + b) * C, aw[b]); }
                        }
                        // unclaimed remainder moves back (move_remainder = TRUE, R/action_mature.R:126-131)
#pragma unroll
                        for (int b = 0; b < NB; b++) num[b] = num[b] + an[b] * remf;
                    }
                    if (finish_here) finish_block(Cl, St, U.lo + r0, nb, num, wgt, false, ren_mask);
                    else {
#pragma unroll
                        for (int b = 0; b < NB; b++)
                            if (b < nb) { stT<T>(pNum + (unsigned)(r0 + b) * C, num[b]); stT<T>(pWgt + (unsigned)(r0 + b) * C, wgt[b]); }
                    }
                }
                if (XF && spawn_here) GS(U.g_sps, sp_acc);
            }

            // ================= columns that receive maturing fish finish once every source column has grown. That part is
            // elementwise per cell, so its (column, block) units are dealt to ALL warps, not only to the owners.
            if ((A.inc_any_mask >> step) & 1) {
                G3T_BAR();
                G3T_PH(5)       // pass 2: growth
                int nitem = 0;          // blocks of the receiving columns (tb: A.inc_n global columns)
                for (int i = 0; i < A.inc_n; i++) {
                    const ColRec& Cl = Tcol[tb[A.inc_off + i]];
                    if ((Cl.inc_mask >> step) & 1) nitem += (Tst[Cl.s].L + NB - 1) / NB;
                }
                for (int k = G3T_GRAB(qn + 3); k < nitem; k = G3T_GRAB(qn + 3)) {
                    int i = 0, blk = k;
                    for (;; i++) {
                        const ColRec& Ci = Tcol[tb[A.inc_off + i]];
                        if (!((Ci.inc_mask >> step) & 1)) continue;
                        const int n = (Tst[Ci.s].L + NB - 1) / NB;
                        if (blk < n) break;
                        blk -= n;
                    }
                    const ColRec& Cl = Tcol[tb[A.inc_off + i]];
                    const StockRec& St = Tst[Cl.s];
                    const int ren_mask = renewals_now(St);
                    const int r0 = blk * NB, nb = min(NB, St.L - r0);
                    T num[NB], wgt[NB];
#pragma unroll
                    for (int b = 0; b < NB; b++) { const int l = min(r0 + b, St.L - 1); num[b] = S(A.s_num + Cl.cell0 + l); wgt[b] = S(A.s_wgt + Cl.cell0 + l); }
                    finish_block(Cl, St, r0, nb, num, wgt, true, ren_mask);
                }
            }

            // ================= g3a_spawn, order 8 (R/action_spawn.R:183-218): the offspring arrive in the youngest age of the
            // output stocks. s = the parent units' parts in unit order, r = the recruitment function, every (output stock, area,
            // block of rows) is dealt to a warp; a warp works out r of a parent when its first block of that parent comes up.
            if (XF && spawn_step) {
                G3T_BAR();
                int kblk = 0;
                for (int i = 0; i < A.sp_n; i++) {
                    const StockRec& Pst = Tst[tb[A.sp_off + i]];
                    const int32_t* Sp = I + Pst.sp_ioff;
                    if (Sp[G3SP_RUN_STEP] != 0 && Sp[G3SP_RUN_STEP] != step + 1) continue;
                    bool have_total = false;
                    T total(0.0);
                    for (int o = 0; o < Sp[G3SP_N_OUT]; o++) {
                        const StockRec& So = Tst[tb[Pst.sp_out_off + 4 * o]];
                        const int gn = tb[Pst.sp_out_off + 4 * o + 1], gw = tb[Pst.sp_out_off + 4 * o + 2];
                        const int nblk = (So.L + NB - 1) / NB;
                        for (int ab = 0; ab < So.narea * nblk; ab++, kblk++) {
                            if (kblk % W != warp) continue;
                            if (!have_total) {
                                have_total = true;
                                T sum(0.0);
                                for (int u = Pst.u0; u < Pst.u0 + Pst.nu; u++) sum = sum + G(Tunit[u].g_sps);
                                const int32_t* rs = Sp + G3SP_RECR_SLOTS;
                                const int kind = Sp[G3SP_RECR_KIND];
                                T r(0.0);
                                if (kind == G3SPAWN_FECUNDITY || kind == G3SPAWN_SIMPLESSB) r = SLOT(rs[0]) * sum;
                                else if (kind == G3SPAWN_RICKER) r = SLOT(rs[0]) * sum * xexp(-SLOT(rs[1]) * sum);
                                else if (kind == G3SPAWN_BEVERTONHOLT) r = SLOT(rs[0]) * sum / (SLOT(rs[1]) + sum);
                                else if (kind == G3SPAWN_BEVERTONHOLT_SS3) {
                                    const T h = SLOT(rs[0]), B0 = SLOT(rs[1]), Ry = SLOT(I[rs[2] + yidx]);
                                    r = 4.0 * h * sum * Ry / (B0 * (1.0 - h) + sum * (5.0 * h - 1.0));
                                } else r = SLOT(rs[0]) * lsa_c(-1000.0 * sum / SLOT(rs[1]), -1000.0) / -1000.0;
                                // stock__offspringnum = r / avoid_zero(number of cells) in every cell of the parent, summed again
                                const double ncell = (double)(Pst.ncol * Pst.L);
                                total = r * 1.0 / avoid_zero(T(ncell)) * ncell;
                            }
                            const int ar = ab / nblk, r0 = (ab - ar * nblk) * NB;
                            const T share = total * SLOT(tb[Pst.sp_out_off + 4 * o + 3]);
                            for (int b = 0; b < NB; b++) {
                                const int l = r0 + b;
                                if (l >= So.L) break;
                                const int cell = So.c0 + ar * So.nage * So.L + l;
                                const T n0 = S(A.s_num + cell), w0 = S(A.s_wgt + cell), spawned = G(gn + l) * share;
                                const T w1 = ratio_add_pop(w0, n0, G(gw + l), spawned);
                                SS(A.s_wgt + cell, w1);
                                SS(A.s_num + cell, n0 + spawned);
                                // a catch in numbers is cons / avoid_zero(wgt) with the weight the likelihood sees (order 10): the
                                // block's finish divided by the weight before the offspring arrived
                                if (xact)
                                    for (int j = 0; j < So.cx_n; j++) {
                                        const int x = tb[So.cx_off + j];
                                        if (!(((xact >> x) & 1) && Txb[x].unit == 0)) continue;
                                        GS(Txb[x].g_x + cell, G(Txb[x].g_x + cell) * avoid_zero(w0) / avoid_zero(w1));
                                    }
                            }
                        }
                    }
                }
            }

            // ================= g3l_distribution collect: model[dest] += lgmatrix . x (R/likelihood_distribution.R:288-364).
            // The planner cut the (dest, term) lists into TASKS of at most a few dozen terms and dealt them to the warps by
            // size (a survey index sums every cell of its stocks into ONE dest: 260 terms on C2). A task of an unsplit dest
            // adds to the model entry; the tasks of a split dest leave partial sums that the dest's owner adds up below.
            // Where the comparison needs the total of a group of dests (sumofsquares: the area group's whole slab;
            // multinomial: each slice), every warp also keeps its share of that total.
            if (cact) {
                G3T_BAR();
                G3T_PH(6)       // hand-over blocks (or pass 2 on steps without hand-over)
                for (int c = 0; c < A.NCOMP; c++) {
                    if (!((cact >> c) & 1)) continue;
                    const CompRec& Cm = Tcomp[c];
                    const int nd = Cm.n_dest;
                    const bool si = Cm.func == G3F_SI_LOG || Cm.func == G3F_SI_LINEAR;
                    const int ms = Cm.g_ms + (si ? I[Cm.tidx_ioff + t] * nd : 0);
                    const int xs = Txb[Cm.xbuf].g_x, xkind = Txb[Cm.xbuf].kind, xunit = Txb[Cm.xbuf].unit;
                    const bool tot_now = Cm.n_groups > 0 && ((cmp_now >> c) & 1);
                    if (tot_now) for (int g = 0; g < Cm.n_groups; g++) GS(Cm.g_pt + g * SHW + warp, T(0.0));
                    const int32_t* __restrict__ ci = M + A.m_cidx;
                    const double* __restrict__ cc = reinterpret_cast<const double*>(M + A.m_ccoef);
                    int gcur = -1;          // this warp's share of the current group's total stays in a register: its tasks
                    T gsum(0.0);            // come sorted by group, so the slot in memory is touched once per group
                    for (int k = tb[Cm.task_off + warp]; k < tb[Cm.task_off + warp + 1]; k += 5) {
                        const int e = tb[k], j0 = tb[k + 1], j1 = tb[k + 2], slot = tb[k + 3], grp = tb[k + 4];
                        const T old = slot < 0 ? G(ms + e) : T(0.0);        // (in flight with the terms' loads)
                        T acc(0.0);
                        // CU terms at a time: their loads are in flight together (a catch collect vector was written by other
                        // warps: every read is an L2 round trip), the sum keeps its order. Abundance needs no collect vector:
                        // it is the state itself, in shared memory, complete at this point of the step.
                        constexpr int CU = NTAN > 0 ? 2 : 8;
                        for (int j = j0; j < j1; j += CU) {
                            int ix[CU]; double cf[CU]; T xv[CU];
#pragma unroll
                            for (int u = 0; u < CU; u++) { const int jj = min(j + u, j1 - 1); ix[u] = ci[jj]; cf[u] = j + u < j1 ? cc[jj] : 0.0; }
                            if (xkind == 0) {
#pragma unroll
                                for (int u = 0; u < CU; u++) xv[u] = G(xs + ix[u]);
                                if (!SMEM && j + CU < j1) {     // (slabs larger than the L2: the next terms on their way under this sum)
#pragma unroll
                                    for (int u = 0; u < CU; u++) g3t_prefetch(GP(xs + ci[min(j + CU + u, j1 - 1)]));
                                }
                            } else {
#pragma unroll
                                for (int u = 0; u < CU; u++) xv[u] = xunit == 0 ? S(A.s_num + ix[u]) : S(A.s_num + ix[u]) * S(A.s_wgt + ix[u]);
                            }
#pragma unroll
                            for (int u = 0; u < CU; u++) if (j + u < j1) acc = acc + cf[u] * xv[u];
                        }
                        // slot < 0: the whole dest; slot = -2 - k: first part of split dest (takes the old value along), k: a later part
                        if (slot < 0) acc = acc + old;
                        if (slot == -1) GS(ms + e, acc);
                        else GS(Cm.g_cpart + (slot < 0 ? -2 - slot : slot), acc);
                        if (tot_now) {
                            if (grp != gcur) {
                                if (gcur >= 0) GS(Cm.g_pt + gcur * SHW + warp, G(Cm.g_pt + gcur * SHW + warp) + gsum);
                                gcur = grp; gsum = T(0.0);
                            }
                            gsum = gsum + acc;
                        }
                    }
                    if (tot_now && gcur >= 0) GS(Cm.g_pt + gcur * SHW + warp, G(Cm.g_pt + gcur * SHW + warp) + gsum);
                }
                G3T_BAR();
                G3T_PH(7)       // collect tasks
                // ================= g3l_distribution compare (R/likelihood_distribution.R:375-394): the dests of a component are
                // dealt warp by warp (continuing from component to component); every warp leaves its share of the component's value
                int rot = 0;
                for (int c = 0; c < A.NCOMP; c++) {
                    if (!((cact >> c) & 1)) continue;
                    const CompRec& Cm = Tcomp[c];
                    const int ti = I[Cm.tidx_ioff + t];
                    const int func = Cm.func, nd = Cm.n_dest, nt = Cm.n_times;
                    const bool si = func == G3F_SI_LOG || func == G3F_SI_LINEAR;
                    const int ms = Cm.g_ms + (si ? ti * nd : 0);
                    const int e0 = warp >= rot ? warp - rot : warp - rot + W;       // dest e belongs to warp (e + rot) % W
                    const int rot0 = rot;
                    rot = (rot + nd) % W;
                    // split dests: the owner adds the parts up, in order
                    for (int k = 0; k < tb[Cm.split_off]; k++) {
                        const int e = tb[Cm.split_off + 1 + 3 * k];
                        if ((e + rot0) % W != warp) continue;
                        const int s0 = tb[Cm.split_off + 2 + 3 * k], sn = tb[Cm.split_off + 3 + 3 * k];
                        T v(0.0);
                        for (int i = 0; i < sn; i++) v = v + G(Cm.g_cpart + s0 + i);
                        GS(ms + e, v);
                    }
                    if (!((cmp_now >> c) & 1)) continue;
                    const int nb = Cm.n_bins;
                    T part(0.0);
                    if (func == G3F_SUMOFSQUARES || (F_LIKE && func == G3F_MULTINOMIAL)) {
                        // R/likelihood_distribution.R:3-39 (sum of squared differences of proportions) and :41-60
                        const double* op = A.obsprop + Cm.obs_off + (size_t)ti * nd;
                        const double eps = R[Cm.param_roff];
                        int gcur = -1;
                        T rt(0.0);
                        for (int e = e0; e < nd; e += W) {
                            const int g = tb[Cm.grp_off + e];
                            if (g != gcur) {        // total of the group: the warps' shares, in warp order
                                gcur = g;
                                T tot(0.0);
                                for (int w = 0; w < W; w++) tot = tot + G(Cm.g_pt + g * SHW + w);
                                rt = 1.0 / avoid_zero(tot);
                            }
                            if (!F_LIKE || func == G3F_SUMOFSQUARES) {
                                const T dlt = G(ms + e) * rt - op[e];
                                part = part + dlt * dlt;
                            } else
                                part = part + op[e] * xlog(dif_pmax_c(G(ms + e) * rt, 1.0 / (nb * eps), 1e4));
                            GS(ms + e, T(0.0));      // zero counters for the next period
                        }
                    } else if (F_LIKE && func == G3F_SUMSQLOGRATIOS) {    // R/likelihood_distribution.R:127-136 (obsprop holds log(obs + eps))
                        const double eps = R[Cm.param_roff];
                        const double* op = A.obsprop + Cm.obs_off + (size_t)ti * nd;
                        for (int e = e0; e < nd; e += W) {
                            const T dlt = op[e] - xlog(G(ms + e) + eps);
                            part = part + dlt * dlt;
                            GS(ms + e, T(0.0));
                        }
                    } else if (ti == nt - 1) {              // surveyindices: R/likelihood_distribution.R:82-125
                        const double fa = R[Cm.param_roff], fb = R[Cm.param_roff + 1];
                        const int series = Cm.g_ms;
                        const double* op = A.obsprop + Cm.obs_off;
                        for (int e = e0; e < nd; e += W) {
                            T sN(0.0); double sI = 0.0;
                            for (int i = 0; i < nt; i++) {
                                const T mv = G(series + i * nd + e);
                                sN = sN + (!F_LIKE || func == G3F_SI_LOG ? xlog(avoid_zero(mv)) : mv);
                                sI += op[(size_t)i * nd + e];
                            }
                            const T mN = sN / (double)nt;
                            const double mI = sI / (double)nt;
                            T beta(fb), alpha(fa);
                            if (isnan(fb)) {
                                T sxy(0.0), sxx(0.0);
                                for (int i = 0; i < nt; i++) {
                                    const T mv = G(series + i * nd + e);
                                    const T N = !F_LIKE || func == G3F_SI_LOG ? xlog(avoid_zero(mv)) : mv;
                                    sxy = sxy + (op[(size_t)i * nd + e] - mI) * (N - mN); sxx = sxx + (N - mN) * (N - mN);
                                }
                                beta = sxy / avoid_zero(sxx);
                            }
                            if (isnan(fa)) alpha = mI - beta * mN;
                            for (int i = 0; i < nt; i++) {
                                const T mv = G(series + i * nd + e);
                                const T N = !F_LIKE || func == G3F_SI_LOG ? xlog(avoid_zero(mv)) : mv;
                                const T r = alpha + beta * N - op[(size_t)i * nd + e];
                                part = part + r * r;
                            }
                        }
                    }
                    GS(Cm.g_pv + warp, part);
                }
            }
            G3T_PH(8)           // compare (collect steps) / hand-over blocks (other steps)
            G3T_BAR();          // the step's state, collect sums and component values are complete
            G3T_PH(9)           // waiting at the end-of-step barrier

            // ================= the keeper warp: the objective, in the reference's summation order (the unit warps go on)
            if (keeper) {
                for (int c = 0; c < A.NCOMP; c++) {
                    if (!((cact >> c) & 1)) continue;
                    const CompRec& Cm = Tcomp[c];
                    if (!((cmp_now >> c) & 1)) continue;
                    const bool si = Cm.func == G3F_SI_LOG || Cm.func == G3F_SI_LINEAR;
                    if (si && I[Cm.tidx_ioff + t] != Cm.n_times - 1) continue;      // survey indices score once, at their last time
                    if (lane_on && ((cw_lane >> c) & 1)) {
                        T cnll(0.0);        // the warps' shares of the component's value, in warp order
                        for (int w = 0; w < W; w++) cnll = cnll + G(Cm.g_pv + w);
                        if (Cm.func == G3F_MULTINOMIAL) cnll = 2.0 * (-cnll + A.obsconst[Cm.obsconst_off + I[Cm.tidx_ioff + t]]);
                        nll = nll + mkpar_g<T>(prow, seed, Cm.weight_par) * cnll;
                        GS(A.g_ctot + c, G(A.g_ctot + c) + cnll);
                    }
                }
                // ---- g3l_random_dnorm / g3l_random_walk (R/likelihood_random.R:14-109): formulas of parameters only
                if (F_LIKE)
                    for (int r = 0; r < I[G3H_NRAND]; r++) {
                        const int32_t* const Rn = I + I[G3H_RAND_OFF] + r * G3RN_LEN;
                        const int run = I[Rn[G3RN_RUN_OFF] + t];
                        if (!run) continue;
                        const T sd = SLOT(I[Rn[G3RN_SIGMA_OFF] + t]);
                        const T resid = (SLOT(I[Rn[G3RN_PARAM_OFF] + t]) - SLOT(I[Rn[G3RN_MEAN_OFF] + t])) / sd;
                        const T logans = T(-0.91893853320467274178) - xlog(sd) - 0.5 * resid * resid;     // TMB dnorm
                        const T n = Rn[G3RN_LOG] ? T(0.0) - logans : xexp(logans);
                        if (lane_on && (run != 2 || t == total_steps)) {        // (2: on the last step of this vector's run)
                            nll = nll + SLOT(Rn[G3RN_WEIGHT]) * n;
                            GS(A.g_ctot + A.NCOMP + r, G(A.g_ctot + A.NCOMP + r) + n);
                        }
                    }
                // ---- g3l_understocking (R/likelihood_understocking.R:36-46)
                if (have_us) {
                    T us_tot(0.0);
                    if (pred_now)
                        for (int s = 0; s < NS; s++) {
                            const StockRec& St = Tst[s];
                            if (!St.us_flag || St.pp_n == 0) continue;
                            T o1(0.0), o2(0.0);
                            for (int u = St.u0; u < St.u0 + St.nu; u++) { o1 = o1 + G(Tunit[u].g_usp); o2 = o2 + G(Tunit[u].g_usp + 1); }
                            us_tot = us_tot + (o1 - o2);
                        }
                    const T u = (us_power == 2.0) ? us_tot * us_tot : xpow(us_tot, T(us_power));
                    if (lane_on) {
                        nll = nll + us_weight * u;
                        GS(A.g_ctot + A.NNLL - 2, G(A.g_ctot + A.NNLL - 2) + u);
                    }
                }
            }

            // ================= g3a_age (R/action_age.R:51-85) + movement hand-off: length group by length group
            if (final_step) {
                bool age_any = false;
                for (int s = 0; s < NS; s++) age_any = age_any || Tst[s].age_enable;
                if (age_any) {
                    for (int l = warp; l < A.maxL; l += W) {
                        for (int s = 0; s < NS; s++) {
                            const StockRec& St = Tst[s];
                            if (!St.age_enable || l >= St.L) continue;
                            const int L = St.L, nage = St.nage, narea = St.narea;
                            for (int r = 0; r < narea; r++) {
                                const int i = r * L + l;
                                const int base = St.c0 + r * nage * L + l;
                                const int top = base + (nage - 1) * L;
                                const int nb_ = A.s_num, wb_ = A.s_wgt;
                                if (St.age_n_out > 0) { SS(St.s_mv + i, S(nb_ + top)); SS(St.s_mv + narea * L + i, S(wb_ + top)); }
                                if (nage == 1) {
                                    if (St.age_n_out > 0) SS(nb_ + top, T(0.0));
                                    continue;
                                }
                                if (St.age_n_out > 0) { SS(nb_ + top, S(nb_ + top - L)); SS(wb_ + top, S(wb_ + top - L)); }
                                else {
                                    const T n0 = S(nb_ + top), n1_ = S(nb_ + top - L);
                                    SS(wb_ + top, ratio_add_pop(S(wb_ + top), n0, S(wb_ + top - L), n1_));
                                    SS(nb_ + top, n0 + n1_);
                                }
                                for (int a = nage - 2; a >= 1; a--) {
                                    SS(nb_ + base + a * L, S(nb_ + base + (a - 1) * L));
                                    SS(wb_ + base + a * L, S(wb_ + base + (a - 1) * L));
                                }
                                SS(nb_ + base, T(0.0));
                            }
                        }
                        for (int i = tb[A.agedst_l_off + l]; i < tb[A.agedst_l_off + l + 1]; i++) {
                            const int4 e = Tagedst[i];     // (destination cell, stash entry num, stash entry wgt, ratio slot); same l
                            if (e.x < 0) continue;
                            const T n0 = S(A.s_num + e.x);
                            const T moved = S(e.y) * SLOT(e.w);
                            SS(A.s_wgt + e.x, ratio_add_pop(S(A.s_wgt + e.x), n0, S(e.z), moved));
                            SS(A.s_num + e.x, n0 + moved);
                        }
                    }
                    G3T_BAR();
                }
            }
            G3T_PH(10)          // keeper / ageing
        }   // time loop

        if (keeper && live) {
            const T out = bad_time ? T(nan("")) : nll;
            if (tile_t == 0) {
                A.nll[b] = val(out);
                if (A.comp_out) for (int c = 0; c < A.NNLL; c++) A.comp_out[b * A.NNLL + c] = bad_time ? nan("") : val(G(A.g_ctot + c));
            }
            if (NTAN > 0 && A.grad)
                for (int k = 0; k < NTAN; k++)
                    if (tile_t * NTAN + k < A.K) A.grad[b * A.K + tile_t * NTAN + k] = tan_of(out, k);
        }
    }
    G3T_PH_FLUSH()
#undef HP
#undef GP
#undef SP
#undef SLOT
#undef G3T_UNITS
}

#ifdef __CUDACC__
template <class T, int NW, bool SMEM, int MAXT, int NBV = 0, bool XF = false, int LEAN = 0>
__global__ void __launch_bounds__(MAXT, 1) eval_tile_kernel(const __grid_constant__ TArgs A) {
    extern __shared__ __align__(16) double g3t_smem[];
    __shared__ int g3t_ctr[8];
    tile_body<T, NW, SMEM, NBV, XF, LEAN>(A, g3t_smem, g3t_ctr, threadIdx.x >> 5, threadIdx.x & 31, blockIdx.x, gridDim.x);
}
#endif

}  // namespace g3t
