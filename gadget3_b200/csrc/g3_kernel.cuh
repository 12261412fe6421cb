// g3_kernel.cuh -- the persistent per-parameter-set CTA that evaluates a gadget3 objective.
//
// One CTA walks the whole time loop of one parameter vector (and one tile of tangent directions)
// with the stock state resident on chip: thread i owns stock cell i (length x age x area, all
// stocks concatenated) in registers; shared memory holds the neighbour-exchange copies, the
// per-evaluation tables (suitability, banded growth, renewal shape) and the likelihood slabs.
// Action order and arithmetic follow the reference's generated objective
// (baseline/multiple-substocks.cpp:692-2161); what is restructured relative to it:
//   * time-invariant terms are hoisted out of the step loop (suitability, exp(-M dt), growth and
//     maturity tables, renewal length distribution, observation proportions);
//   * growth is a banded gather over the n+1 source length groups instead of dense L x L matrices
//     (R/action_grow.R:233-335), the beta-binomial uses the lgamma-free product form
//      and the plus-group tail sums are folded into the band;
//   * logspace_add takes an exact early-out when the log1p term cannot change the result.
#pragma once
#include "g3_math.cuh"
#include "../../include/g3cuda_spec.h"

namespace g3 {

constexpr int MAXS = 8;      // stocks
constexpr int MAXC = 16;     // likelihood components
constexpr int MAXQ = 4;      // predator-prey pairs acting on one stock
constexpr int MAXTS = 8;     // (predator, area) total-suitability accumulators
constexpr int MAXX = 4;      // collect vectors
constexpr int MAXRN = 4;     // renewal records per stock
constexpr int NRED = 8;      // rotating reduction rows

struct KArgs {
    const int32_t* I; const double* R;
    const int4* cell;            // per cell: stock, l, col (= area_idx*nage + age_idx), area_idx
    const int32_t* inc_src;      // maturation: source cell feeding this cell, or -1
    const int32_t* inc_slot;     //   ratio slot
    const int32_t* inc_mask;     //   transition mask of the source stock
    const int32_t* age_src;      // ageing movement: source (oldest-column) cell, or -1
    const int32_t* age_slot;
    const int32_t* rem_off;      // per cell: offset into rem_slots list of ratio slots with a destination
    const int32_t* rem_slots;    //   [-1 terminated]
    const int32_t* pp_tsid;      // [NPP*maxnarea] total-suitability accumulator of (pp, prey area idx) or -1
    const int32_t* pp_eoff;      // [NPP*maxnarea] reals offset of E(t=0) for that predator area
    const int32_t* pp_estride;   // [NPP*maxnarea]
    const double* obsprop;       // per component: obs / avoid_zero(total) (sumofsquares), raw obs (multinomial), log(avoid_zero(obs)) (SI)
    const double* obsconst;      // per component/time: multinomial constant term
    const double* par; long long B;
    const int32_t* tangent_idx; int K;
    double* nll; double* comp; double* grad; double* report;
    int NC, NS, NPP, NTS, NCOMP, NX, NSLOT, NPAR, NT, NSTEPS, NNLL, NBOUND, maxnarea, any_const_mat;
    int stock_pp_off[MAXS + 1]; int stock_pp[MAXS * MAXQ];
    int us_flag[MAXS];
    int obs_off[MAXC]; int obsconst_off[MAXC];
    int trans_any[32];           // per step of year: some stock transitions
    // shared-memory layout, element offsets into the T region (par lives first, as doubles)
    int sm_par_bytes;
    int o_slot, o_num, o_wgt, o_tn, o_tw, o_alt, o_xg, o_pw, o_suit, o_gscr, o_red, o_rwt;
    int o_g[MAXS], o_gr[MAXS], o_gn[MAXS], o_pws[MAXS], o_rd[MAXS], o_ms[MAXC], o_xslot[MAXX];
    int maxL;
    long long rep_dstart[MAXS]; long long rep_model[MAXC]; long long rep_params;
    long long rep_renew[MAXS]; long long rep_pp[32];   // detail section (thread kernel): renewalnum per stock; per pair: suit report, suit, cons
    // ---- warp-per-evaluation kernel (g3_warp_kernel.cuh)
    int w_warp_bytes;                       // shared memory per warp
    int w_G[MAXS], w_ntile[MAXS];           // columns per tile, tiles per stock
    int w_ndt, w_dt_len[4], w_step_idt[32]; // distinct step lengths (months) and the one each step uses
    const unsigned char* w_pred_active;     // [NT] some landing is non-zero at t
    int w_o_mort, w_o_tot, w_o_mv;
    int w_mort_off[MAXS], w_o_dw[MAXS], w_o_rw[MAXS], w_tr_off[MAXS], w_mv_off[MAXS], w_inc_steps[MAXS];
    int w_ts_eoff[MAXTS], w_ts_estr[MAXTS]; // landings table of each (predator, area) accumulator
    int t_cblk;                             // thread kernel, wide bands: columns per growth sweep block
    int w_ts_kind[MAXTS];                   // G3PRED_* of the accumulator's fleet
    int w_x_kind[MAXX], w_x_unit[MAXX]; unsigned w_x_ppmask[MAXX];
    const int32_t* w_inc_tr;                // per cell: compact index of its maturation source in s_tn/s_tw, or -1
    int w_n_agedst; const int4* w_agedst;
    // ---- thread-per-evaluation kernel (g3_thread_kernel.cuh): workspace entries, [entry][thread]
    void* t_ws;
    int t_slot, t_num, t_wgt, t_tn, t_tw, t_mv, t_suit, t_scr, t_ctot;
    int t_mort[MAXS], t_pw[MAXS], t_g[MAXS], t_gr[MAXS], t_gn[MAXS], t_rd[MAXS], t_rw[MAXS], t_ms[MAXC];
    int t_x[MAXX];                          // collect vectors
    int t_suitq[32];                        // per predator-prey pair: suitability [predator length][prey length]
    int t_pts[MAXS], t_pmc[MAXS];           // per stock predator: totals / consumption factors [area][length]; max consumption [length]
    int t_mig[MAXS], t_mig_steps[MAXS];     // migration matrices [step][dest][src]; steps on which the stock migrates
    int t_remf, t_colbase[MAXS], t_cmat[MAXS];      // per global column: unclaimed share; first global column; constant-maturity ratios
    const int4* t_colinc;                   // per global column: (compact index of the source column's first cell or -1, ratio slot, step mask, 0)   // ageing movement: (destination cell, stash index num, stash index wgt, ratio slot)
};

template <class T> __device__ __forceinline__ T mkpar(const double* spar, const int* seed, int i) {
    T r(spar[i]);
    if constexpr (Tan<T>::n > 0) {
#pragma unroll
        for (int k = 0; k < Tan<T>::n; k++) if (seed[k] == i) r.d[k] = 1.0;
    }
    return r;
}

// slot programs (include/g3cuda_spec.h G3OP_*)
template <class T> __device__ T eval_slot(const int32_t* __restrict__ I, const double* __restrict__ R,
                                          const double* spar, const int* seed, int slot) {
    const int32_t* tab = I + I[G3H_SLOT_OFF] + 2 * slot;
    const int32_t* c = I + tab[0];
    int n = tab[1];
    T st[G3_SLOT_STACK];
    int sp = 0;
    for (int i = 0; i < n; i++) {
        int op = c[i];
        switch (op) {
        case G3OP_CONST: st[sp++] = T(R[c[++i]]); break;
        case G3OP_PARAM: st[sp++] = mkpar<T>(spar, seed, c[++i]); break;
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

// Block-wide sum delivered to every thread. Partials go through a rotating row of `red`
// (NRED rows x 32 warps): a row is only reused after NRED-1 further barriers.
template <class T> __device__ __forceinline__ T block_sum(T v, T* red, int& row) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    T* r = red + row * 32;
    if (lane == 0) r[warp] = v;
    __syncthreads();
    T tot = r[0];
    for (int w = 1; w < nw; w++) tot = tot + r[w];
    row = (row + 1 == NRED) ? 0 : row + 1;
    return tot;
}

template <class T, int MAXT>
__global__ void __launch_bounds__(MAXT) eval_kernel(const __grid_constant__ KArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int32_t* __restrict__ I = A.I;
    const double* __restrict__ R = A.R;
    double* spar = reinterpret_cast<double*>(smem_raw);
    T* sm = reinterpret_cast<T*>(smem_raw + A.sm_par_bytes);
    T* slotv = sm + A.o_slot;
    T* s_num = sm + A.o_num;
    T* s_wgt = sm + A.o_wgt;
    T* s_tn = sm + A.o_tn;
    T* s_tw = sm + A.o_tw;
    T* s_alt = sm + A.o_alt;
    T* s_suit = sm + A.o_suit;
    T* s_gscr = sm + A.o_gscr;
    T* s_red = sm + A.o_red;
    T* s_rwt = sm + A.o_rwt;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = (blockDim.x + 31) >> 5;
    const int NC = A.NC, NS = A.NS, NSTEPS = A.NSTEPS, NT = A.NT;
    constexpr int NTAN = Tan<T>::n;
    const int ntiles = NTAN > 0 ? (A.K + NTAN - 1) / NTAN : 1;

    // ---- per-thread constants: which cell this thread owns
    const bool has_cell = tid < NC;
    const int4 ci = has_cell ? A.cell[tid] : make_int4(0, 0, 0, 0);
    const int s = ci.x, l = ci.y, col = ci.z, ridx = ci.w;
    const int32_t* St = I + I[G3H_STOCK_OFF] + s * G3S_LEN;
    const int L = St[G3S_L], nage = St[G3S_NAGE], narea = St[G3S_NAREA], c0 = St[G3S_CELL_OFF];
    const int aidx = col - ridx * nage;
    const double midlen = R[St[G3S_MIDLEN_ROFF] + l];
    const double plusdl = R[St[G3S_PLUSDL_ROFF]];
    const int grow_n = has_cell ? St[G3S_GROW_N] : -1;
    const int mat_kind = St[G3S_MAT_KIND], trans_mask = St[G3S_TRANS_MASK], n_out = St[G3S_N_OUT];
    const int mort_slot = (has_cell && St[G3S_MORT_OFF] >= 0) ? I[St[G3S_MORT_OFF] + col] : -1;
    const int my_inc_src = has_cell ? A.inc_src[tid] : -1;
    const int my_inc_slot = has_cell ? A.inc_slot[tid] : -1;
    const int my_inc_mask = has_cell ? A.inc_mask[tid] : 0;
    const int my_age_src = has_cell ? A.age_src[tid] : -1;
    const int my_age_slot = has_cell ? A.age_slot[tid] : -1;
    const int my_rem_off = has_cell ? A.rem_off[tid] : -1;
    const bool is_us = has_cell && A.us_flag[s];
    const int nq = has_cell ? A.stock_pp_off[s + 1] - A.stock_pp_off[s] : 0;
    int myq[MAXQ], myts[MAXQ], myeoff[MAXQ], myestr[MAXQ];
#pragma unroll
    for (int k = 0; k < MAXQ; k++) {
        myq[k] = -1; myts[k] = -1; myeoff[k] = 0; myestr[k] = 0;
        if (k < nq) {
            int q = A.stock_pp[A.stock_pp_off[s] + k];
            myq[k] = q;
            myts[k] = A.pp_tsid[q * A.maxnarea + ridx];
            myeoff[k] = A.pp_eoff[q * A.maxnarea + ridx];
            myestr[k] = A.pp_estride[q * A.maxnarea + ridx];
        }
    }
    const T* my_pw = sm + A.o_pws[s];
    const T* my_g = sm + A.o_g[s];
    const T* my_gr = sm + A.o_gr[s];
    const T* my_gn = sm + A.o_gn[s];
    const int age_enable = has_cell ? St[G3S_AGE_ENABLE] : 0, age_n_out = St[G3S_AGE_N_OUT];
    const int n_renew = has_cell ? St[G3S_N_RENEW] : 0;
    const double us_power = R[I[G3H_US_ROFF]], us_weight = R[I[G3H_US_ROFF] + 1];
    const bool have_us = I[G3H_US_N] > 0;

    for (long long work = blockIdx.x; work < A.B * ntiles; work += gridDim.x) {
        const long long b = work / ntiles;
        const int tile = (int)(work - b * ntiles);
        int seed[NTAN > 0 ? NTAN : 1];
#pragma unroll
        for (int k = 0; k < (NTAN > 0 ? NTAN : 1); k++) {
            int g = tile * NTAN + k;
            seed[k] = (NTAN > 0 && g < A.K) ? A.tangent_idx[g] : -1;
        }
        int rrow = 0;
        __syncthreads();      // previous evaluation fully retired before shared memory is reused
        for (int i = tid; i < A.NPAR; i += blockDim.x) spar[i] = A.par[b * A.NPAR + i];
        __syncthreads();

        // ---- g3a_time (R/action_time.R:75-91)
        const double retro = I[G3H_PAR_RETRO] >= 0 ? spar[I[G3H_PAR_RETRO]] : 0.0;
        const double proj = I[G3H_PAR_PROJECT] >= 0 ? spar[I[G3H_PAR_PROJECT]] : 0.0;
        const int total_steps = NSTEPS * (I[G3H_END_YEAR] - (int)floor(retro) - I[G3H_START_YEAR] + (int)floor(proj)) + NSTEPS - 1;
        const bool bad_time = !(retro >= 0.0) || !(proj >= 0.0) || total_steps + 1 > NT;

        // ---- scalar slots, once per parameter vector
        for (int i = tid; i < A.NSLOT; i += blockDim.x) slotv[i] = eval_slot<T>(I, R, spar, seed, i);
        __syncthreads();

        // ---- parameter-dependent tables that do not depend on the step length
        // suitability per predator-prey pair (R/suitability.R:5-13), midlen^wbeta for growth,
        // renewal shape dnorm(midlen, mean, avoid_zero(sd)) (R/action_renewal.R:56-69)
        for (int i = tid; i < A.NPP * A.maxL; i += blockDim.x) {
            int q = i / A.maxL, ll = i - q * A.maxL;
            const int32_t* Q = I + I[G3H_PP_OFF] + q * G3PP_LEN;
            const int32_t* Ps = I + I[G3H_STOCK_OFF] + Q[G3PP_PREY] * G3S_LEN;
            if (ll < Ps[G3S_L]) {
                const int32_t* ss = I + Q[G3PP_SUIT_OFF];
                double ml = R[Ps[G3S_MIDLEN_ROFF] + ll];
                T v;
                if (Q[G3PP_SUIT_KIND] == G3SUIT_EXPL50) v = 1.0 / (1.0 + xexp(-slotv[ss[0]] * (ml - slotv[ss[1]])));
                else v = slotv[ss[0]];
                s_suit[i] = v;
            }
        }
        for (int ss2 = 0; ss2 < NS; ss2++) {
            const int32_t* S2 = I + I[G3H_STOCK_OFF] + ss2 * G3S_LEN;
            int L2 = S2[G3S_L];
            if (S2[G3S_GROW_N] >= 0 && tid < L2)
                sm[A.o_pws[ss2] + tid] = xpow(R[S2[G3S_MIDLEN_ROFF] + tid], slotv[I[S2[G3S_GROW_OFF] + 4]]);
            for (int k = 0; k < S2[G3S_N_RENEW]; k++) {
                const int32_t* Rn = I + S2[G3S_RENEW_OFF] + k * G3R_LEN;
                if (tid < L2) {
                    double ml = R[S2[G3S_MIDLEN_ROFF] + tid];
                    sm[A.o_rd[ss2] + k * L2 + tid] = dnorm(ml, slotv[Rn[G3R_MEAN]], avoid_zero(slotv[Rn[G3R_SD]]));
                    s_rwt[A.o_rd[ss2] - A.o_rd[0] + k * L2 + tid] = slotv[Rn[G3R_WALPHA]] * xpow(ml, slotv[Rn[G3R_WBETA]]);
                }
            }
        }
        // ---- g3a_initialconditions (R/action_renewal.R:83-163): unnormalised density first
        T num(0.0), wgt(1.0);
        T dens(0.0);
        const bool has_init = has_cell && St[G3S_INIT_OFF] >= 0;
        if (has_init) {
            const int32_t* sl = I + St[G3S_INIT_OFF] + col * 5;
            dens = dnorm(midlen, slotv[sl[0]], avoid_zero(slotv[sl[1]]));
            s_tn[tid] = dens;
        }
        __syncthreads();
        if (has_init) {
            const int32_t* sl = I + St[G3S_INIT_OFF] + col * 5;
            T tot(0.0);
            const T* colp = s_tn + (tid - l);
            for (int k = 0; k < L; k++) tot = tot + colp[k];
            num = dens / avoid_zero(tot) * 10000.0 * slotv[sl[2]];          // normalize_vec, R/aab_env.R:37-41
            wgt = slotv[sl[3]] * xpow(midlen, slotv[sl[4]]);
        }
        // renewal shapes: normalise in place (each thread reads the whole column, then all write)
        T rnorm[MAXRN];
#pragma unroll
        for (int k = 0; k < MAXRN; k++) {
            rnorm[k] = T(0.0);
            if (k < n_renew && aidx == I[St[G3S_RENEW_OFF] + k * G3R_LEN + G3R_AGE_IDX]) {
                const T* d = sm + A.o_rd[s] + k * L;
                T tot(0.0);
                for (int j = 0; j < L; j++) tot = tot + d[j];
                rnorm[k] = d[l] / avoid_zero(tot) * 10000.0;
            }
        }
        __syncthreads();

        T nll(0.0);
        T comp_us(0.0), comp_bounds(0.0);
        // ---- g3l_bounds_penalty (R/likelihood_bounds.R:13-16), cur_time == 0
        if (A.NBOUND > 0) {
            T pen(0.0);
            const double bw = R[I[G3H_BOUND_WS_ROFF]], bs = R[I[G3H_BOUND_WS_ROFF] + 1];
            for (int i = tid; i < A.NBOUND; i += blockDim.x) {
                T p = mkpar<T>(spar, seed, I[I[G3H_BOUND_OFF] + i]);
                double lo = R[I[G3H_BOUND_ROFF] + 2 * i], up = R[I[G3H_BOUND_ROFF] + 2 * i + 1];
                T a = lsa_c(bs * (p - up) / (up - lo), 0.0) + lsa_c(bs * (lo - p) / (up - lo), 0.0);
                pen = pen + a * a;
            }
            pen = block_sum(pen, s_red, rrow);
            nll = nll + bw * pen;
            comp_bounds = pen;
            if (A.report && tid == 0) A.report[1 + (size_t)(A.NNLL - 1) * NT] = val(pen);
        }

        T comp_tot[MAXC];
#pragma unroll
        for (int c = 0; c < MAXC; c++) comp_tot[c] = T(0.0);
        // zero the likelihood slabs
        for (int c = 0; c < A.NCOMP; c++) {
            const int32_t* C = I + I[G3H_COMP_OFF] + c * G3C_LEN;
            bool si = C[G3C_FUNC] == G3F_SI_LOG || C[G3C_FUNC] == G3F_SI_LINEAR;
            int n = C[G3C_N_DEST] * (si ? C[G3C_N_TIMES] : 1);
            for (int i = tid; i < n; i += blockDim.x) sm[A.o_ms[c] + i] = T(0.0);
        }
        int lastcalc = -1;
        T mortf(1.0);
        int mort_calc = -1;
        T tn(0.0), tw(0.0);

        for (int t = 0; t <= total_steps && !bad_time; t++) {
            const int step = t % NSTEPS, yidx = t / NSTEPS;
            const int steplen = I[I[G3H_STEPLEN_OFF] + step];
            const double dt = steplen / 12.0;
            const bool final_step = step == NSTEPS - 1;

            if (A.report && has_cell) {     // dstart_<stock>__num/wgt (R/action_report.R:177-213)
                size_t ncs = (size_t)L * nage * narea;
                A.report[A.rep_dstart[s] + (size_t)t * ncs + (tid - c0)] = val(num);
                A.report[A.rep_dstart[s] + (size_t)NT * ncs + (size_t)t * ncs + (tid - c0)] = val(wgt);
            }

            // ================= g3a_predate (R/action_predate.R:240-420)
            const T B0 = num * wgt;
            {   // step 1: suitable biomass per (predator, area), one barrier for all accumulators
                T part[MAXTS];
#pragma unroll
                for (int k = 0; k < MAXTS; k++) part[k] = T(0.0);
#pragma unroll
                for (int k = 0; k < MAXQ; k++)
                    if (myts[k] >= 0) {
                        T sv = s_suit[myq[k] * A.maxL + l] * B0;
#pragma unroll
                        for (int j = 0; j < MAXTS; j++) if (j == myts[k]) part[j] = part[j] + sv;
                    }
#pragma unroll
                for (int k = 0; k < MAXTS; k++)
                    if (k < A.NTS) {
                        T v = warp_sum(part[k]);
                        if (lane == 0) s_red[(NRED + k) * 32 + warp] = v;
                    }
            }
            __syncthreads();
            T cfac[MAXQ];
            T tp(0.0);
#pragma unroll
            for (int k = 0; k < MAXQ; k++) {
                cfac[k] = T(0.0);
                if (myts[k] >= 0) {
                    const T* r = s_red + (NRED + myts[k]) * 32;
                    T tot = r[0];
                    for (int w = 1; w < nwarp; w++) tot = tot + r[w];
                    double E = R[myeoff[k] + (size_t)t * myestr[k]];
                    cfac[k] = s_suit[myq[k] * A.maxL + l] / tot * E;     // cons = E * (suit / totalsuit)
                    tp = tp + cfac[k] * B0;
                }
            }
            T consconv(0.0);
            T us1(0.0), us2(0.0);
            if (nq > 0) {   // step 4: overconsumption (R/action_predate.R:381-406)
                T cr = tp / avoid_zero(B0);
                cr = dif_pmax_c(cr, 0.95, -1e3);
                T tpn = B0 * cr;
                consconv = 1.0 / avoid_zero(tp) * tpn;
                num = num * (1.0 - cr);
                if (is_us) { us1 = tp; us2 = tpn; }
            }
            // ================= g3a_naturalmortality (R/action_naturalmortality.R:26)
            if (mort_slot >= 0) {
                if (mort_calc != steplen) { mortf = xexp(-slotv[mort_slot] * dt); mort_calc = steplen; }
                num = num * mortf;
            }
            // ================= publish state for the banded growth gather
            const bool trans_now = grow_n >= 0 && n_out > 0 && ((trans_mask >> step) & 1);
            if (has_cell) {
                s_wgt[tid] = wgt;
                if (trans_now && mat_kind == G3MAT_CONSTANT) {
                    // g3a_mature_constant ratio of the SOURCE cell (R/action_mature.R:44-74, R/action_grow.R:438-457)
                    const int32_t* ms = I + St[G3S_MAT_OFF];
                    T inner(0.0);
                    if (ms[0] >= 0) inner = inner - slotv[ms[0]] * (midlen - slotv[ms[1]]);
                    if (ms[2] >= 0) inner = inner - slotv[ms[2]] * ((double)(St[G3S_MINAGE] + aidx) - slotv[ms[3]]);
                    T ratio = 1.0 / (1.0 + xexp(inner));
                    s_num[tid] = num * (1.0 - ratio);
                    s_alt[tid] = num * ratio;
                } else {
                    s_num[tid] = num;
                }
            }
            if (have_us) {
                T a = warp_sum(us1), c2 = warp_sum(us2);
                if (lane == 0) { s_red[(NRED + MAXTS) * 32 + warp] = a; s_red[(NRED + MAXTS + 1) * 32 + warp] = c2; }
            }
            // growth tables depend on the step length only (R/action_grow.R:400-404)
            const bool recalc = lastcalc != steplen;
            __syncthreads();
            // ================= g3l_understocking (R/likelihood_understocking.R:36-46); thread 0 keeps nll
            if (have_us && tid == 0) {
                const T* r1 = s_red + (NRED + MAXTS) * 32; const T* r2 = r1 + 32;
                T o1 = r1[0], o2 = r2[0];
                for (int w = 1; w < nwarp; w++) { o1 = o1 + r1[w]; o2 = o2 + r2[w]; }
                T tot = o1 - o2;
                T u = (us_power == 2.0) ? tot * tot : xpow(tot, T(us_power));
                nll = nll + us_weight * u;
                comp_us = comp_us + u;
                if (A.report) A.report[1 + (size_t)(A.NNLL - 2) * NT + t] = val(u);
            }
            if (recalc) {
                // ---- growth_bbinom in product form + maturity split, source-oriented scratch
                for (int ss2 = 0; ss2 < NS; ss2++) {
                    const int32_t* S2 = I + I[G3H_STOCK_OFF] + ss2 * G3S_LEN;
                    const int n2 = S2[G3S_GROW_N];
                    if (n2 < 0) continue;
                    const int L2 = S2[G3S_L];
                    const int32_t* gs = I + S2[G3S_GROW_OFF];
                    const bool cont = S2[G3S_MAT_KIND] == G3MAT_CONTINUOUS && S2[G3S_N_OUT] > 0;
                    T* scr = s_gscr;                 // [3][(n2+1)*L2] source-oriented G, G*ratio, G*(1-ratio)
                    const int gsz = (n2 + 1) * L2;
                    if (tid < L2) {
                        const double ml = R[S2[G3S_MIDLEN_ROFF] + tid], pdl = R[S2[G3S_PLUSDL_ROFF]];
                        // g3a_grow_lengthvbsimple (R/action_grow.R:54) inside the bbinom wrapper (R/action_grow.R:220)
                        T dl = avoid_zero((slotv[gs[0]] - ml) * (1.0 - xexp(-(slotv[gs[1]]) * dt)));
                        T delta = avoid_zero(dl / pdl);
                        T beta = avoid_zero(slotv[gs[2]]);
                        T alpha = (beta * delta) / ((double)n2 - delta);
                        // P(x) = C(n,x) prod_{i<x}|alpha+i| prod_{j<n-x}(beta+j) / prod_{k<n}|alpha+beta+k|
                        // (the reference's lgamma is log|Gamma|, so a mean jump beyond n -- alpha < 0 --
                        // yields absolute values, tests/test-action_grow.R:34-62; beta > 0 always)
                        T g = T(1.0);
                        for (int j = 0; j < n2; j++) g = g * ((beta + (double)j) / xabs(alpha + beta + (double)j));
                        T m0(0.0), malpha(0.0);
                        if (cont) {
                            const int32_t* ms = I + S2[G3S_MAT_OFF];
                            malpha = slotv[ms[0]];
                            m0 = 1.0 / (1.0 + xexp(0.0 - malpha * (ml - slotv[ms[1]])));   // R/action_mature.R:46-72
                        }
                        for (int x = 0; x <= n2; x++) {
                            scr[x * L2 + tid] = g;
                            if (cont) {
                                // g3a_mature_continuous (R/action_mature.R:1-42), beta = 0
                                T ratio = dif_pminmax_c(m0 * (malpha * (pdl * x)), 0.0, 1.0, 1e5);
                                scr[gsz + x * L2 + tid] = g * ratio;
                                scr[2 * gsz + x * L2 + tid] = g * (1.0 - ratio);
                            }
                            if (x < n2)
                                g = g * ((double)(n2 - x) / (double)(x + 1)) * (xabs(alpha + (double)x) / (beta + (double)(n2 - x - 1)));
                        }
                    }
                    __syncthreads();
                    // ---- destination-oriented band with the plus-group tail sums folded in
                    // (g3a_grow_matrix_len, R/action_grow.R:233-271)
                    for (int i = tid; i < gsz; i += blockDim.x) {
                        int d = i / L2, j = i - d * L2, src = j - d;
                        for (int v = 0; v < (cont ? 3 : 1); v++) {
                            T g(0.0);
                            if (src >= 0) {
                                if (j < L2 - 1) g = scr[v * gsz + d * L2 + src];
                                else for (int x = d; x <= n2; x++) g = g + scr[v * gsz + x * L2 + src];
                            }
                            int o = v == 0 ? A.o_g[ss2] : (v == 1 ? A.o_gr[ss2] : A.o_gn[ss2]);
                            sm[o + i] = g;
                        }
                    }
                    __syncthreads();
                }
                lastcalc = steplen;
            }

            // ================= g3a_growmature (R/action_grow.R:340-497): banded gather
            if (grow_n >= 0) {
                const T wa = slotv[I[St[G3S_GROW_OFF] + 3]];
                const T pwl = my_pw[l];
                const int dmax = l < grow_n ? l : grow_n;
                if (trans_now && mat_kind == G3MAT_CONTINUOUS) {
                    T an(0.0), aw(0.0), bn(0.0), bw(0.0);
                    for (int d = 0; d <= dmax; d++) {
                        T ns = s_num[tid - d];
                        T w = (wa * (pwl - my_pw[l - d]) + s_wgt[tid - d]) * ns;
                        T gr = my_gr[d * L + l], gn = my_gn[d * L + l];
                        an = an + gr * ns; aw = aw + gr * w;
                        bn = bn + gn * ns; bw = bw + gn * w;
                    }
                    tn = an; tw = aw / avoid_zero(an);
                    num = bn; wgt = bw / avoid_zero(bn);
                } else if (trans_now && mat_kind == G3MAT_CONSTANT) {
                    T an(0.0), aw(0.0), bn(0.0), bw(0.0);
                    for (int d = 0; d <= dmax; d++) {
                        T g = my_g[d * L + l];
                        T wsrc = wa * (pwl - my_pw[l - d]) + s_wgt[tid - d];
                        T n1 = s_alt[tid - d] * g, n2 = s_num[tid - d] * g;
                        an = an + n1; aw = aw + wsrc * n1;
                        bn = bn + n2; bw = bw + wsrc * n2;
                    }
                    tn = an; tw = aw / avoid_zero(an);
                    num = bn; wgt = bw / avoid_zero(bn);
                } else {
                    T bn(0.0), bw(0.0);
                    for (int d = 0; d <= dmax; d++) {
                        T n2 = s_num[tid - d] * my_g[d * L + l];
                        bn = bn + n2;
                        bw = bw + (wa * (pwl - my_pw[l - d]) + s_wgt[tid - d]) * n2;
                    }
                    num = bn; wgt = bw / avoid_zero(bn);
                }
            }
            // ================= g3a_step_transition (R/action_mature.R:78-134)
            if (A.trans_any[step]) {
                if (trans_now) { s_tn[tid] = tn; s_tw[tid] = tw; }
                __syncthreads();
                if (my_inc_src >= 0 && ((my_inc_mask >> step) & 1)) {
                    T moved = s_tn[my_inc_src] * slotv[my_inc_slot];
                    wgt = ratio_add_pop(wgt, num, s_tw[my_inc_src], moved);
                    num = num + moved;
                }
                if (trans_now) {      // unclaimed remainder moves back (move_remainder = TRUE)
                    T rem = tn;
                    for (int k = my_rem_off; A.rem_slots[k] >= 0; k++) rem = rem - tn * slotv[A.rem_slots[k]];
                    num = num + rem;
                }
            }
            // ================= g3a_renewal (R/action_renewal.R:166-188)
#pragma unroll
            for (int k = 0; k < MAXRN; k++) {
                if (k < n_renew) {
                    const int32_t* Rn = I + St[G3S_RENEW_OFF] + k * G3R_LEN;
                    if (aidx == Rn[G3R_AGE_IDX] && (Rn[G3R_RUN_STEP] == 0 || Rn[G3R_RUN_STEP] == step + 1)) {
                        int fs = I[Rn[G3R_FACTOR_OFF] + yidx * narea + ridx];
                        if (fs >= 0) {
                            T rn = rnorm[k] * slotv[fs];
                            T rw = s_rwt[A.o_rd[s] - A.o_rd[0] + k * L + l];
                            wgt = ratio_add_pop(wgt, num, rw, rn);
                            num = num + rn;
                        }
                    }
                }
            }

            // ================= g3l_distribution (R/likelihood_distribution.R:275-394)
            T xv[MAXX];
            bool any_gather = false;
#pragma unroll
            for (int x = 0; x < MAXX; x++) {
                xv[x] = T(0.0);
                if (x < A.NX) {
                    const int32_t* X = I + I[G3H_XBUF_OFF] + x * G3X_LEN;
                    if (I[X[G3X_ACTIVE_OFF] + t]) {
                        if (has_cell) {
                            if (X[G3X_KIND] == 0) {     // catch: sum of cons over the component's predators
                                T f(0.0);
#pragma unroll
                                for (int k = 0; k < MAXQ; k++)
                                    if (myq[k] >= 0)
                                        for (int j = 0; j < X[G3X_N_PP]; j++)
                                            if (I[X[G3X_PP_OFF] + j] == myq[k]) f = f + cfac[k];
                                T cons = f * B0 * consconv;
                                xv[x] = X[G3X_UNIT] == 0 ? cons / avoid_zero(wgt) : cons;
                            } else {
                                xv[x] = X[G3X_UNIT] == 0 ? num : num * wgt;
                            }
                            if (A.o_xslot[x] >= 0) sm[A.o_xslot[x] + tid] = xv[x];
                        }
                        if (A.o_xslot[x] >= 0) any_gather = true;
                    }
                }
            }
            if (any_gather) __syncthreads();
            for (int c = 0; c < A.NCOMP; c++) {
                const int32_t* C = I + I[G3H_COMP_OFF] + c * G3C_LEN;
                const int ti = I[C[G3C_TIDX_OFF] + t];
                if (ti < 0 || !(spar[C[G3C_WEIGHT_PAR]] > 0.0)) continue;
                const int func = C[G3C_FUNC], nd = C[G3C_N_DEST], nt = C[G3C_N_TIMES];
                const bool si = func == G3F_SI_LOG || func == G3F_SI_LINEAR;
                T* ms = sm + A.o_ms[c] + (si ? ti * nd : 0);
                T m(0.0);
                if (C[G3C_MODE] == 0) {
                    if (tid < nd) {
                        const int32_t* ptr = I + C[G3C_CSR_PTR_OFF];
                        const int32_t* idx = I + C[G3C_CSR_IDX_OFF];
                        const double* cf = R + C[G3C_CSR_COEF_ROFF];
                        const T* xs = sm + A.o_xslot[C[G3C_XBUF]];
                        T acc = ms[tid];
                        for (int j = ptr[tid]; j < ptr[tid + 1]; j++) acc = acc + cf[j] * xs[idx[j]];
                        ms[tid] = acc;
                        m = acc;
                    }
                } else {
                    // few destinations, every cell contributes: block-reduce per destination
                    const int32_t* cd = I + C[G3C_CELL_DEST_OFF];
                    const double* cf = R + C[G3C_CELL_COEF_ROFF];
                    for (int e = 0; e < nd; e++) {
                        T v(0.0);
                        if (has_cell && cd[tid] == e) v = cf[tid] * xv[C[G3C_XBUF]];
                        T tot = block_sum(v, s_red, rrow);
                        if (tid == e) { m = ms[e] + tot; ms[e] = m; }
                    }
                }
                if (!I[C[G3C_CMP_OFF] + t]) continue;
                const T weight = mkpar<T>(spar, seed, C[G3C_WEIGHT_PAR]);
                const int nb = C[G3C_N_BINS], nsl = C[G3C_N_SLICES], nag = C[G3C_N_AREAG];
                T cnll(0.0);
                bool scored = false;
                if (func == G3F_SUMOFSQUARES) {     // R/likelihood_distribution.R:3-39
                    T atot(0.0);
                    for (int ag = 0; ag < nag; ag++) {
                        bool mine = tid < nd && tid / (nsl * nb) == ag;
                        T tot = block_sum(mine ? m : T(0.0), s_red, rrow);
                        if (mine) atot = avoid_zero(tot);
                    }
                    T d2(0.0);
                    if (tid < nd) {
                        T dlt = m / atot - A.obsprop[A.obs_off[c] + (size_t)ti * nd + tid];
                        d2 = dlt * dlt;
                    }
                    cnll = block_sum(d2, s_red, rrow);
                    scored = true;
                } else if (func == G3F_MULTINOMIAL) {   // R/likelihood_distribution.R:41-60
                    __syncthreads();
                    T term(0.0);
                    if (tid < nd) {
                        const T* sl = ms + (tid / nb) * nb;
                        T tot(0.0);
                        for (int k = 0; k < nb; k++) tot = tot + sl[k];
                        double o = A.obsprop[A.obs_off[c] + (size_t)ti * nd + tid];
                        double eps = R[C[G3C_PARAM_ROFF]];
                        term = o * xlog(dif_pmax_c(m / avoid_zero(tot), 1.0 / (nb * eps), 1e4));
                    }
                    T likely = block_sum(term, s_red, rrow);
                    cnll = 2.0 * (-likely + A.obsconst[A.obsconst_off[c] + ti]);
                    scored = true;
                } else if (ti == nt - 1) {              // surveyindices: R/likelihood_distribution.R:82-125
                    __syncthreads();
                    const double fa = R[C[G3C_PARAM_ROFF]], fb = R[C[G3C_PARAM_ROFF] + 1];
                    const T* series = sm + A.o_ms[c];
                    for (int e = 0; e < nd; e++) {
                        T N(0.0); double Iv = 0.0;
                        if (tid < nt) {
                            T mv = series[tid * nd + e];
                            N = func == G3F_SI_LOG ? xlog(avoid_zero(mv)) : mv;
                            Iv = A.obsprop[A.obs_off[c] + (size_t)tid * nd + e];
                        }
                        T mN = block_sum(N, s_red, rrow) / (double)nt;
                        T mI = block_sum(T(Iv), s_red, rrow) / (double)nt;
                        T beta(fb), alpha(fa);
                        if (isnan(fb)) {
                            T sxy = block_sum(tid < nt ? (Iv - mI) * (N - mN) : T(0.0), s_red, rrow);
                            T sxx = block_sum(tid < nt ? (N - mN) * (N - mN) : T(0.0), s_red, rrow);
                            beta = sxy / avoid_zero(sxx);
                        }
                        if (isnan(fa)) alpha = mI - beta * mN;
                        T r2(0.0);
                        if (tid < nt) { T r = alpha + beta * N - Iv; r2 = r * r; }
                        cnll = cnll + block_sum(r2, s_red, rrow);
                        if (A.report && tid == 0) {
                            A.report[A.rep_params + 2 * c] = val(alpha);
                            A.report[A.rep_params + 2 * c + 1] = val(beta);
                        }
                    }
                    scored = true;
                }
                if (A.report) {
                    if (tid < nd) A.report[A.rep_model[c] + (size_t)ti * nd + tid] = val(ms[tid]);
                    if (tid == 0 && scored) A.report[1 + (size_t)c * NT + t] = val(cnll);
                }
                if (scored) {
                    nll = nll + weight * cnll;
#pragma unroll
                    for (int k = 0; k < MAXC; k++) if (k == c) comp_tot[k] = comp_tot[k] + cnll;
                }
                if (!si && tid < nd) ms[tid] = T(0.0);     // zero counters for the next period
            }

            // ================= g3a_age (R/action_age.R:51-85) + movement hand-off
            if (final_step) {
                __syncthreads();
                if (has_cell) { s_num[tid] = num; s_wgt[tid] = wgt; }
                __syncthreads();
                if (age_enable) {
                    if (nage == 1) {
                        if (age_n_out > 0) num = T(0.0);
                    } else if (aidx == nage - 1) {
                        if (age_n_out > 0) { num = s_num[tid - L]; wgt = s_wgt[tid - L]; }
                        else {
                            wgt = ratio_add_pop(wgt, num, s_wgt[tid - L], s_num[tid - L]);
                            num = num + s_num[tid - L];
                        }
                    } else if (aidx == 0) {
                        num = T(0.0);
                    } else {
                        num = s_num[tid - L]; wgt = s_wgt[tid - L];
                    }
                }
                if (my_age_src >= 0) {
                    T moved = s_num[my_age_src] * slotv[my_age_slot];
                    wgt = ratio_add_pop(wgt, num, s_wgt[my_age_src], moved);
                    num = num + moved;
                }
            }
        }   // time loop

        if (tid == 0) {
            T out = bad_time ? T(nan("")) : nll;
            if (tile == 0) {
                A.nll[b] = val(out);
                if (A.comp) {
                    for (int c = 0; c < A.NCOMP; c++) A.comp[b * A.NNLL + c] = val(comp_tot[c]);
                    A.comp[b * A.NNLL + A.NNLL - 2] = val(comp_us);
                    A.comp[b * A.NNLL + A.NNLL - 1] = val(comp_bounds);
                }
                if (A.report) A.report[0] = val(out);
            }
            if (NTAN > 0 && A.grad)
                for (int k = 0; k < NTAN; k++)
                    if (tile * NTAN + k < A.K) A.grad[b * A.K + tile * NTAN + k] = tan_of(out, k);
        }
    }
}

}  // namespace g3
