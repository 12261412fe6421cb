// g3_warp_kernel.cuh -- one WARP per parameter vector: the production evaluation kernel.
//
// Mapping. A lane is a length group: the 32 lanes of a warp hold G = floor(32 / L) whole
// (length x one age x one area) columns of a stock at a time ("tile"), and the warp walks the
// tiles of every stock, then the time steps, for one parameter vector. The stock state
// (num/wgt over length x age x area, all stocks) lives in that warp's slice of shared memory
// for all time steps. Consequences:
//   * no block-level barrier anywhere -- warps are independent, a CTA is just WPC of them;
//   * every branch on model structure (which stock, does it mature, is there a catch this step)
//     is warp-uniform; lanes only differ by data;
//   * neighbour access (banded growth gather, ageing, maturation hand-over) is a shared-memory
//     read of the same warp's slice, ordered by __syncwarp; reductions are warp shuffles;
//   * the per-step bookkeeping is paid once per tile, not once per cell.
// Arithmetic and action order are those of the reference's generated objective
// (baseline/multiple-substocks.cpp:692-2161); see g3_kernel.cuh for what is hoisted/restructured.
// Exact shortcuts on top of those: predation is skipped on steps where every landing is zero
// (cons = 0, consratio = 0, num unchanged -- R/action_predate.R:335-406 with E = 0), and
// consconv is only formed on steps where a catch likelihood consumes it.
#pragma once
#include "g3_kernel.cuh"

namespace g3 {

template <class T> __device__ __forceinline__ T shfl_xor_t(const T& x, int o);
template <> __device__ __forceinline__ double shfl_xor_t<double>(const double& x, int o) { return __shfl_xor_sync(0xffffffffu, x, o); }
template <int N> __device__ __forceinline__ Dual<N> shfl_xor_d(const Dual<N>& x, int o) {
    Dual<N> r; r.v = __shfl_xor_sync(0xffffffffu, x.v, o);
#pragma unroll
    for (int i = 0; i < N; i++) r.d[i] = __shfl_xor_sync(0xffffffffu, x.d[i], o);
    return r;
}
__device__ __forceinline__ double warp_sum_all(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
template <int N> __device__ __forceinline__ Dual<N> warp_sum_all(Dual<N> x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = x + shfl_xor_d(x, o);
    return x;
}

constexpr int WTAP = 5;     // growth taps kept in registers (maxlengthgroupgrowth <= 4); wider bands read shared memory

template <class T, int WPC>
__global__ void __launch_bounds__(WPC * 32) eval_warp_kernel(const __grid_constant__ KArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int32_t* __restrict__ I = A.I;
    const double* __restrict__ R = A.R;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    unsigned char* wbase = smem_raw + (size_t)wib * A.w_warp_bytes;
    double* spar = reinterpret_cast<double*>(wbase);
    T* sm = reinterpret_cast<T*>(wbase + A.sm_par_bytes);
    T* slotv = sm + A.o_slot;
    T* s_num = sm + A.o_num;
    T* s_wgt = sm + A.o_wgt;
    T* s_tn = sm + A.o_tn;          // compact: maturing stocks only (w_tr_off)
    T* s_tw = sm + A.o_tw;
    T* s_alt = sm + A.o_alt;
    T* s_suit = sm + A.o_suit;
    T* s_scr = sm + A.o_gscr;
    T* s_mort = sm + A.w_o_mort;
    T* s_tot = sm + A.w_o_tot;      // [NTS] E / total suitable biomass ; [NNLL] component totals after it
    T* s_ctot = s_tot + MAXTS;
    T* s_mv = sm + A.w_o_mv;        // ageing movement stash (num then wgt)

    const int NC = A.NC, NS = A.NS, NSTEPS = A.NSTEPS, NT = A.NT;
    constexpr int NTAN = Tan<T>::n;
    const int ntiles = NTAN > 0 ? (A.K + NTAN - 1) / NTAN : 1;
    const double us_power = R[I[G3H_US_ROFF]], us_weight = R[I[G3H_US_ROFF] + 1];
    const bool have_us = I[G3H_US_N] > 0;
    const long long nwarps = (long long)gridDim.x * WPC;

    for (long long work = (long long)blockIdx.x * WPC + wib; work < A.B * ntiles; work += nwarps) {
        const long long b = work / ntiles;
        const int tile_t = (int)(work - b * ntiles);
        int seed[NTAN > 0 ? NTAN : 1];
#pragma unroll
        for (int k = 0; k < (NTAN > 0 ? NTAN : 1); k++) {
            int g = tile_t * NTAN + k;
            seed[k] = (NTAN > 0 && g < A.K) ? A.tangent_idx[g] : -1;
        }
        __syncwarp();
        for (int i = lane; i < A.NPAR; i += 32) spar[i] = A.par[b * A.NPAR + i];
        __syncwarp();

        // ---- g3a_time (R/action_time.R:75-91)
        const double retro = I[G3H_PAR_RETRO] >= 0 ? spar[I[G3H_PAR_RETRO]] : 0.0;
        const double proj = I[G3H_PAR_PROJECT] >= 0 ? spar[I[G3H_PAR_PROJECT]] : 0.0;
        const int total_steps = NSTEPS * (I[G3H_END_YEAR] - (int)floor(retro) - I[G3H_START_YEAR] + (int)floor(proj)) + NSTEPS - 1;
        const bool bad_time = !(retro >= 0.0) || !(proj >= 0.0) || total_steps + 1 > NT;

        for (int i = lane; i < A.NSLOT; i += 32) slotv[i] = eval_slot<T>(I, R, spar, seed, i);
        __syncwarp();

        // ================= per-evaluation tables
        for (int i = lane; i < A.NPP * A.maxL; i += 32) {       // suitability (R/suitability.R:5-13)
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
        for (int s = 0; s < NS; s++) {
            const int32_t* St = I + I[G3H_STOCK_OFF] + s * G3S_LEN;
            const int L = St[G3S_L], ncol = St[G3S_NAGE] * St[G3S_NAREA];
            const double ml = lane < L ? R[St[G3S_MIDLEN_ROFF] + lane] : 1.0;
            const double pdl = R[St[G3S_PLUSDL_ROFF]];
            // natural mortality factor per column and distinct step length (R/action_naturalmortality.R:26)
            if (St[G3S_MORT_OFF] >= 0)
                for (int i = lane; i < ncol * A.w_ndt; i += 32) {
                    int idt = i / ncol, col = i - idt * ncol;
                    s_mort[A.w_mort_off[s] + i] = xexp(-slotv[I[St[G3S_MORT_OFF] + col]] * (A.w_dt_len[idt] / 12.0));
                }
            // renewal length distribution, normalised, and weight (R/action_renewal.R:56-80)
            for (int k = 0; k < St[G3S_N_RENEW]; k++) {
                const int32_t* Rn = I + St[G3S_RENEW_OFF] + k * G3R_LEN;
                T d(0.0);
                if (lane < L) { d = dnorm(ml, slotv[Rn[G3R_MEAN]], avoid_zero(slotv[Rn[G3R_SD]])); s_scr[lane] = d; }
                __syncwarp();
                if (lane < L) {
                    T tot(0.0);
                    for (int j = 0; j < L; j++) tot = tot + s_scr[j];
                    sm[A.o_rd[s] + k * L + lane] = d / avoid_zero(tot) * 10000.0;
                    sm[A.w_o_rw[s] + k * L + lane] = slotv[Rn[G3R_WALPHA]] * xpow(ml, slotv[Rn[G3R_WBETA]]);
                }
                __syncwarp();
            }
            const int n2 = St[G3S_GROW_N];
            if (n2 < 0) continue;
            const int32_t* gs = I + St[G3S_GROW_OFF];
            const bool cont = St[G3S_MAT_KIND] == G3MAT_CONTINUOUS && St[G3S_N_OUT] > 0;
            const int gsz = (n2 + 1) * L;
            T pwl(0.0);
            if (lane < L) { pwl = xpow(ml, slotv[gs[4]]); sm[A.o_pws[s] + lane] = pwl; }
            __syncwarp();
            // weight gain of a fish arriving in l from l - d (R/action_grow.R:58-71): walpha (l^wbeta - (l-d)^wbeta)
            for (int i = lane; i < gsz; i += 32) {
                int d = i / L, j = i - d * L;
                sm[A.w_o_dw[s] + i] = j - d >= 0 ? slotv[gs[3]] * (sm[A.o_pws[s] + j] - sm[A.o_pws[s] + j - d]) : T(0.0);
            }
            for (int idt = 0; idt < A.w_ndt; idt++) {
                const double dt = A.w_dt_len[idt] / 12.0;
                T* scr = s_scr;                 // [3][(n+1) L] source-oriented G, G ratio, G (1 - ratio)
                if (lane < L) {
                    // g3a_grow_lengthvbsimple (R/action_grow.R:54) inside the bbinom wrapper (R/action_grow.R:220)
                    T dl = avoid_zero((slotv[gs[0]] - ml) * (1.0 - xexp(-(slotv[gs[1]]) * dt)));
                    T delta = avoid_zero(dl / pdl);
                    T beta = avoid_zero(slotv[gs[2]]);
                    T alpha = (beta * delta) / ((double)n2 - delta);
                    // growth_bbinom in product form, see g3_kernel.cuh
                    T g = T(1.0);
                    for (int j = 0; j < n2; j++) g = g * ((beta + (double)j) / xabs(alpha + beta + (double)j));
                    T m0(0.0), malpha(0.0);
                    if (cont) {
                        const int32_t* ms = I + St[G3S_MAT_OFF];
                        malpha = slotv[ms[0]];
                        m0 = 1.0 / (1.0 + xexp(0.0 - malpha * (ml - slotv[ms[1]])));       // R/action_mature.R:46-72
                    }
                    for (int x = 0; x <= n2; x++) {
                        scr[x * L + lane] = g;
                        if (cont) {     // g3a_mature_continuous (R/action_mature.R:1-42), beta = 0
                            T ratio = dif_pminmax_c(m0 * (malpha * (pdl * x)), 0.0, 1.0, 1e5);
                            scr[gsz + x * L + lane] = g * ratio;
                            scr[2 * gsz + x * L + lane] = g * (1.0 - ratio);
                        }
                        if (x < n2)
                            g = g * ((double)(n2 - x) / (double)(x + 1)) * (xabs(alpha + (double)x) / (beta + (double)(n2 - x - 1)));
                    }
                }
                __syncwarp();
                // destination-oriented band, plus-group tail sums folded in (R/action_grow.R:233-271)
                for (int i = lane; i < gsz; i += 32) {
                    int d = i / L, j = i - d * L, src = j - d;
                    for (int v = 0; v < (cont ? 3 : 1); v++) {
                        T g(0.0);
                        if (src >= 0) {
                            if (j < L - 1) g = scr[v * gsz + d * L + src];
                            else for (int x = d; x <= n2; x++) g = g + scr[v * gsz + x * L + src];
                        }
                        int o = v == 0 ? A.o_g[s] : (v == 1 ? A.o_gr[s] : A.o_gn[s]);
                        sm[o + idt * gsz + i] = g;
                    }
                }
                __syncwarp();
            }
        }

        // ================= g3a_initialconditions (R/action_renewal.R:83-163)
        for (int s = 0; s < NS; s++) {
            const int32_t* St = I + I[G3H_STOCK_OFF] + s * G3S_LEN;
            const int L = St[G3S_L], c0 = St[G3S_CELL_OFF], ncell = L * St[G3S_NAGE] * St[G3S_NAREA];
            const int CH = A.w_G[s] * L;
            const int gcol = lane / L, l = lane - gcol * L;
            const double ml = R[St[G3S_MIDLEN_ROFF] + l];
            for (int k = 0; k < A.w_ntile[s]; k++) {
                const int cell = c0 + k * CH + lane;
                const bool valid = lane < CH && cell < c0 + ncell;
                const int col = k * A.w_G[s] + gcol;
                T num(0.0), wgt(1.0), dens(0.0);
                const bool has_init = valid && St[G3S_INIT_OFF] >= 0;
                if (has_init) {
                    const int32_t* sl = I + St[G3S_INIT_OFF] + col * 5;
                    dens = dnorm(ml, slotv[sl[0]], avoid_zero(slotv[sl[1]]));
                    s_scr[lane] = dens;
                }
                __syncwarp();
                if (has_init) {
                    const int32_t* sl = I + St[G3S_INIT_OFF] + col * 5;
                    T tot(0.0);
                    for (int j = 0; j < L; j++) tot = tot + s_scr[gcol * L + j];
                    num = dens / avoid_zero(tot) * 10000.0 * slotv[sl[2]];        // normalize_vec, R/aab_env.R:37-41
                    wgt = slotv[sl[3]] * xpow(ml, slotv[sl[4]]);
                }
                __syncwarp();
                if (valid) { s_num[cell] = num; s_wgt[cell] = wgt; }
            }
        }

        T nll(0.0);
        for (int i = lane; i < A.NNLL; i += 32) s_ctot[i] = T(0.0);
        // ---- g3l_bounds_penalty (R/likelihood_bounds.R:13-16), cur_time == 0
        if (A.NBOUND > 0) {
            T pen(0.0);
            const double bw = R[I[G3H_BOUND_WS_ROFF]], bs = R[I[G3H_BOUND_WS_ROFF] + 1];
            for (int i = lane; i < A.NBOUND; i += 32) {
                T p = mkpar<T>(spar, seed, I[I[G3H_BOUND_OFF] + i]);
                double lo = R[I[G3H_BOUND_ROFF] + 2 * i], up = R[I[G3H_BOUND_ROFF] + 2 * i + 1];
                T a = lsa_c(bs * (p - up) / (up - lo), 0.0) + lsa_c(bs * (lo - p) / (up - lo), 0.0);
                pen = pen + a * a;
            }
            pen = warp_sum_all(pen);
            nll = nll + bw * pen;
            __syncwarp();
            if (lane == 0) s_ctot[A.NNLL - 1] = pen;
        }
        for (int c = 0; c < A.NCOMP; c++) {         // zero the likelihood slabs
            const int32_t* C = I + I[G3H_COMP_OFF] + c * G3C_LEN;
            bool si = C[G3C_FUNC] == G3F_SI_LOG || C[G3C_FUNC] == G3F_SI_LINEAR;
            int n = C[G3C_N_DEST] * (si ? C[G3C_N_TIMES] : 1);
            for (int i = lane; i < n; i += 32) sm[A.o_ms[c] + i] = T(0.0);
        }
        __syncwarp();

        for (int t = 0; t <= total_steps && !bad_time; t++) {
            const int step = t % NSTEPS, yidx = t / NSTEPS;
            const int idt = A.w_step_idt[step];
            const bool final_step = step == NSTEPS - 1;
            const bool pred_now = A.w_pred_active[t] != 0;
            // which collect vectors are wanted at this step
            bool xact[MAXX];
            bool any_catch_x = false, any_x = false;
#pragma unroll
            for (int x = 0; x < MAXX; x++) {
                xact[x] = false;
                if (x < A.NX) {
                    const int32_t* X = I + I[G3H_XBUF_OFF] + x * G3X_LEN;
                    xact[x] = I[X[G3X_ACTIVE_OFF] + t] != 0;
                    any_x = any_x || xact[x];
                    if (xact[x] && X[G3X_KIND] == 0) any_catch_x = true;
                }
            }

            // ================= g3a_predate step 1: suitable biomass per (predator, area) (R/action_predate.R:314-333)
            if (pred_now) {
                T part[MAXTS];
#pragma unroll
                for (int k = 0; k < MAXTS; k++) part[k] = T(0.0);
                for (int s = 0; s < NS; s++) {
                    const int nq = A.stock_pp_off[s + 1] - A.stock_pp_off[s];
                    if (nq == 0) continue;
                    const int32_t* St = I + I[G3H_STOCK_OFF] + s * G3S_LEN;
                    const int L = St[G3S_L], nage = St[G3S_NAGE], c0 = St[G3S_CELL_OFF], ncell = L * nage * St[G3S_NAREA];
                    const int CH = A.w_G[s] * L;
                    const int gcol = lane / L, l = lane - gcol * L;
                    for (int k = 0; k < A.w_ntile[s]; k++) {
                        const int cell = c0 + k * CH + lane;
                        if (!(lane < CH && cell < c0 + ncell)) continue;
                        const int ridx = (k * A.w_G[s] + gcol) / nage;
                        const T B0 = s_num[cell] * s_wgt[cell];
                        for (int j = 0; j < nq; j++) {
                            const int q = A.stock_pp[A.stock_pp_off[s] + j];
                            const int ts = A.pp_tsid[q * A.maxnarea + ridx];
                            if (ts < 0) continue;
                            T sv = s_suit[q * A.maxL + l] * B0;
#pragma unroll
                            for (int m = 0; m < MAXTS; m++) if (m == ts) part[m] = part[m] + sv;
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < MAXTS; k++)
                    if (k < A.NTS) {
                        T tot = warp_sum_all(part[k]);
                        // landings over total suitable biomass: cons = E * (suit / totalsuit)
                        if (lane == 0) s_tot[k] = R[A.w_ts_eoff[k] + (size_t)t * A.w_ts_estr[k]] / tot;
                    }
                __syncwarp();
            }

            // ================= pass A, stock by stock: predation, natural mortality, growth
            T us1(0.0), us2(0.0);
            for (int s = 0; s < NS; s++) {
                const int32_t* St = I + I[G3H_STOCK_OFF] + s * G3S_LEN;
                const int L = St[G3S_L], nage = St[G3S_NAGE], c0 = St[G3S_CELL_OFF], ncell = L * nage * St[G3S_NAREA];
                const int G = A.w_G[s], CH = G * L;
                const int gcol = lane / L, l = lane - gcol * L;
                const int nq = A.stock_pp_off[s + 1] - A.stock_pp_off[s];
                const bool prey_now = pred_now && nq > 0;
                const bool is_us = A.us_flag[s] != 0;
                const int grow_n = St[G3S_GROW_N], mat_kind = St[G3S_MAT_KIND];
                const bool trans_now = grow_n >= 0 && St[G3S_N_OUT] > 0 && ((St[G3S_TRANS_MASK] >> step) & 1);
                const bool has_mort = St[G3S_MORT_OFF] >= 0;
                if (!prey_now && !has_mort && grow_n < 0) {
                    if (any_catch_x && nq > 0) {        // no catch this step: cons = 0
                        for (int k = 0; k < A.w_ntile[s]; k++) {
                            const int cell = c0 + k * CH + lane;
                            if (lane < CH && cell < c0 + ncell) {
#pragma unroll
                                for (int x = 0; x < MAXX; x++) if (xact[x] && A.w_x_kind[x] == 0) sm[A.o_xslot[x] + cell] = T(0.0);
                            }
                        }
                    }
                    continue;
                }
                // per-lane tables of this stock (l is the same in every tile)
                T sq[MAXQ];
#pragma unroll
                for (int j = 0; j < MAXQ; j++) sq[j] = (prey_now && j < nq && l < L) ? s_suit[A.stock_pp[A.stock_pp_off[s] + j] * A.maxL + l] : T(0.0);
                const int dmax = grow_n >= 0 ? (l < grow_n ? l : grow_n) : -1;
                const bool regtap = grow_n >= 0 && grow_n < WTAP;
                const int gsz = (grow_n + 1) * L;
                const bool cont_now = trans_now && mat_kind == G3MAT_CONTINUOUS;
                const bool const_now = trans_now && mat_kind == G3MAT_CONSTANT;
                const T* tg = sm + (cont_now ? A.o_gn[s] : A.o_g[s]) + idt * gsz;     // staying part
                const T* tgr = sm + A.o_gr[s] + idt * gsz;                              // maturing part (continuous)
                const T* tdw = sm + A.w_o_dw[s];
                T gt[WTAP], grt[WTAP], dwt[WTAP];
#pragma unroll
                for (int d = 0; d < WTAP; d++) {
                    gt[d] = grt[d] = dwt[d] = T(0.0);
                    if (regtap && d <= dmax && lane < CH) {
                        gt[d] = tg[d * L + l];
                        dwt[d] = tdw[d * L + l];
                        if (cont_now) grt[d] = tgr[d * L + l];
                    }
                }
                const int32_t* msl = I + St[G3S_MAT_OFF];

                for (int k = 0; k < A.w_ntile[s]; k++) {
                    const int cell = c0 + k * CH + lane;
                    const bool valid = lane < CH && cell < c0 + ncell;
                    const int col = k * G + gcol;
                    const int ridx = col / nage, aidx = col - ridx * nage;
                    T num(0.0), wgt(1.0);
                    if (valid) {
                        num = s_num[cell]; wgt = s_wgt[cell];
                        if (prey_now) {     // R/action_predate.R:335-406
                            const T B0 = num * wgt;
                            T tp(0.0), fx[MAXX];
#pragma unroll
                            for (int x = 0; x < MAXX; x++) fx[x] = T(0.0);
#pragma unroll
                            for (int j = 0; j < MAXQ; j++)
                                if (j < nq) {
                                    const int q = A.stock_pp[A.stock_pp_off[s] + j];
                                    const int ts = A.pp_tsid[q * A.maxnarea + ridx];
                                    if (ts >= 0) {
                                        T cf = sq[j] * s_tot[ts];
                                        tp = tp + cf * B0;
#pragma unroll
                                        for (int x = 0; x < MAXX; x++) if (xact[x] && ((A.w_x_ppmask[x] >> q) & 1)) fx[x] = fx[x] + cf;
                                    }
                                }
                            T cr = tp / avoid_zero(B0);
                            cr = dif_pmax_c(cr, 0.95, -1e3);
                            const T tpn = B0 * cr;
                            num = num * (1.0 - cr);
                            if (is_us) { us1 = us1 + tp; us2 = us2 + tpn; }
                            if (any_catch_x) {
                                const T cc = 1.0 / avoid_zero(tp) * tpn * B0;      // consconv * B0
#pragma unroll
                                for (int x = 0; x < MAXX; x++) if (xact[x] && A.w_x_kind[x] == 0) sm[A.o_xslot[x] + cell] = fx[x] * cc;
                            }
                        } else if (any_catch_x && nq > 0) {
#pragma unroll
                            for (int x = 0; x < MAXX; x++) if (xact[x] && A.w_x_kind[x] == 0) sm[A.o_xslot[x] + cell] = T(0.0);
                        }
                        if (has_mort) num = num * s_mort[A.w_mort_off[s] + idt * (nage * St[G3S_NAREA]) + col];
                        if (grow_n >= 0) {
                            if (const_now) {
                                // g3a_mature_constant ratio of the SOURCE cell (R/action_mature.R:44-74, R/action_grow.R:438-457)
                                T inner(0.0);
                                if (msl[0] >= 0) inner = inner - slotv[msl[0]] * (R[St[G3S_MIDLEN_ROFF] + l] - slotv[msl[1]]);
                                if (msl[2] >= 0) inner = inner - slotv[msl[2]] * ((double)(St[G3S_MINAGE] + aidx) - slotv[msl[3]]);
                                T ratio = 1.0 / (1.0 + xexp(inner));
                                s_num[cell] = num * (1.0 - ratio);
                                s_alt[cell] = num * ratio;
                            } else {
                                s_num[cell] = num;
                            }
                        } else {
                            s_num[cell] = num;
                        }
                    }
                    if (grow_n < 0) continue;
                    __syncwarp();
                    // ---- g3a_growmature (R/action_grow.R:340-497): banded gather over the source lengths l - d
                    T tn(0.0), tw(0.0);
                    if (valid) {
                        if (cont_now) {
                            T an(0.0), aw(0.0), bn(0.0), bw(0.0);
                            if (regtap) {
#pragma unroll
                                for (int d = 0; d < WTAP; d++)
                                    if (d <= dmax) {
                                        T ns = s_num[cell - d];
                                        T w = (dwt[d] + s_wgt[cell - d]) * ns;
                                        an = an + grt[d] * ns; aw = aw + grt[d] * w;
                                        bn = bn + gt[d] * ns; bw = bw + gt[d] * w;
                                    }
                            } else {
                                for (int d = 0; d <= dmax; d++) {
                                    T ns = s_num[cell - d];
                                    T w = (tdw[d * L + l] + s_wgt[cell - d]) * ns;
                                    T gr = tgr[d * L + l], gn = tg[d * L + l];
                                    an = an + gr * ns; aw = aw + gr * w;
                                    bn = bn + gn * ns; bw = bw + gn * w;
                                }
                            }
                            tn = an; tw = aw / avoid_zero(an);
                            num = bn; wgt = bw / avoid_zero(bn);
                        } else if (const_now) {
                            T an(0.0), aw(0.0), bn(0.0), bw(0.0);
                            for (int d = 0; d <= dmax; d++) {
                                T g = tg[d * L + l];
                                T wsrc = tdw[d * L + l] + s_wgt[cell - d];
                                T n1 = s_alt[cell - d] * g, n2 = s_num[cell - d] * g;
                                an = an + n1; aw = aw + wsrc * n1;
                                bn = bn + n2; bw = bw + wsrc * n2;
                            }
                            tn = an; tw = aw / avoid_zero(an);
                            num = bn; wgt = bw / avoid_zero(bn);
                        } else {
                            T bn(0.0), bw(0.0);
                            if (regtap) {
#pragma unroll
                                for (int d = 0; d < WTAP; d++)
                                    if (d <= dmax) {
                                        T n2 = s_num[cell - d] * gt[d];
                                        bn = bn + n2;
                                        bw = bw + (dwt[d] + s_wgt[cell - d]) * n2;
                                    }
                            } else {
                                for (int d = 0; d <= dmax; d++) {
                                    T n2 = s_num[cell - d] * tg[d * L + l];
                                    bn = bn + n2;
                                    bw = bw + (tdw[d * L + l] + s_wgt[cell - d]) * n2;
                                }
                            }
                            num = bn; wgt = bw / avoid_zero(bn);
                        }
                    }
                    __syncwarp();
                    if (valid) {
                        s_num[cell] = num; s_wgt[cell] = wgt;
                        if (trans_now) { s_tn[A.w_tr_off[s] + cell - c0] = tn; s_tw[A.w_tr_off[s] + cell - c0] = tw; }
                    }
                }
            }
            // ================= g3l_understocking (R/likelihood_understocking.R:36-46)
            if (have_us) {
                T u(0.0);
                if (pred_now) {
                    T o1 = warp_sum_all(us1), o2 = warp_sum_all(us2);
                    T tot = o1 - o2;
                    u = (us_power == 2.0) ? tot * tot : xpow(tot, T(us_power));
                }
                nll = nll + us_weight * u;
                if (lane == 0) s_ctot[A.NNLL - 2] = s_ctot[A.NNLL - 2] + u;
            }
            __syncwarp();

            // ================= pass B: maturation hand-over, renewal, collect vectors
            for (int s = 0; s < NS; s++) {
                const int32_t* St = I + I[G3H_STOCK_OFF] + s * G3S_LEN;
                const int L = St[G3S_L], nage = St[G3S_NAGE], narea = St[G3S_NAREA], c0 = St[G3S_CELL_OFF], ncell = L * nage * narea;
                const int G = A.w_G[s], CH = G * L;
                const int gcol = lane / L, l = lane - gcol * L;
                const bool trans_now = St[G3S_GROW_N] >= 0 && St[G3S_N_OUT] > 0 && ((St[G3S_TRANS_MASK] >> step) & 1);
                const bool inc_now = (A.w_inc_steps[s] >> step) & 1;
                const int n_renew = St[G3S_N_RENEW];
                bool ren_now = false;
                for (int k = 0; k < n_renew; k++) {
                    int rs = I[St[G3S_RENEW_OFF] + k * G3R_LEN + G3R_RUN_STEP];
                    ren_now = ren_now || rs == 0 || rs == step + 1;
                }
                if (!trans_now && !inc_now && !ren_now && !any_x) continue;
                for (int k = 0; k < A.w_ntile[s]; k++) {
                    const int cell = c0 + k * CH + lane;
                    if (!(lane < CH && cell < c0 + ncell)) continue;
                    const int col = k * G + gcol;
                    const int ridx = col / nage, aidx = col - ridx * nage;
                    T num = s_num[cell], wgt = s_wgt[cell];
                    // g3a_step_transition (R/action_mature.R:78-134)
                    if (inc_now) {
                        const int src = A.w_inc_tr[cell];
                        if (src >= 0 && ((A.inc_mask[cell] >> step) & 1)) {
                            T moved = s_tn[src] * slotv[A.inc_slot[cell]];
                            wgt = ratio_add_pop(wgt, num, s_tw[src], moved);
                            num = num + moved;
                        }
                    }
                    if (trans_now) {      // unclaimed remainder moves back (move_remainder = TRUE)
                        const T tn = s_tn[A.w_tr_off[s] + cell - c0];
                        T rem = tn;
                        for (int j = A.rem_off[cell]; A.rem_slots[j] >= 0; j++) rem = rem - tn * slotv[A.rem_slots[j]];
                        num = num + rem;
                    }
                    // g3a_renewal (R/action_renewal.R:166-188)
                    if (ren_now)
                        for (int r = 0; r < n_renew; r++) {
                            const int32_t* Rn = I + St[G3S_RENEW_OFF] + r * G3R_LEN;
                            if (aidx == Rn[G3R_AGE_IDX] && (Rn[G3R_RUN_STEP] == 0 || Rn[G3R_RUN_STEP] == step + 1)) {
                                int fs = I[Rn[G3R_FACTOR_OFF] + yidx * narea + ridx];
                                if (fs >= 0) {
                                    T rn = sm[A.o_rd[s] + r * L + l] * slotv[fs];
                                    wgt = ratio_add_pop(wgt, num, sm[A.w_o_rw[s] + r * L + l], rn);
                                    num = num + rn;
                                }
                            }
                        }
                    s_num[cell] = num; s_wgt[cell] = wgt;
                    // collect vectors (R/likelihood_distribution.R:275-287)
#pragma unroll
                    for (int x = 0; x < MAXX; x++)
                        if (xact[x]) {
                            T* xs = sm + A.o_xslot[x];
                            if (A.w_x_kind[x] == 0) { if (A.w_x_unit[x] == 0) xs[cell] = xs[cell] / avoid_zero(wgt); }
                            else xs[cell] = A.w_x_unit[x] == 0 ? num : num * wgt;
                        }
                }
            }
            __syncwarp();

            // ================= g3l_distribution (R/likelihood_distribution.R:275-394)
            if (any_x)
            for (int c = 0; c < A.NCOMP; c++) {
                const int32_t* C = I + I[G3H_COMP_OFF] + c * G3C_LEN;
                const int ti = I[C[G3C_TIDX_OFF] + t];
                if (ti < 0 || !(spar[C[G3C_WEIGHT_PAR]] > 0.0)) continue;
                const int func = C[G3C_FUNC], nd = C[G3C_N_DEST], nt = C[G3C_N_TIMES];
                const bool si = func == G3F_SI_LOG || func == G3F_SI_LINEAR;
                T* ms = sm + A.o_ms[c] + (si ? ti * nd : 0);
                const T* xs = sm + A.o_xslot[C[G3C_XBUF]];
                if (C[G3C_MODE] == 0) {
                    const int32_t* ptr = I + C[G3C_CSR_PTR_OFF];
                    const int32_t* idx = I + C[G3C_CSR_IDX_OFF];
                    const double* cf = R + C[G3C_CSR_COEF_ROFF];
                    for (int e = lane; e < nd; e += 32) {
                        T acc = ms[e];
                        for (int j = ptr[e]; j < ptr[e + 1]; j++) acc = acc + cf[j] * xs[idx[j]];
                        ms[e] = acc;
                    }
                } else {
                    const int32_t* cd = I + C[G3C_CELL_DEST_OFF];
                    const double* cf = R + C[G3C_CELL_COEF_ROFF];
                    for (int e = 0; e < nd; e++) {
                        T v(0.0);
                        for (int i = lane; i < NC; i += 32) if (cd[i] == e) v = v + cf[i] * xs[i];
                        v = warp_sum_all(v);
                        if (lane == 0) ms[e] = ms[e] + v;
                    }
                }
                __syncwarp();
                if (!I[C[G3C_CMP_OFF] + t]) continue;
                const T weight = mkpar<T>(spar, seed, C[G3C_WEIGHT_PAR]);
                const int nb = C[G3C_N_BINS], nsl = C[G3C_N_SLICES], nag = C[G3C_N_AREAG];
                T cnll(0.0);
                bool scored = false;
                if (func == G3F_SUMOFSQUARES) {     // R/likelihood_distribution.R:3-39
                    const int per = nsl * nb;
                    for (int ag = 0; ag < nag; ag++) {
                        T tot(0.0);
                        for (int e = lane; e < per; e += 32) tot = tot + ms[ag * per + e];
                        const T atot = avoid_zero(warp_sum_all(tot));
                        T d2(0.0);
                        for (int e = lane; e < per; e += 32) {
                            T dlt = ms[ag * per + e] / atot - A.obsprop[A.obs_off[c] + (size_t)ti * nd + ag * per + e];
                            d2 = d2 + dlt * dlt;
                        }
                        cnll = cnll + warp_sum_all(d2);
                    }
                    scored = true;
                } else if (func == G3F_MULTINOMIAL) {   // R/likelihood_distribution.R:41-60
                    T term(0.0);
                    const double eps = R[C[G3C_PARAM_ROFF]];
                    for (int e = lane; e < nd; e += 32) {
                        const T* sl = ms + (e / nb) * nb;
                        T tot(0.0);
                        for (int k = 0; k < nb; k++) tot = tot + sl[k];
                        double o = A.obsprop[A.obs_off[c] + (size_t)ti * nd + e];
                        term = term + o * xlog(dif_pmax_c(ms[e] / avoid_zero(tot), 1.0 / (nb * eps), 1e4));
                    }
                    T likely = warp_sum_all(term);
                    cnll = 2.0 * (-likely + A.obsconst[A.obsconst_off[c] + ti]);
                    scored = true;
                } else if (ti == nt - 1) {              // surveyindices: R/likelihood_distribution.R:82-125
                    const double fa = R[C[G3C_PARAM_ROFF]], fb = R[C[G3C_PARAM_ROFF] + 1];
                    const T* series = sm + A.o_ms[c];
                    for (int e = 0; e < nd; e++) {
                        T sN(0.0); double sI = 0.0;
                        for (int i = lane; i < nt; i += 32) {
                            T mv = series[i * nd + e];
                            sN = sN + (func == G3F_SI_LOG ? xlog(avoid_zero(mv)) : mv);
                            sI += A.obsprop[A.obs_off[c] + (size_t)i * nd + e];
                        }
                        const T mN = warp_sum_all(sN) / (double)nt;
                        const T mI = T(warp_sum_all(sI)) / (double)nt;
                        T beta(fb), alpha(fa);
                        if (isnan(fb)) {
                            T sxy(0.0), sxx(0.0);
                            for (int i = lane; i < nt; i += 32) {
                                T mv = series[i * nd + e];
                                T N = func == G3F_SI_LOG ? xlog(avoid_zero(mv)) : mv;
                                double Iv = A.obsprop[A.obs_off[c] + (size_t)i * nd + e];
                                sxy = sxy + (Iv - mI) * (N - mN); sxx = sxx + (N - mN) * (N - mN);
                            }
                            beta = warp_sum_all(sxy) / avoid_zero(warp_sum_all(sxx));
                        }
                        if (isnan(fa)) alpha = mI - beta * mN;
                        T r2(0.0);
                        for (int i = lane; i < nt; i += 32) {
                            T mv = series[i * nd + e];
                            T N = func == G3F_SI_LOG ? xlog(avoid_zero(mv)) : mv;
                            T r = alpha + beta * N - A.obsprop[A.obs_off[c] + (size_t)i * nd + e];
                            r2 = r2 + r * r;
                        }
                        cnll = cnll + warp_sum_all(r2);
                    }
                    scored = true;
                }
                if (scored) {
                    nll = nll + weight * cnll;
                    if (lane == 0) s_ctot[c] = s_ctot[c] + cnll;
                }
                if (!si) for (int e = lane; e < nd; e += 32) ms[e] = T(0.0);     // zero counters for the next period
                __syncwarp();
            }

            // ================= g3a_age (R/action_age.R:51-85) + movement hand-off
            if (final_step) {
                for (int s = 0; s < NS; s++) {
                    const int32_t* St = I + I[G3H_STOCK_OFF] + s * G3S_LEN;
                    if (!St[G3S_AGE_ENABLE]) continue;
                    const int L = St[G3S_L], nage = St[G3S_NAGE], narea = St[G3S_NAREA], c0 = St[G3S_CELL_OFF];
                    const int age_n_out = St[G3S_AGE_N_OUT];
                    // each lane walks one (area, length) chain of ages, oldest first
                    for (int i = lane; i < narea * L; i += 32) {
                        const int r = i / L, l = i - r * L;
                        const int base = c0 + r * nage * L + l;
                        const int top = base + (nage - 1) * L;
                        if (age_n_out > 0) { s_mv[A.w_mv_off[s] + i] = s_num[top]; s_mv[A.w_mv_off[s] + narea * L + i] = s_wgt[top]; }
                        if (nage == 1) {
                            if (age_n_out > 0) s_num[top] = T(0.0);
                            continue;
                        }
                        if (age_n_out > 0) { s_num[top] = s_num[top - L]; s_wgt[top] = s_wgt[top - L]; }
                        else {
                            s_wgt[top] = ratio_add_pop(s_wgt[top], s_num[top], s_wgt[top - L], s_num[top - L]);
                            s_num[top] = s_num[top] + s_num[top - L];
                        }
                        for (int a = nage - 2; a >= 1; a--) {
                            s_num[base + a * L] = s_num[base + (a - 1) * L];
                            s_wgt[base + a * L] = s_wgt[base + (a - 1) * L];
                        }
                        s_num[base] = T(0.0);
                    }
                }
                __syncwarp();
                for (int i = lane; i < A.w_n_agedst; i += 32) {
                    const int4 e = A.w_agedst[i];
                    const int cell = e.x;
                    T moved = s_mv[e.y] * slotv[e.w];
                    s_wgt[cell] = ratio_add_pop(s_wgt[cell], s_num[cell], s_mv[e.z], moved);
                    s_num[cell] = s_num[cell] + moved;
                }
                __syncwarp();
            }
        }   // time loop

        __syncwarp();
        if (lane == 0) {
            T out = bad_time ? T(nan("")) : nll;
            if (tile_t == 0) {
                A.nll[b] = val(out);
                if (A.comp) for (int c = 0; c < A.NNLL; c++) A.comp[b * A.NNLL + c] = val(s_ctot[c]);
            }
            if (NTAN > 0 && A.grad)
                for (int k = 0; k < NTAN; k++)
                    if (tile_t * NTAN + k < A.K) A.grad[b * A.K + tile_t * NTAN + k] = tan_of(out, k);
        }
    }
}

}  // namespace g3
