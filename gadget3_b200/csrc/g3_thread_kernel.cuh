// g3_thread_kernel.cuh -- one THREAD per parameter vector (round 1's production kernel). Today it serves the
// single-vector report run (g3cuda_report: per-step nll rows, state histories, model arrays), forward-mode
// gradient batches, and models the tile kernel (g3_tile.cuh) refuses.
//
// Mapping. A lane is an independent evaluation: the 32 lanes of a warp walk the SAME cell of the SAME time step
// of 32 different parameter vectors, so everything the reference's objective
// (baseline/multiple-substocks.cpp:692-2161) does for one vector is straight scalar code per thread: control flow
// is warp-uniform by construction, nothing is exchanged between lanes, sums are serial in the reference's order.
// State. The per-evaluation working set (a few thousand values) lives in a global workspace [entry][thread]
// (every warp access is one coalesced line). With tens of thousands of evaluations in flight that workspace is
// far larger than L2, so every step streams it from HBM (measured: 3.45 MB of DRAM traffic per C2 evaluation,
// the kernel is latency-bound on those loads, fp64 pipe 15 % busy -- profiles/r01_thread_kernel_v7_*): the
// reason the tile kernel, which keeps the state in shared memory, replaced it for value batches.
// Per step there is ONE fused pass over the cells (predation -> natural mortality -> banded growth with the
// maturing split -> maturation hand-over -> renewal -> likelihood collection), possible because length growth
// only looks down the length axis and maturation sources are processed before their destinations; plus a totals
// pre-pass on steps with landings and the ageing pass at the end of a year.
#pragma once
#include "g3_kargs.cuh"

namespace g3 {

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// CB: block-size cap the variant is compiled for (__launch_bounds__); NW: growth window = maxlengthgroupgrowth + 1 source cells kept in registers per column
// CONSTMAT: some stock matures with g3a_mature_constant (needs a third register window)
// REPORT: single-vector report run (g3cuda_report): per-step nll rows, state histories, model arrays
// STRIDE: thread stride of the workspace (>= launched threads); the report variant runs one warp
template <class T, int CB, int NW, bool CONSTMAT, bool REPORT = false, unsigned STRIDE = WS_STRIDE>
__global__ void __launch_bounds__(CB) eval_thread_kernel(const __grid_constant__ KArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int32_t* __restrict__ I = A.I;
    const double* __restrict__ R = A.R;
    const unsigned gtid = blockIdx.x * blockDim.x + threadIdx.x;
    // The thread stride of the workspace is a COMPILE-TIME constant (>= the launched threads), so that
    // neighbouring entries sit at immediate offsets from a running pointer: no address arithmetic
    // in the inner loops. (entries x WS_STRIDE < 2^32 is checked at launch.)
    constexpr unsigned NTH = STRIDE;
    T* __restrict__ ws = reinterpret_cast<T*>(A.t_ws) + gtid;
#define W(e) ws[(unsigned)(e) * NTH]
    T* s_eot = reinterpret_cast<T*>(smem_raw) + threadIdx.x;      // [MAXTS][blockDim]: landings / total suitable biomass
#define EOT(k) s_eot[(k) * blockDim.x]

    const int NS = A.NS, NSTEPS = A.NSTEPS, NT = A.NT;
    constexpr int NTAN = Tan<T>::n;
    const int ntiles = NTAN > 0 ? (A.K + NTAN - 1) / NTAN : 1;
    const double us_power = R[I[G3H_US_ROFF]], us_weight = R[I[G3H_US_ROFF] + 1];
    const bool have_us = I[G3H_US_N] > 0;

    // Every thread of a block runs the same trip counts (threads past the end of the batch redo
    // vector 0 without writing), so that the block can re-align its warps once per time step:
    // warps that execute the same code at the same time share instruction fetches.
    for (long long base0 = (long long)blockIdx.x * blockDim.x; base0 < A.B * ntiles; base0 += (long long)gridDim.x * blockDim.x) {
        const bool live = base0 + threadIdx.x < A.B * ntiles;
        const long long work = live ? base0 + threadIdx.x : 0;
        const long long b = work / ntiles;
        const int tile_t = (int)(work - b * ntiles);
        int seed[NTAN > 0 ? NTAN : 1];
#pragma unroll
        for (int k = 0; k < (NTAN > 0 ? NTAN : 1); k++) {
            int g = tile_t * NTAN + k;
            seed[k] = (NTAN > 0 && g < A.K) ? A.tangent_idx[g] : -1;
        }
        const double* __restrict__ prow = A.par + b * A.NPAR;

        // ---- g3a_time (R/action_time.R:75-91)
        const double retro = I[G3H_PAR_RETRO] >= 0 ? prow[I[G3H_PAR_RETRO]] : 0.0;
        const double proj = I[G3H_PAR_PROJECT] >= 0 ? prow[I[G3H_PAR_PROJECT]] : 0.0;
        const int total_steps = NSTEPS * (I[G3H_END_YEAR] - (int)floor(retro) - I[G3H_START_YEAR] + (int)floor(proj)) + NSTEPS - 1;
        const bool bad_time = !(retro >= 0.0) || !(proj >= 0.0) || total_steps + 1 > NT;

        for (int i = 0; i < A.NSLOT; i++) W(A.t_slot + i) = eval_slot_g<T>(I, R, prow, seed, i);
#define SLOT(i) W(A.t_slot + (i))

        // the set of suitability parameters in force at step t (time-varying selectivity), 0 where there is one
        auto suit_set = [&](const int32_t* Q, const int t) { return Q[G3PP_SUIT_TMAP] >= 0 ? I[Q[G3PP_SUIT_TMAP] + t] : 0; };
        // ================= per-evaluation tables
        for (int q = 0; q < A.NPP; q++) {           // suitability (R/suitability.R:5-30): [predator length][prey length]
            const int32_t* Q = I + I[G3H_PP_OFF] + q * G3PP_LEN;
            const int32_t* P = I + I[G3H_PRED_OFF] + Q[G3PP_PRED] * G3P_LEN;
            const int32_t* Ps = I + I[G3H_STOCK_OFF] + Q[G3PP_PREY] * G3S_LEN;
            const bool spred = P[G3P_KIND] == G3PRED_STOCK;
            const int32_t* Pp = I + I[G3H_STOCK_OFF] + (spred ? P[G3P_STOCK] : 0) * G3S_LEN;
            const int PL = spred ? Pp[G3S_L] : 1, L = Ps[G3S_L];
            // prey_preference p (R/action_predate.R:220-225): (S E B)^p = S^p E^p B^p; the table then holds S^p E^(p-1) -- x E x B^p
            // in the predator's total, x B^p in the consumption (one E cancels against cons' 1 / E)
            const bool pref = Q[G3PP_PREF] >= 0;
            const T pp_ = pref ? SLOT(Q[G3PP_PREF]) : T(1.0), ef = pref ? xpow(SLOT(Q[G3PP_ENERGY]), pp_ - 1.0) : T(1.0);
            // one table per set of suitability parameters: one set, unless the selectivity varies with time (by_year / by_step
            // tables, g3_timevariable); the report shows the set of the last step
            const int last_set = Q[G3PP_SUIT_TMAP] >= 0 ? I[Q[G3PP_SUIT_TMAP] + min(total_steps, NT - 1)] : 0;
            for (int sset = 0; sset < A.w_suit_nset[q]; sset++) {
                const int32_t* ss = I + Q[G3PP_SUIT_OFF] + 6 * sset;
                T sp[5];
#pragma unroll
                for (int k = 0; k < 5; k++) sp[k] = ss[k] >= 0 ? SLOT(ss[k]) : T(0.0);
                for (int pl = 0; pl < PL; pl++)
                    for (int l = 0; l < L; l++) {
                        const T sv = suitability_of<T>(Q[G3PP_SUIT_KIND], sp, R[Ps[G3S_MIDLEN_ROFF] + l],
                            spred && Q[G3PP_SUIT_KIND] != G3SUIT_ANDERSENFLEET ? R[Pp[G3S_MIDLEN_ROFF] + pl] : R[Ps[G3S_MIDLEN_ROFF] + L - 1]);
                        if (REPORT && live && pref && sset == last_set) A.report[A.rep_pp[q] + pl * L + l] = val(sv);       // suit_<prey>_<pred>__report is S itself
                        W(A.t_suitq[q] + (sset * PL + pl) * L + l) = pref ? xpow(sv, pp_) * ef : sv;
                    }
            }
        }
        // stock predators: maximum consumption per predator length without the step length,
        // m0 exp(m1 T - m2 T^3) l^m3 (R/action_predate.R:190-205)
        for (int pi = 0; pi < I[G3H_NPRED]; pi++) {
            const int32_t* P = I + I[G3H_PRED_OFF] + pi * G3P_LEN;
            if (P[G3P_KIND] != G3PRED_STOCK) continue;
            const int32_t* Pp = I + I[G3H_STOCK_OFF] + P[G3P_STOCK] * G3S_LEN;
            const int32_t* ps = I + P[G3P_SLOT_OFF];
            const T m0 = SLOT(ps[0]), m1 = SLOT(ps[1]), m2 = SLOT(ps[2]), m3 = SLOT(ps[3]), temp = SLOT(ps[5]);
            const T tf = m0 * xexp(m1 * temp - m2 * temp * temp * temp);
            for (int pl = 0; pl < Pp[G3S_L]; pl++) W(A.t_pmc[pi] + pl) = tf * xpow(T(R[Pp[G3S_MIDLEN_ROFF] + pl]), m3);
        }
        // migration matrices per step of the year: proportion moving src -> dest (R/action_migrate.R:2-15,49-56)
        for (int s = 0; s < NS; s++) {
            const int32_t* St = I + I[G3H_STOCK_OFF] + s * G3S_LEN;
            if (St[G3S_MIGRATE_OFF] < 0) continue;
            const int NR = St[G3S_NAREA];
            for (int st = 0; st < NSTEPS; st++) {
                const int32_t* mig = I + St[G3S_MIGRATE_OFF] + st * NR * NR;
                for (int src = 0; src < NR; src++) {
                    T tot(0.0);
                    for (int d = 0; d < NR; d++) {
                        T v = (d != src && mig[d * NR + src] >= 0) ? SLOT(mig[d * NR + src]) : T(0.0);
                        v = v * v;
                        tot = tot + v;
                        W(A.t_mig[s] + (st * NR + d) * NR + src) = v / 1.0;
                    }
                    W(A.t_mig[s] + (st * NR + src) * NR + src) = (1.0 - tot) / 1.0;
                }
            }
        }
        for (int s = 0; s < NS; s++) {
            const int32_t* St = I + I[G3H_STOCK_OFF] + s * G3S_LEN;
            const int L = St[G3S_L], ncol = St[G3S_NAGE] * St[G3S_NAREA];
            const double* midlen = R + St[G3S_MIDLEN_ROFF];
            const double pdl = R[St[G3S_PLUSDL_ROFF]];
            // natural mortality factor per column and distinct step length (R/action_naturalmortality.R:26)
            if (St[G3S_MORT_OFF] >= 0)
                for (int ms = 0; ms < A.w_mort_nset[s]; ms++)       // (one set unless M varies with time)
                    for (int idt = 0; idt < A.w_ndt; idt++)
                        for (int col = 0; col < ncol; col++)
                            W(A.t_mort[s] + (ms * A.w_ndt + idt) * ncol + col) = xexp(-SLOT(I[St[G3S_MORT_OFF] + ms * ncol + col]) * (A.w_dt_len[idt] / 12.0));
            // renewal length distribution, normalised, and weight (R/action_renewal.R:56-80)
            for (int k = 0; k < St[G3S_N_RENEW]; k++) {
                const int32_t* Rn = I + St[G3S_RENEW_OFF] + k * G3R_LEN;
                const T mean = SLOT(Rn[G3R_MEAN]), sd = avoid_zero(SLOT(Rn[G3R_SD]));
                const T wa = SLOT(Rn[G3R_WALPHA]), wb = SLOT(Rn[G3R_WBETA]);
                T tot(0.0);
                for (int l = 0; l < L; l++) { T d = dnorm(midlen[l], mean, sd); W(A.t_rd[s] + k * L + l) = d; tot = tot + d; }
                const T den = avoid_zero(tot);
                for (int l = 0; l < L; l++) {
                    W(A.t_rd[s] + k * L + l) = W(A.t_rd[s] + k * L + l) / den * 10000.0;
                    W(A.t_rw[s] + k * L + l) = wa * xpow(midlen[l], wb);
                }
            }
            const int n2 = St[G3S_GROW_N];
            if (n2 < 0) continue;
            const bool cont = St[G3S_MAT_KIND] == G3MAT_CONTINUOUS && St[G3S_N_OUT] > 0;
            const int gsz = (n2 + 1) * L;
            for (int gset = 0; gset < A.w_grow_nset[s]; gset++) {       // (one set unless the growth parameters vary with time)
            const int32_t* gs = I + St[G3S_GROW_OFF] + 5 * gset;
            {
                const T wb = SLOT(gs[4]);
                for (int l = 0; l < L; l++) W(A.t_pw[s] + gset * L + l) = xpow(midlen[l], wb);
            }
            for (int idt = 0; idt < A.w_ndt; idt++) {
                const double dt = A.w_dt_len[idt] / 12.0;
                const T linf = SLOT(gs[0]), kk = 1.0 - xexp(-(SLOT(gs[1])) * dt), beta = avoid_zero(SLOT(gs[2]));
                T malpha(0.0), ml50(0.0), mbeta(0.0), ma50(0.0);
                const int ncset = A.w_ncset[s];     // > 1: the maturing share has an age term, one pair of bands per column
                if (cont) {
                    const int32_t* ms = I + St[G3S_MAT_OFF] + 4 * gset; malpha = SLOT(ms[0]); ml50 = SLOT(ms[1]);
                    if (ms[2] >= 0) { mbeta = SLOT(ms[2]); ma50 = SLOT(ms[3]); }
                }
                for (int cset = 0; cset < ncset; cset++) {
                const double c_age = (double)(St[G3S_MINAGE] + cset % St[G3S_NAGE]);
                // source-oriented band in scratch: [v][x * L + l], v = 0 all, 1 maturing, 2 staying
                for (int l = 0; l < L; l++) {
                    // g3a_grow_lengthvbsimple (R/action_grow.R:54) inside the bbinom wrapper (R/action_grow.R:220)
                    T dl = avoid_zero((linf - midlen[l]) * kk);
                    T delta = avoid_zero(dl / pdl);
                    T alpha = (beta * delta) / ((double)n2 - delta);
                    // growth_bbinom in product form, see g3_kernel.cuh
                    T g = T(1.0);
                    for (int j = 0; j < n2; j++) g = g * ((beta + (double)j) / xabs(alpha + beta + (double)j));
                    T m0(0.0);
                    if (cont) {         // g3a_mature_constant as M0 (R/action_mature.R:46-72)
                        T inner = 0.0 - malpha * (midlen[l] - ml50);
                        if (ncset > 1) inner = inner - mbeta * (c_age - ma50);
                        m0 = 1.0 / (1.0 + xexp(inner));
                    }
                    for (int x = 0; x <= n2; x++) {
                        W(A.t_scr + x * L + l) = g;
                        if (cont) {     // g3a_mature_continuous (R/action_mature.R:1-42): alpha ldelta + beta cur_step_size
                            T ratio = ncset > 1 ? dif_pminmax_c(m0 * (malpha * (pdl * x) + mbeta * dt), 0.0, 1.0, 1e5)
                                                : dif_pminmax_c(m0 * (malpha * (pdl * x)), 0.0, 1.0, 1e5);
                            W(A.t_scr + gsz + x * L + l) = g * ratio;
                            W(A.t_scr + 2 * gsz + x * L + l) = g * (1.0 - ratio);
                        }
                        if (x < n2)
                            g = g * ((double)(n2 - x) / (double)(x + 1)) * (xabs(alpha + (double)x) / (beta + (double)(n2 - x - 1)));
                    }
                }
                // destination-oriented band [d][l] = P(l - d -> l), plus-group tail sums folded in (R/action_grow.R:233-271)
                for (int v = (cset == 0 ? 0 : 1); v < (cont ? 3 : 1); v++) {
                    const int o = v == 0 ? A.t_g[s] + (gset * A.w_ndt + idt) * gsz
                                         : (v == 1 ? A.t_gr[s] : A.t_gn[s]) + ((gset * ncset + cset) * A.w_ndt + idt) * gsz;
                    for (int d = 0; d <= n2; d++)
                        for (int j = 0; j < L; j++) {
                            const int src = j - d;
                            T g(0.0);
                            if (src >= 0) {
                                if (j < L - 1) g = W(A.t_scr + v * gsz + d * L + src);
                                else for (int x = d; x <= n2; x++) g = g + W(A.t_scr + v * gsz + x * L + src);
                            }
                            W(o + d * L + j) = g;
                        }
                }
                }   // cset
            }
            }   // gset
        }

        // ================= g3a_initialconditions (R/action_renewal.R:83-163)
        for (int s = 0; s < NS; s++) {
            const int32_t* St = I + I[G3H_STOCK_OFF] + s * G3S_LEN;
            const int L = St[G3S_L], c0 = St[G3S_CELL_OFF], ncol = St[G3S_NAGE] * St[G3S_NAREA];
            const double* midlen = R + St[G3S_MIDLEN_ROFF];
            for (int col = 0; col < ncol; col++) {
                if (St[G3S_INIT_OFF] < 0) {
                    for (int l = 0; l < L; l++) { W(A.t_num + c0 + col * L + l) = T(0.0); W(A.t_wgt + c0 + col * L + l) = T(1.0); }
                    continue;
                }
                const int32_t* sl = I + St[G3S_INIT_OFF] + col * 5;
                const T mean = SLOT(sl[0]), sd = avoid_zero(SLOT(sl[1])), factor = SLOT(sl[2]), wa = SLOT(sl[3]), wb = SLOT(sl[4]);
                T tot(0.0);
                for (int l = 0; l < L; l++) { T d = dnorm(midlen[l], mean, sd); W(A.t_num + c0 + col * L + l) = d; tot = tot + d; }
                const T den = avoid_zero(tot);                           // normalize_vec, R/aab_env.R:37-41
                for (int l = 0; l < L; l++) {
                    const int cell = c0 + col * L + l;
                    W(A.t_num + cell) = W(A.t_num + cell) / den * 10000.0 * factor;
                    W(A.t_wgt + cell) = wa * xpow(midlen[l], wb);
                }
            }
        }

        // maturation: share of the maturing fish of a column that no output stock claims (R/action_mature.R:126-131),
        // and the g3a_mature_constant ratio per (column, length) (R/action_mature.R:44-74)
        for (int s = 0; s < NS; s++) {
            const int32_t* St = I + I[G3H_STOCK_OFF] + s * G3S_LEN;
            if (St[G3S_GROW_N] < 0 || St[G3S_N_OUT] == 0) continue;
            const int L = St[G3S_L], nage = St[G3S_NAGE], ncol = nage * St[G3S_NAREA], c0 = St[G3S_CELL_OFF];
            for (int col = 0; col < ncol; col++) {
                T remf(1.0);
                for (int k = A.rem_off[c0 + col * L]; A.rem_slots[k] >= 0; k++) remf = remf - SLOT(A.rem_slots[k]);
                W(A.t_remf + A.t_colbase[s] + col) = remf;
                if (CONSTMAT && St[G3S_MAT_KIND] == G3MAT_CONSTANT) {
                    const int32_t* msl = I + St[G3S_MAT_OFF];
                    const int aidx = col % nage;
                    for (int l = 0; l < L; l++) {
                        T inner(0.0);
                        if (msl[0] >= 0) inner = inner - SLOT(msl[0]) * (R[St[G3S_MIDLEN_ROFF] + l] - SLOT(msl[1]));
                        if (msl[2] >= 0) inner = inner - SLOT(msl[2]) * ((double)(St[G3S_MINAGE] + aidx) - SLOT(msl[3]));
                        W(A.t_cmat[s] + col * L + l) = 1.0 / (1.0 + xexp(inner));
                    }
                }
            }
        }

        T nll(0.0);
        for (int i = 0; i < A.NNLL; i++) W(A.t_ctot + i) = T(0.0);
        // ---- g3l_bounds_penalty (R/likelihood_bounds.R:13-16), cur_time == 0
        if (A.NBOUND > 0) {
            T pen(0.0);
            const double bw = R[I[G3H_BOUND_WS_ROFF]], bs = R[I[G3H_BOUND_WS_ROFF] + 1];
            for (int i = 0; i < A.NBOUND; i++) {
                T p = mkpar_g<T>(prow, seed, I[I[G3H_BOUND_OFF] + i]);
                double lo = R[I[G3H_BOUND_ROFF] + 2 * i], up = R[I[G3H_BOUND_ROFF] + 2 * i + 1];
                T a = lsa_c(bs * (p - up) / (up - lo), 0.0) + lsa_c(bs * (lo - p) / (up - lo), 0.0);
                pen = pen + a * a;
            }
            nll = nll + bw * pen;
            W(A.t_ctot + A.NNLL - 1) = pen;
            if (REPORT && live) A.report[1 + (size_t)(A.NNLL - 1) * NT] = val(pen);
        }
        for (int c = 0; c < A.NCOMP; c++) {         // zero the likelihood slabs
            const int32_t* C = I + I[G3H_COMP_OFF] + c * G3C_LEN;
            bool si = C[G3C_FUNC] == G3F_SI_LOG || C[G3C_FUNC] == G3F_SI_LINEAR;
            int n = C[G3C_N_DEST] * (si ? C[G3C_N_TIMES] : 1);
            for (int i = 0; i < n; i++) W(A.t_ms[c] + i) = T(0.0);
        }

        for (int t = 0; t < NT; t++) {
            __syncthreads();
            if (t > total_steps || bad_time) continue;
            const int step = t % NSTEPS, yidx = t / NSTEPS;
            const int idt = A.w_step_idt[step];
            const bool final_step = step == NSTEPS - 1;
            const bool pred_now = A.w_pred_active[t] != 0;
            // g3a_otherfood (R/action_renewal.R:276-308): number and mean weight assigned from formulas, every step
            for (int s = 0; s < NS; s++) {
                const int32_t* St = I + I[G3H_STOCK_OFF] + s * G3S_LEN;
                if (St[G3S_OTHERFOOD_OFF] < 0) continue;
                const int per = St[G3S_NAGE] * St[G3S_L];
                for (int i = 0; i < St[G3S_NAREA] * per; i++) {
                    const int32_t* sl = I + St[G3S_OTHERFOOD_OFF] + ((size_t)t * St[G3S_NAREA] + i / per) * 2;
                    W(A.t_num + St[G3S_CELL_OFF] + i) = SLOT(sl[0]); W(A.t_wgt + St[G3S_CELL_OFF] + i) = SLOT(sl[1]);
                }
            }
            if (REPORT && live)         // dstart_<stock>__num/wgt (R/action_report.R:177-213)
                for (int s = 0; s < NS; s++) {
                    const int32_t* St = I + I[G3H_STOCK_OFF] + s * G3S_LEN;
                    const size_t ncs = (size_t)St[G3S_L] * St[G3S_NAGE] * St[G3S_NAREA];
                    for (size_t i = 0; i < ncs; i++) {
                        A.report[A.rep_dstart[s] + (size_t)t * ncs + i] = val(W(A.t_num + St[G3S_CELL_OFF] + i));
                        A.report[A.rep_dstart[s] + (size_t)NT * ncs + (size_t)t * ncs + i] = val(W(A.t_wgt + St[G3S_CELL_OFF] + i));
                        // <stock>__renewalnum keeps its last renewal until the next one overwrites it
                        if (t > 0) A.report[A.rep_renew[s] + (size_t)t * ncs + i] = A.report[A.rep_renew[s] + (size_t)(t - 1) * ncs + i];
                    }
                }
            // components collecting at this step (bit c), and whether any of them reads a catch
            unsigned cact = 0, xact = 0;
            bool any_catch = false;
            for (int c = 0; c < A.NCOMP; c++) {
                const int32_t* C = I + I[G3H_COMP_OFF] + c * G3C_LEN;
                if (I[C[G3C_TIDX_OFF] + t] >= 0 && prow[C[G3C_WEIGHT_PAR]] > 0.0) {
                    cact |= 1u << c;
                    xact |= 1u << C[G3C_XBUF];
                    if (A.w_x_kind[C[G3C_XBUF]] == 0) any_catch = true;
                }
            }

            // ================= g3a_migrate (R/action_migrate.R:17-82): every (age, length) mixes its areas in place
            for (int s = 0; s < NS; s++) {
                if (!((A.t_mig_steps[s] >> step) & 1)) continue;
                const int32_t* St = I + I[G3H_STOCK_OFF] + s * G3S_LEN;
                const int L = St[G3S_L], nage = St[G3S_NAGE], NR = St[G3S_NAREA], c0 = St[G3S_CELL_OFF];
                const int mo = A.t_mig[s] + step * NR * NR;
                for (int i = 0; i < nage * L; i++) {
                    T sn[MAXMIG], sw[MAXMIG];
#pragma unroll
                    for (int r = 0; r < MAXMIG; r++)
                        if (r < NR) { sn[r] = W(A.t_num + c0 + r * nage * L + i); sw[r] = W(A.t_wgt + c0 + r * nage * L + i); }
#pragma unroll
                    for (int d = 0; d < MAXMIG; d++)
                        if (d < NR) {
                            T pn(0.0), pw(0.0);
#pragma unroll
                            for (int r = 0; r < MAXMIG; r++)
                                if (r < NR) {
                                    const T moved = sn[r] * W(mo + d * NR + r);
                                    pw = ratio_add_pop(pw, pn, sw[r], moved);       // stock_combine_subpop
                                    pn = pn + moved;
                                }
                            W(A.t_num + c0 + d * nage * L + i) = pn; W(A.t_wgt + c0 + d * nage * L + i) = pw;
                        }
                }
            }

            // ================= g3a_predate step 1: suitable biomass per (predator, area[, predator length]) (R/action_predate.R:314-333)
            if (pred_now) {
#pragma unroll
                for (int k = 0; k < MAXTS; k++) if (k < A.NTS) EOT(k) = T(0.0);
                for (int pi = 0; pi < I[G3H_NPRED]; pi++) {
                    const int32_t* P = I + I[G3H_PRED_OFF] + pi * G3P_LEN;
                    if (P[G3P_KIND] != G3PRED_STOCK) continue;
                    const int32_t* Pp = I + I[G3H_STOCK_OFF] + P[G3P_STOCK] * G3S_LEN;
                    for (int i = 0; i < Pp[G3S_NAREA] * Pp[G3S_L]; i++) W(A.t_pts[pi] + i) = T(0.0);
                }
                for (int s = 0; s < NS; s++) {
                    const int nq = A.stock_pp_off[s + 1] - A.stock_pp_off[s];
                    if (nq == 0) continue;
                    const int32_t* St = I + I[G3H_STOCK_OFF] + s * G3S_LEN;
                    const int L = St[G3S_L], nage = St[G3S_NAGE], c0 = St[G3S_CELL_OFF], ncol = nage * St[G3S_NAREA];
                    for (int j = 0; j < nq; j++) {
                        const int q = A.stock_pp[A.stock_pp_off[s] + j];
                        const int32_t* Q = I + I[G3H_PP_OFF] + q * G3PP_LEN;
                        const int32_t* P = I + I[G3H_PRED_OFF] + Q[G3PP_PRED] * G3P_LEN;
                        const bool spred = P[G3P_KIND] == G3PRED_STOCK;
                        const bool by_number = P[G3P_KIND] == G3PRED_NUMBERFLEET;       // suit = S * num (R/action_predate.R:7-11)
                        if (P[G3P_KIND] == G3PRED_LINEARFLEET || P[G3P_KIND] == G3PRED_EFFORTFLEET) continue;     // no total needed
                        const int PL = spred ? I[I[G3H_STOCK_OFF] + P[G3P_STOCK] * G3S_LEN + G3S_L] : 1;
                        const T energy = spred ? SLOT(Q[G3PP_ENERGY]) : T(1.0);
                        for (int col = 0; col < ncol; col++) {
                            const int ts = A.pp_tsid[q * A.maxnarea + col / nage];     // accumulator (fleet) / predator area idx (stock predator)
                            if (ts < 0) continue;
                            for (int pl = 0; pl < PL; pl++) {
                                T cs(0.0);
                                for (int l = 0; l < L; l++) {
                                    const int cell = c0 + col * L + l;
                                    T bm = by_number ? W(A.t_num + cell) : W(A.t_num + cell) * W(A.t_wgt + cell);
                                    if (Q[G3PP_PREF] >= 0) bm = xpow(bm, SLOT(Q[G3PP_PREF]));
                                    cs = cs + W(A.t_suitq[q] + suit_set(Q, t) * PL * L + pl * L + l) * bm;
                                }
                                if (spred) W(A.t_pts[Q[G3PP_PRED]] + ts * PL + pl) = W(A.t_pts[Q[G3PP_PRED]] + ts * PL + pl) + cs * energy;
                                else EOT(ts) = EOT(ts) + cs;
                            }
                        }
                    }
                }
                // landings over total suitable biomass: cons = E * (suit / totalsuit)
#pragma unroll
                for (int k = 0; k < MAXTS; k++)
                    if (k < A.NTS) {
                        const double E = R[A.w_ts_eoff[k] + (size_t)t * A.w_ts_estr[k]];
                        if (A.w_ts_kind[k] == G3PRED_LINEARFLEET || A.w_ts_kind[k] == G3PRED_EFFORTFLEET)
                            EOT(k) = T(E * (A.w_dt_len[idt] / 12.0));       // harvest rate x step length (R/action_predate.R:13-18,117-126)
                        else EOT(k) = E / EOT(k);
                    }
                // stock predators: predN * max_consumption * feeding_level / total_predsuit per (area, predator length)
                // (R/action_predate.R:207-238; energycontent cancels between suit and cons)
                for (int pi = 0; pi < I[G3H_NPRED]; pi++) {
                    const int32_t* P = I + I[G3H_PRED_OFF] + pi * G3P_LEN;
                    if (P[G3P_KIND] != G3PRED_STOCK) continue;
                    const int32_t* Pp = I + I[G3H_STOCK_OFF] + P[G3P_STOCK] * G3S_LEN;
                    const int PL = Pp[G3S_L], PA = Pp[G3S_NAGE];
                    const double dt = A.w_dt_len[idt] / 12.0;
                    const T hf = SLOT(I[P[G3P_SLOT_OFF] + 4]);
                    for (int pr = 0; pr < Pp[G3S_NAREA]; pr++)
                        for (int pl = 0; pl < PL; pl++) {
                            const T ts = W(A.t_pts[pi] + pr * PL + pl);
                            T predn(0.0);
                            for (int pa = 0; pa < PA; pa++) predn = predn + W(A.t_num + Pp[G3S_CELL_OFF] + (pr * PA + pa) * PL + pl);
                            const T feeding = ts / (hf * dt + ts);
                            W(A.t_pts[pi] + pr * PL + pl) = predn * (W(A.t_pmc[pi] + pl) * dt) * feeding / ts;
                        }
                }
            }

            // ================= stock by stock
            T us_tot(0.0);
            for (int s = 0; s < NS; s++) {
                const int32_t* St = I + I[G3H_STOCK_OFF] + s * G3S_LEN;
                const int L = St[G3S_L], nage = St[G3S_NAGE], narea = St[G3S_NAREA], c0 = St[G3S_CELL_OFF], ncol = nage * narea;
                const int nq = A.stock_pp_off[s + 1] - A.stock_pp_off[s];
                const bool prey_now = pred_now && nq > 0;
                const bool is_us = A.us_flag[s] != 0;
                const int grow_n = St[G3S_GROW_N], mat_kind = St[G3S_MAT_KIND];
                const bool trans_now = grow_n >= 0 && St[G3S_N_OUT] > 0 && ((St[G3S_TRANS_MASK] >> step) & 1);
                const bool cont_now = trans_now && mat_kind == G3MAT_CONTINUOUS;
                const bool const_now = CONSTMAT && trans_now && mat_kind == G3MAT_CONSTANT;
                const bool has_mort = St[G3S_MORT_OFF] >= 0;
                const bool inc_now = (A.w_inc_steps[s] >> step) & 1;
                const int n_renew = St[G3S_N_RENEW];
                int ren_mask = 0;       // renewal records that run at this step
                for (int k = 0; k < n_renew; k++) {
                    int rs = I[St[G3S_RENEW_OFF] + k * G3R_LEN + G3R_RUN_STEP];
                    if (rs == 0 || rs == step + 1) ren_mask |= 1 << k;
                }
                const bool catch_now = any_catch && nq > 0;
                const bool spawns_now = St[G3S_SPAWN_OFF] >= 0 && (I[St[G3S_SPAWN_OFF] + G3SP_RUN_STEP] == 0 || I[St[G3S_SPAWN_OFF] + G3SP_RUN_STEP] == step + 1);
                if (!prey_now && !has_mort && grow_n < 0 && !inc_now && !ren_mask && xact == 0 && !spawns_now) continue;
                const unsigned LWS = (unsigned)L * NTH;

                // ---------- g3a_predate (R/action_predate.R:335-406): cell order, as the reference sums
                if (prey_now) {
                    unsigned kx[MAXQ], ksuit[MAXQ];     // per predator: collect vectors it feeds, suitability table
                    int kpl[MAXQ], kpts[MAXQ];          // stock predators: predator length groups, table of their consumption factors
                    int kq[MAXQ];
#pragma unroll
                    for (int k = 0; k < MAXQ; k++) {
                        kx[k] = 0; ksuit[k] = 0; kpl[k] = 0; kpts[k] = 0; kq[k] = 0;
                        if (k < nq) {
                            const int q = A.stock_pp[A.stock_pp_off[s] + k];
                            kq[k] = q;
                            const int32_t* Q = I + I[G3H_PP_OFF] + q * G3PP_LEN;
                            const int32_t* P = I + I[G3H_PRED_OFF] + Q[G3PP_PRED] * G3P_LEN;
                            ksuit[k] = (unsigned)(A.t_suitq[q] + suit_set(Q, t) * (P[G3P_KIND] == G3PRED_STOCK ? I[I[G3H_STOCK_OFF] + P[G3P_STOCK] * G3S_LEN + G3S_L] : 1) * L) * NTH;
                            if (P[G3P_KIND] == G3PRED_STOCK) { kpl[k] = I[I[G3H_STOCK_OFF] + P[G3P_STOCK] * G3S_LEN + G3S_L]; kpts[k] = A.t_pts[Q[G3PP_PRED]]; }
                            for (int x = 0; x < A.NX; x++) if ((A.w_x_ppmask[x] >> q) & 1) kx[k] |= 1u << x;
                        }
                    }
                    T o1(0.0), o2(0.0);
                    T* pn = ws + (unsigned)(A.t_num + c0) * NTH;
                    const T* pw_ = ws + (unsigned)(A.t_wgt + c0) * NTH;
                    for (int col = 0; col < ncol; col++) {
                        const int ridx = col / nage;
                        T eotk[MAXQ];       // landings / total suitable biomass of each fleet in this column's area
                        int tsk[MAXQ];
#pragma unroll
                        for (int k = 0; k < MAXQ; k++) {
                            eotk[k] = T(0.0); tsk[k] = -1;
                            if (k < nq) {
                                tsk[k] = A.pp_tsid[A.stock_pp[A.stock_pp_off[s] + k] * A.maxnarea + ridx];
                                if (tsk[k] >= 0 && kpl[k] == 0) {
                                    eotk[k] = EOT(tsk[k]);
                                    if (A.w_ts_kind[tsk[k]] == G3PRED_EFFORTFLEET) eotk[k] = eotk[k] * SLOT(I[I[G3H_PP_OFF] + kq[k] * G3PP_LEN + G3PP_ENERGY]);
                                }
                            }
                        }
                        unsigned lN = 0;
                        for (int l = 0; l < L; l++, pn += NTH, pw_ += NTH, lN += NTH) {
                            const T num = *pn;
                            const T B0 = num * *pw_;
                            T tp(0.0), fx[MAXX];
#pragma unroll
                            for (int x = 0; x < MAXX; x++) fx[x] = T(0.0);
#pragma unroll
                            for (int k = 0; k < MAXQ; k++)
                                if (k < nq && tsk[k] >= 0) {
                                    T cf;
                                    if (kpl[k] == 0) cf = ws[ksuit[k] + lN] * eotk[k];
                                    else {      // sum over the predator's length groups of suitability x consumption factor
                                        cf = T(0.0);
                                        for (int pl = 0; pl < kpl[k]; pl++)
                                            cf = cf + ws[ksuit[k] + (unsigned)(pl * L) * NTH + lN] * W(kpts[k] + tsk[k] * kpl[k] + pl);
                                    }
                                    const int pfs = kpl[k] ? I[I[G3H_PP_OFF] + kq[k] * G3PP_LEN + G3PP_PREF] : -1;
                                    tp = tp + cf * (pfs >= 0 ? xpow(B0, SLOT(pfs)) : B0);        // biomass^prey_preference
                                    if (catch_now) {
#pragma unroll
                                        for (int x = 0; x < MAXX; x++) if ((kx[k] >> x) & 1) fx[x] = fx[x] + cf;
                                    }
                                }
                            T cr = xdiv_x(tp, avoid_zero_x(B0));       // (exact-division forms: see g3_math.cuh)
                            cr = dif_pmax_x(cr, 0.95, -1e3);
                            const T tpn = B0 * cr;
                            *pn = num * (1.0 - cr);
                            if (is_us) { o1 = o1 + tp; o2 = o2 + tpn; }
                            if (catch_now) {
                                const T cc = xdiv(T(1.0), avoid_zero(tp)) * tpn * B0;      // consconv * B0
#pragma unroll
                                for (int x = 0; x < MAXX; x++)
                                    if (((xact >> x) & 1) && A.w_x_kind[x] == 0) W(A.t_x[x] + c0 + col * L + l) = fx[x] * cc;
                            }
                            if (REPORT && live) {       // detail_<prey>_<pred>__suit / __cons (R/action_report.R:177-213)
                                const T cc = 1.0 / avoid_zero(tp) * tpn * B0;
                                const size_t ncs = (size_t)ncol * L, ci = (size_t)t * ncs + col * L + l;
                                for (int k = 0; k < nq; k++)
                                    if (tsk[k] >= 0) {
                                        const size_t base = A.rep_pp[kq[k]] + (size_t)(kpl[k] ? kpl[k] : 1) * L;
                                        T cf;
                                        if (kpl[k] == 0) {
                                            cf = ws[ksuit[k] + lN] * eotk[k];
                                            A.report[base + ci] = val(A.w_ts_kind[tsk[k]] == G3PRED_NUMBERFLEET ? ws[ksuit[k] + lN] * num : ws[ksuit[k] + lN] * num * *pw_);
                                        } else {
                                            cf = T(0.0);
                                            for (int pl = 0; pl < kpl[k]; pl++)
                                                cf = cf + ws[ksuit[k] + (unsigned)(pl * L) * NTH + lN] * W(kpts[k] + tsk[k] * kpl[k] + pl);
                                            const int pfs = I[I[G3H_PP_OFF] + kq[k] * G3PP_LEN + G3PP_PREF];
                                            if (pfs >= 0) {
                                                A.report[base + (size_t)NT * ncs + ci] = val(cf * xpow(B0, SLOT(pfs)) * (1.0 / avoid_zero(tp) * tpn));
                                                continue;
                                            }
                                        }
                                        A.report[base + (size_t)NT * ncs + ci] = val(cf * cc);
                                    }
                            }
                        }
                    }
                    if (is_us) us_tot = us_tot + (o1 - o2);
                } else if (catch_now) {
                    for (int x = 0; x < A.NX; x++)
                        if (((xact >> x) & 1) && A.w_x_kind[x] == 0)
                            for (int i = 0; i < ncol * L; i++) W(A.t_x[x] + c0 + i) = T(0.0);       // no landings: cons = 0
                }
                if (REPORT && live && !prey_now)        // steps without landings: the fleets' suit arrays are still filled
                    for (int k = 0; k < nq; k++) {
                        const int q = A.stock_pp[A.stock_pp_off[s] + k];
                        const int32_t* Q = I + I[G3H_PP_OFF] + q * G3PP_LEN;
                        const int pk = I[I[G3H_PRED_OFF] + Q[G3PP_PRED] * G3P_LEN + G3P_KIND];
                        if (pk == G3PRED_STOCK) continue;
                        for (int col = 0; col < ncol; col++) {
                            if (A.pp_tsid[q * A.maxnarea + col / nage] < 0) continue;
                            for (int l = 0; l < L; l++) {
                                const int cell = c0 + col * L + l;
                                const T sn = W(A.t_suitq[q] + suit_set(Q, t) * L + l) * W(A.t_num + cell);
                                A.report[A.rep_pp[q] + L + ((size_t)t * ncol + col) * L + l] = val(pk == G3PRED_NUMBERFLEET ? sn : sn * W(A.t_wgt + cell));
                            }
                        }
                    }
                // g3a_spawn, order 6 (R/action_spawn.R:121-168): does this stock spawn at this step? (This kernel serves the reports:
                // proportion, mortality and fecundity weights are worked out cell by cell instead of from tables.)
                bool spawn_here = false;
                T sp_prop[5], sp_mort[5], sp_p[5], sp_acc(0.0);
                int sp_kind = 0, sp_pk = 0, sp_mk = -1;
                if (St[G3S_SPAWN_OFF] >= 0) {
                    const int32_t* Sp = I + St[G3S_SPAWN_OFF];
                    spawn_here = Sp[G3SP_RUN_STEP] == 0 || Sp[G3SP_RUN_STEP] == step + 1;
                    sp_kind = Sp[G3SP_RECR_KIND]; sp_pk = Sp[G3SP_PROP_KIND]; sp_mk = Sp[G3SP_MORT_KIND];
#pragma unroll
                    for (int k = 0; k < 5; k++) {
                        sp_prop[k] = Sp[G3SP_PROP_SLOTS + k] >= 0 ? SLOT(Sp[G3SP_PROP_SLOTS + k]) : T(0.0);
                        sp_mort[k] = sp_mk >= 0 && Sp[G3SP_MORT_SLOTS + k] >= 0 ? SLOT(Sp[G3SP_MORT_SLOTS + k]) : T(0.0);
                        sp_p[k] = sp_kind == G3SPAWN_FECUNDITY ? SLOT(Sp[G3SP_RECR_SLOTS + k]) : T(0.0);
                    }
                }
                auto spawn_cell = [&](const int l, const int aidx, T& num, T& wgt) {
                    const double ml = R[St[G3S_MIDLEN_ROFF] + l];
                    const T spn = num * suitability_of<T>(sp_pk, sp_prop, ml, ml);
                    if (sp_kind == G3SPAWN_FECUNDITY)
                        sp_acc = sp_acc + xpow(T(ml), sp_p[1]) * xpow(T((double)(St[G3S_MINAGE] + aidx)), sp_p[2]) * xpow(spn, sp_p[3]) * xpow(wgt, sp_p[4]);
                    else sp_acc = sp_acc + wgt * spn;
                    if (sp_mk >= 0) num = num - spn * (1.0 - xexp(-suitability_of<T>(sp_mk, sp_mort, ml, ml)));
                    const int32_t* Spw = I + St[G3S_SPAWN_OFF];
                    if (Spw[G3SP_WLOSS] >= 0) {     // spawning weight loss (R/action_spawn.R:166-181, R/action_weightloss.R:19-44)
                        const T mw = SLOT(Spw[G3SP_WLOSS + 1]);
                        const T lost = ((wgt - mw) * (1.0 - SLOT(Spw[G3SP_WLOSS]))) + mw;
                        wgt = ratio_add_pop(wgt, num - spn, lost, spn);
                    }
                };
                if (!has_mort && grow_n < 0 && !inc_now && !ren_mask && xact == 0 && !spawn_here) continue;

                // ---------- natural mortality, growth (+ maturing split), hand-over, renewal, collect vectors.
                // Length growth only looks DOWN the length axis, so walking l from the top the update is in
                // place: every source l - d is still the old value when l is rebuilt. Columns are the inner
                // loop: the band P(l - d -> l) and the weight gains are loaded once per l for all columns.
                const int gsz = (grow_n + 1) * L;
                const int gset = grow_n >= 0 && St[G3S_GROW_TMAP_OFF] >= 0 ? I[St[G3S_GROW_TMAP_OFF] + t] : 0;     // time-varying growth: this step's tables
                const T wal = grow_n >= 0 ? SLOT(I[St[G3S_GROW_OFF] + 5 * gset + 3]) : T(0.0);
                const int mset = St[G3S_MORT_TMAP_OFF] >= 0 ? I[St[G3S_MORT_TMAP_OFF] + t] : 0;        // time-varying M: this step's table
                const T* pmort = ws + (unsigned)(A.t_mort[s] + (mset * A.w_ndt + idt) * ncol) * NTH;
                const int4* cinc = A.t_colinc + A.t_colbase[s];
                // Wide bands (streamed per column, below) sweep the columns in blocks of t_cblk: the n+1 rows a
                // column re-reads while l walks down then stay in L1/L2 instead of coming back from HBM
                // (65,536 threads x 19 columns x 16 rows x 2 x 8 B = 318 MB per length otherwise).
                const int cper = NW <= 8 ? ncol : A.t_cblk;
                for (int cb0 = 0; cb0 < ncol; cb0 += cper) {
                const int cb1 = min(cb0 + cper, ncol);
                const int ncset = grow_n >= 0 ? A.w_ncset[s] : 1;       // > 1: bands per column (maturity with an age term)
                const bool percol = cont_now && ncset > 1;
                const T* pg = ws + (unsigned)((cont_now ? A.t_gn[s] + (gset * ncset * A.w_ndt + idt) * gsz : A.t_g[s] + (gset * A.w_ndt + idt) * gsz) + (L - 1)) * NTH;     // staying part, [d][l]
                const T* pgr = ws + (unsigned)(A.t_gr[s] + (gset * ncset * A.w_ndt + idt) * gsz + (L - 1)) * NTH;                            // maturing part
                const unsigned cset_stride = (unsigned)(A.w_ndt * gsz) * NTH;      // from one column's bands to the next
                const T* ppw = ws + (unsigned)(A.t_pw[s] + gset * L + (L - 1)) * NTH;
                for (int l = L - 1; l >= 0; l--, pg -= NTH, pgr -= NTH, ppw -= NTH) {
                    // Narrow bands keep the band and the weight gains of this length in registers for all columns;
                    // wide ones (maxlengthgroupgrowth > 7) would spill them, so they stream them per column instead.
                    constexpr bool WINDOW = NW <= 8;
                    constexpr int NWR = WINDOW ? NW : 1;
                    T gt[NWR], grt[NWR], dwt[NWR];
                    const T pw_l = ppw[0];
                    if (l > 0 && grow_n >= 0) {     // next length's band on its way
#pragma unroll
                        for (int d = 0; d < NW; d++)
                            if (d <= grow_n && d < l) { prefetch_l1(pg - NTH + d * LWS); if (cont_now) prefetch_l1(pgr - NTH + d * LWS); }
                    }
                    if constexpr (WINDOW) {
#pragma unroll
                        for (int d = 0; d < NW; d++) {
                            gt[d] = T(0.0); grt[d] = T(0.0); dwt[d] = T(0.0);
                            if (d <= grow_n && d <= l) {
                                gt[d] = pg[d * LWS];
                                if (cont_now) grt[d] = pgr[d * LWS];
                                dwt[d] = wal * (pw_l - *(ppw - d * NTH));        // walpha (l^wbeta - (l-d)^wbeta), R/action_grow.R:58-71
                            }
                        }
                    }
                    T* pn = ws + (unsigned)(A.t_num + c0 + cb0 * L + l) * NTH;
                    T* pw_ = ws + (unsigned)(A.t_wgt + c0 + cb0 * L + l) * NTH;
                    T* ptn = ws + (unsigned)(A.t_tn + A.w_tr_off[s] + cb0 * L + l) * NTH;
                    T* ptw = ws + (unsigned)(A.t_tw + A.w_tr_off[s] + cb0 * L + l) * NTH;
                    int aidx = cb0 % nage, ridx = cb0 / nage;
                    T* const pn_first = pn; T* const pw_first = pw_;
                    for (int col = cb0; col < cb1; col++, pn += LWS, pw_ += LWS, ptn += LWS, ptw += LWS) {
                        {   // the next column's (or, at the end of the sweep, the next length's first column's) sources on their way
                            const T* qn = col + 1 < cb1 ? pn + LWS : pn_first - NTH;
                            const T* qw = col + 1 < cb1 ? pw_ + LWS : pw_first - NTH;
                            if (col + 1 < cb1 || l > 0) {
                                if (grow_n >= 0) {
#pragma unroll
                                    for (int d = 0; d < NW; d++)
                                        if (d <= grow_n && d < l) { prefetch_l1(qn - d * NTH); prefetch_l1(qw - d * NTH); }
                                } else { prefetch_l1(qn); prefetch_l1(qw); }
                            }
                        }
                        const T mortf = has_mort ? pmort[col * NTH] : T(1.0);
                        T num, wgt;
                        if (grow_n >= 0) {
                            // ---- g3a_growmature (R/action_grow.R:340-497): banded gather over the sources l - d
                            T an(0.0), aw(0.0), bn(0.0), bw(0.0);
#pragma unroll
                            for (int d = 0; d < NW; d++) {
                                if (d <= grow_n && d <= l) {
                                    const T ns = *(pn - d * NTH) * mortf;
                                    T gd, grd, dwd;
                                    if constexpr (WINDOW) { gd = gt[d]; grd = grt[d]; dwd = dwt[d]; }
                                    else {
                                        gd = pg[d * LWS]; grd = cont_now ? pgr[d * LWS] : T(0.0);
                                        dwd = wal * (pw_l - *(ppw - d * NTH));
                                    }
                                    if (percol) { gd = (pg + col * cset_stride)[d * LWS]; grd = (pgr + col * cset_stride)[d * LWS]; }
                                    const T wsrc = dwd + *(pw_ - d * NTH);
                                    if (cont_now) {
                                        const T w = wsrc * ns;
                                        an = an + grd * ns; aw = aw + grd * w;
                                        bn = bn + gd * ns; bw = bw + gd * w;
                                    } else if (CONSTMAT && const_now) {
                                        // g3a_mature_constant ratio of the SOURCE cell (R/action_mature.R:44-74, R/action_grow.R:438-457)
                                        const T ratio = W(A.t_cmat[s] + col * L + l - d);
                                        const T n1 = (ns * ratio) * gd, n2 = (ns * (1.0 - ratio)) * gd;
                                        an = an + n1; aw = aw + wsrc * n1;
                                        bn = bn + n2; bw = bw + wsrc * n2;
                                    } else {
                                        const T n2 = ns * gd;
                                        bn = bn + n2; bw = bw + wsrc * n2;
                                    }
                                }
                            }
                            num = bn; wgt = xdiv(bw, avoid_zero(bn));
                            if (spawn_here) spawn_cell(l, aidx, num, wgt);        // (before the maturing fish move: order 6 < 7)
                            if (trans_now) {
                                *ptn = an; *ptw = xdiv(aw, avoid_zero(an));
                                // unclaimed remainder moves back (move_remainder = TRUE, R/action_mature.R:126-131)
                                num = num + an * W(A.t_remf + A.t_colbase[s] + col);
                            }
                        } else {
                            num = *pn * mortf; wgt = *pw_;
                            if (spawn_here) spawn_cell(l, aidx, num, wgt);
                        }
                        // ---- g3a_step_transition (R/action_mature.R:78-134): maturing fish arriving from their source column
                        if (inc_now) {
                            const int4 ci = cinc[col];
                            if (ci.x >= 0 && ((ci.z >> step) & 1)) {
                                const T moved = W(A.t_tn + ci.x + l) * SLOT(ci.y);
                                wgt = ratio_add_pop(wgt, num, W(A.t_tw + ci.x + l), moved);
                                num = num + moved;
                            }
                        }
                        // ---- g3a_renewal (R/action_renewal.R:166-188)
                        if (ren_mask)
                            for (int r = 0; r < n_renew; r++)
                                if (((ren_mask >> r) & 1) && aidx == I[St[G3S_RENEW_OFF] + r * G3R_LEN + G3R_AGE_IDX]) {
                                    const int fs = I[I[St[G3S_RENEW_OFF] + r * G3R_LEN + G3R_FACTOR_OFF] + yidx * narea + ridx];
                                    if (fs >= 0) {
                                        const T rn = W(A.t_rd[s] + r * L + l) * SLOT(fs);
                                        if (REPORT && live) A.report[A.rep_renew[s] + ((size_t)t * ncol + col) * L + l] = val(rn);
                                        wgt = ratio_add_pop(wgt, num, W(A.t_rw[s] + r * L + l), rn);
                                        num = num + rn;
                                    }
                                }
                        *pn = num; *pw_ = wgt;
                        // ---- collect vectors for g3l_distribution (R/likelihood_distribution.R:275-287)
                        if (xact) {
                            const int cell = c0 + col * L + l;
#pragma unroll
                            for (int x = 0; x < MAXX; x++)
                                if ((xact >> x) & 1) {
                                    if (A.w_x_kind[x] == 0) { if (A.w_x_unit[x] == 0 && nq > 0) W(A.t_x[x] + cell) = xdiv(W(A.t_x[x] + cell), avoid_zero(wgt)); }
                                    else W(A.t_x[x] + cell) = A.w_x_unit[x] == 0 ? num : num * wgt;
                                }
                        }
                        if (++aidx == nage) { aidx = 0; ridx++; }
                    }
                }
                }
                if (spawn_here) W(A.t_sps + s) = sp_acc;
            }

            // ================= g3a_spawn, order 8 (R/action_spawn.R:183-218): the offspring arrive in the youngest age of the output stocks
            for (int s = 0; s < NS; s++) {
                const int32_t* St = I + I[G3H_STOCK_OFF] + s * G3S_LEN;
                if (St[G3S_SPAWN_OFF] < 0) continue;
                const int32_t* Sp = I + St[G3S_SPAWN_OFF];
                if (Sp[G3SP_RUN_STEP] != 0 && Sp[G3SP_RUN_STEP] != step + 1) continue;
                const int32_t* rs = Sp + G3SP_RECR_SLOTS;
                const int kind = Sp[G3SP_RECR_KIND];
                const T sum = W(A.t_sps + s);
                T r(0.0);
                if (kind == G3SPAWN_FECUNDITY || kind == G3SPAWN_SIMPLESSB) r = SLOT(rs[0]) * sum;
                else if (kind == G3SPAWN_RICKER) r = SLOT(rs[0]) * sum * xexp(-SLOT(rs[1]) * sum);
                else if (kind == G3SPAWN_BEVERTONHOLT) r = SLOT(rs[0]) * sum / (SLOT(rs[1]) + sum);
                else if (kind == G3SPAWN_BEVERTONHOLT_SS3) {
                    const T h = SLOT(rs[0]), B0 = SLOT(rs[1]), Ry = SLOT(I[rs[2] + yidx]);
                    r = 4.0 * h * sum * Ry / (B0 * (1.0 - h) + sum * (5.0 * h - 1.0));
                } else r = SLOT(rs[0]) * lsa_c(-1000.0 * sum / SLOT(rs[1]), -1000.0) / -1000.0;
                const double ncell = (double)(St[G3S_L] * St[G3S_NAGE] * St[G3S_NAREA]);
                const T total = r * 1.0 / avoid_zero(T(ncell)) * ncell;     // sum of stock__offspringnum = r / avoid_zero(cells) per cell
                for (int o = 0; o < Sp[G3SP_N_OUT]; o++) {
                    const int32_t* Or = I + Sp[G3SP_OUT_OFF] + 6 * o;
                    const int32_t* So = I + I[G3H_STOCK_OFF] + Or[0] * G3S_LEN;
                    const int oL = So[G3S_L], oA = So[G3S_NAGE], oR = So[G3S_NAREA];
                    const double* oml = R + So[G3S_MIDLEN_ROFF];
                    const T mean = SLOT(Or[1]), isd = 1.0 / SLOT(Or[2]), wa = SLOT(Or[3]), wb = SLOT(Or[4]);
                    T tot(0.0);
                    for (int ar = 0; ar < oR; ar++)
                        for (int l = 0; l < oL; l++) { const T z = (oml[l] - mean) * isd; tot = tot + xexp(-(z * z) * 0.5); }
                    const T share = total * SLOT(Or[5]) / avoid_zero(tot);
                    for (int ar = 0; ar < oR; ar++)
                        for (int l = 0; l < oL; l++) {
                            const int cell = So[G3S_CELL_OFF] + ar * oA * oL + l;
                            const T z = (oml[l] - mean) * isd;
                            const T spawned = xexp(-(z * z) * 0.5) * share;
                            const T n0 = W(A.t_num + cell), w0 = W(A.t_wgt + cell);
                            const T w1 = ratio_add_pop(w0, n0, wa * xpow(T(oml[l]), wb), spawned), n1 = n0 + spawned;
                            W(A.t_num + cell) = n1; W(A.t_wgt + cell) = w1;
                            // the collect vectors of these cells were filled when the stock's own step ended: bring them up to date
#pragma unroll
                            for (int x = 0; x < MAXX; x++)
                                if ((xact >> x) & 1) {
                                    if (A.w_x_kind[x] == 0) { if (A.w_x_unit[x] == 0) W(A.t_x[x] + cell) = W(A.t_x[x] + cell) * avoid_zero(w0) / avoid_zero(w1); }
                                    else W(A.t_x[x] + cell) = A.w_x_unit[x] == 0 ? n1 : n1 * w1;
                                }
                        }
                }
            }

            // ================= g3l_distribution collect: model[dest] += lgmatrix . x (R/likelihood_distribution.R:288-364)
            for (int c = 0; c < A.NCOMP; c++) {
                if (!((cact >> c) & 1)) continue;
                const int32_t* C = I + I[G3H_COMP_OFF] + c * G3C_LEN;
                const int nd = C[G3C_N_DEST];
                const bool si = C[G3C_FUNC] == G3F_SI_LOG || C[G3C_FUNC] == G3F_SI_LINEAR;
                const int ms = A.t_ms[c] + (si ? I[C[G3C_TIDX_OFF] + t] * nd : 0);
                const int xs = A.t_x[C[G3C_XBUF]];
                if (C[G3C_MODE] == 0) {
                    const int32_t* ptr = I + C[G3C_CSR_PTR_OFF];
                    const int32_t* idx = I + C[G3C_CSR_IDX_OFF];
                    const double* cf = R + C[G3C_CSR_COEF_ROFF];
                    for (int e = 0; e < nd; e++) {
                        T acc = W(ms + e);
                        for (int j = ptr[e]; j < ptr[e + 1]; j++) acc = acc + cf[j] * W(xs + idx[j]);
                        W(ms + e) = acc;
                    }
                } else {
                    const int32_t* cd = I + C[G3C_CELL_DEST_OFF];
                    const double* cf = R + C[G3C_CELL_COEF_ROFF];
                    if (nd == 1) {
                        T acc = W(ms);
                        for (int i = 0; i < A.NC; i++) if (cd[i] == 0) acc = acc + cf[i] * W(xs + i);
                        W(ms) = acc;
                    } else {
                        for (int i = 0; i < A.NC; i++) if (cd[i] >= 0) W(ms + cd[i]) = W(ms + cd[i]) + cf[i] * W(xs + i);
                    }
                }
                if (REPORT && live)
                    for (int e = 0; e < nd; e++)
                        A.report[A.rep_model[c] + (size_t)I[C[G3C_TIDX_OFF] + t] * nd + e] = val(W(ms + e));
            }

            // ================= g3l_distribution compare (R/likelihood_distribution.R:375-394)
            for (int c = 0; c < A.NCOMP; c++) {
                if (!((cact >> c) & 1)) continue;
                const int32_t* C = I + I[G3H_COMP_OFF] + c * G3C_LEN;
                if (!I[C[G3C_CMP_OFF] + t]) continue;
                const int ti = I[C[G3C_TIDX_OFF] + t];
                const int func = C[G3C_FUNC], nd = C[G3C_N_DEST], nt = C[G3C_N_TIMES];
                const bool si = func == G3F_SI_LOG || func == G3F_SI_LINEAR;
                const int ms = A.t_ms[c] + (si ? ti * nd : 0);
                const T weight = mkpar_g<T>(prow, seed, C[G3C_WEIGHT_PAR]);
                const int nb = C[G3C_N_BINS], nsl = C[G3C_N_SLICES], nag = C[G3C_N_AREAG];
                T cnll(0.0);
                bool scored = false;
                if (func == G3F_SUMOFSQUARES && R[C[G3C_PARAM_ROFF]] != 0.0) {
                    // stratified (over includes 'length'): proportions within each length group, R/likelihood_distribution.R:8-16
                    const double* op = A.obsprop + A.obs_off[c] + (size_t)ti * nd;     // (obs / avoid_zero(its length group's total))
                    for (int ag = 0; ag < nag; ag++)
                        for (int b = 0; b < nb; b++) {
                            T tot(0.0);
                            for (int sl = 0; sl < nsl; sl++) tot = tot + W(ms + (ag * nsl + sl) * nb + b);
                            const T rt = 1.0 / avoid_zero(tot);
                            for (int sl = 0; sl < nsl; sl++) {
                                const int i = (ag * nsl + sl) * nb + b;
                                T dlt = W(ms + i) * rt - op[i];
                                cnll = cnll + dlt * dlt;
                            }
                        }
                    scored = true;
                } else if (func == G3F_SUMOFSQUARES) {     // R/likelihood_distribution.R:3-39
                    const int per = nsl * nb;
                    const double* op = A.obsprop + A.obs_off[c] + (size_t)ti * nd;
                    for (int ag = 0; ag < nag; ag++) {
                        T tot(0.0);
                        for (int e = 0; e < per; e++) tot = tot + W(ms + ag * per + e);
                        const T rt = 1.0 / avoid_zero(tot);
                        for (int sl = 0; sl < nsl; sl++) {
                            T v(0.0);
                            for (int e = 0; e < nb; e++) {
                                const int i = ag * per + sl * nb + e;
                                T dlt = W(ms + i) * rt - op[i];
                                v = v + dlt * dlt;
                            }
                            cnll = cnll + v;
                        }
                    }
                    scored = true;
                } else if (func == G3F_MULTINOMIAL) {   // R/likelihood_distribution.R:41-60
                    const double eps = R[C[G3C_PARAM_ROFF]];
                    const double* op = A.obsprop + A.obs_off[c] + (size_t)ti * nd;
                    T likely(0.0);
                    for (int sl = 0; sl < nag * nsl; sl++) {
                        T tot(0.0);
                        for (int e = 0; e < nb; e++) tot = tot + W(ms + sl * nb + e);
                        const T at = avoid_zero(tot);
                        for (int e = 0; e < nb; e++)
                            likely = likely + op[sl * nb + e] * xlog(dif_pmax_c(W(ms + sl * nb + e) / at, 1.0 / (nb * eps), 1e4));
                    }
                    cnll = 2.0 * (-likely + A.obsconst[A.obsconst_off[c] + ti]);
                    scored = true;
                } else if (func == G3F_SUMSQLOGRATIOS) {    // R/likelihood_distribution.R:127-136 (obsprop holds log(obs + eps))
                    const double eps = R[C[G3C_PARAM_ROFF]];
                    const double* op = A.obsprop + A.obs_off[c] + (size_t)ti * nd;
                    for (int e = 0; e < nd; e++) {
                        const T dlt = op[e] - xlog(W(ms + e) + eps);
                        cnll = cnll + dlt * dlt;
                    }
                    scored = true;
                } else if (ti == nt - 1) {              // surveyindices: R/likelihood_distribution.R:82-125
                    const double fa = R[C[G3C_PARAM_ROFF]], fb = R[C[G3C_PARAM_ROFF] + 1];
                    const int series = A.t_ms[c];
                    const double* op = A.obsprop + A.obs_off[c];
                    for (int e = 0; e < nd; e++) {
                        T sN(0.0); double sI = 0.0;
                        for (int i = 0; i < nt; i++) {
                            T mv = W(series + i * nd + e);
                            sN = sN + (func == G3F_SI_LOG ? xlog(avoid_zero(mv)) : mv);
                            sI += op[(size_t)i * nd + e];
                        }
                        const T mN = sN / (double)nt;
                        const double mI = sI / (double)nt;
                        T beta(fb), alpha(fa);
                        if (isnan(fb)) {
                            T sxy(0.0), sxx(0.0);
                            for (int i = 0; i < nt; i++) {
                                T mv = W(series + i * nd + e);
                                T N = func == G3F_SI_LOG ? xlog(avoid_zero(mv)) : mv;
                                sxy = sxy + (op[(size_t)i * nd + e] - mI) * (N - mN); sxx = sxx + (N - mN) * (N - mN);
                            }
                            beta = sxy / avoid_zero(sxx);
                        }
                        if (isnan(fa)) alpha = mI - beta * mN;
                        if (REPORT && live) { A.report[A.rep_params + 2 * c] = val(alpha); A.report[A.rep_params + 2 * c + 1] = val(beta); }
                        for (int i = 0; i < nt; i++) {
                            T mv = W(series + i * nd + e);
                            T N = func == G3F_SI_LOG ? xlog(avoid_zero(mv)) : mv;
                            T r = alpha + beta * N - op[(size_t)i * nd + e];
                            cnll = cnll + r * r;
                        }
                    }
                    scored = true;
                }
                if (scored) {
                    nll = nll + weight * cnll;
                    W(A.t_ctot + c) = W(A.t_ctot + c) + cnll;
                    if (REPORT && live) A.report[1 + (size_t)c * NT + t] = val(cnll);
                }
                if (!si) for (int e = 0; e < nd; e++) W(ms + e) = T(0.0);     // zero counters for the next period
            }
            // ================= g3l_random_dnorm / g3l_random_walk (R/likelihood_random.R:14-109)
            for (int r = 0; r < I[G3H_NRAND]; r++) {
                const int32_t* Rn = I + I[G3H_RAND_OFF] + r * G3RN_LEN;
                const int run = I[Rn[G3RN_RUN_OFF] + t];
                if (!run || (run == 2 && t != total_steps)) continue;
                const T sd = SLOT(I[Rn[G3RN_SIGMA_OFF] + t]);
                const T resid = (SLOT(I[Rn[G3RN_PARAM_OFF] + t]) - SLOT(I[Rn[G3RN_MEAN_OFF] + t])) / sd;
                const T logans = T(-0.91893853320467274178) - xlog(sd) - 0.5 * resid * resid;     // TMB dnorm
                const T n = Rn[G3RN_LOG] ? T(0.0) - logans : xexp(logans);
                nll = nll + SLOT(Rn[G3RN_WEIGHT]) * n;
                W(A.t_ctot + A.NCOMP + r) = W(A.t_ctot + A.NCOMP + r) + n;
                if (REPORT && live) A.report[1 + (size_t)(A.NCOMP + r) * NT + t] = val(n);
            }
            // ================= g3l_understocking (R/likelihood_understocking.R:36-46)
            if (have_us) {
                T u = (us_power == 2.0) ? us_tot * us_tot : xpow(us_tot, T(us_power));
                nll = nll + us_weight * u;
                W(A.t_ctot + A.NNLL - 2) = W(A.t_ctot + A.NNLL - 2) + u;
                if (REPORT && live) A.report[1 + (size_t)(A.NNLL - 2) * NT + t] = val(u);
            }

            // ================= g3a_age (R/action_age.R:51-85) + movement hand-off
            if (final_step) {
                for (int s = 0; s < NS; s++) {
                    const int32_t* St = I + I[G3H_STOCK_OFF] + s * G3S_LEN;
                    if (!St[G3S_AGE_ENABLE]) continue;
                    const int L = St[G3S_L], nage = St[G3S_NAGE], narea = St[G3S_NAREA], c0 = St[G3S_CELL_OFF];
                    const int age_n_out = St[G3S_AGE_N_OUT];
                    for (int r = 0; r < narea; r++)
                        for (int l = 0; l < L; l++) {
                            const int i = r * L + l;
                            const int base = c0 + r * nage * L + l;
                            const int top = base + (nage - 1) * L;
                            if (age_n_out > 0) { W(A.t_mv + A.w_mv_off[s] + i) = W(A.t_num + top); W(A.t_mv + A.w_mv_off[s] + narea * L + i) = W(A.t_wgt + top); }
                            if (nage == 1) {
                                if (age_n_out > 0) W(A.t_num + top) = T(0.0);
                                continue;
                            }
                            if (age_n_out > 0) { W(A.t_num + top) = W(A.t_num + top - L); W(A.t_wgt + top) = W(A.t_wgt + top - L); }
                            else {
                                const T n0 = W(A.t_num + top), n1 = W(A.t_num + top - L);
                                W(A.t_wgt + top) = ratio_add_pop(W(A.t_wgt + top), n0, W(A.t_wgt + top - L), n1);
                                W(A.t_num + top) = n0 + n1;
                            }
                            for (int a = nage - 2; a >= 1; a--) {
                                W(A.t_num + base + a * L) = W(A.t_num + base + (a - 1) * L);
                                W(A.t_wgt + base + a * L) = W(A.t_wgt + base + (a - 1) * L);
                            }
                            W(A.t_num + base) = T(0.0);
                        }
                }
                for (int i = 0; i < A.w_n_agedst; i++) {
                    const int4 e = A.w_agedst[i];
                    const T n0 = W(A.t_num + e.x);
                    const T moved = W(A.t_mv + e.y) * SLOT(e.w);
                    W(A.t_wgt + e.x) = ratio_add_pop(W(A.t_wgt + e.x), n0, W(A.t_mv + e.z), moved);
                    W(A.t_num + e.x) = n0 + moved;
                }
            }
        }   // time loop

        const T out = bad_time ? T(nan("")) : nll;
        if (!live) continue;
        if (REPORT) A.report[0] = val(out);
        if (REPORT && live)         // suit_<prey>_<pred>__report (R/action_predate.R:292-299): [predator length][prey length]
            for (int q = 0; q < A.NPP; q++) {
                const int32_t* Q = I + I[G3H_PP_OFF] + q * G3PP_LEN;
                const int32_t* P = I + I[G3H_PRED_OFF] + Q[G3PP_PRED] * G3P_LEN;
                const int n = I[I[G3H_STOCK_OFF] + Q[G3PP_PREY] * G3S_LEN + G3S_L] *
                              (P[G3P_KIND] == G3PRED_STOCK ? I[I[G3H_STOCK_OFF] + P[G3P_STOCK] * G3S_LEN + G3S_L] : 1);
                if (Q[G3PP_PREF] >= 0) continue;        // (written when the table was filled: the table is not S itself)
                for (int i = 0; i < n; i++) A.report[A.rep_pp[q] + i] = val(W(A.t_suitq[q] + suit_set(Q, min(total_steps, NT - 1)) * n + i));
            }
        if (tile_t == 0) {
            A.nll[b] = val(out);
            if (A.comp) for (int c = 0; c < A.NNLL; c++) A.comp[b * A.NNLL + c] = bad_time ? nan("") : val(W(A.t_ctot + c));
        }
        if (NTAN > 0 && A.grad)
            for (int k = 0; k < NTAN; k++)
                if (tile_t * NTAN + k < A.K) A.grad[b * A.K + tile_t * NTAN + k] = tan_of(out, k);
    }
#undef W
#undef EOT
#undef SLOT
}

}  // namespace g3
