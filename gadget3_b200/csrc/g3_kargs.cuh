// g3_kargs.cuh -- launch arguments of the thread-per-evaluation kernel (g3_thread_kernel.cuh) and the
// compile-time capacities of its register arrays. (The tile kernel, g3_tile.cuh, has its own record tables and
// no such capacities.)
#pragma once
#include "g3_common.cuh"

namespace g3 {

constexpr int MAXS = 16;     // stocks
constexpr int MAXC = 32;     // likelihood components
constexpr int MAXQ = 8;      // predator-prey pairs acting on one stock
constexpr int MAXTS = 8;     // (predator, area) total-suitability accumulators
constexpr int MAXX = 12;     // collect vectors
constexpr int MAXRN = 30;    // renewal records per stock (a bit each in the per-step mask)
constexpr int MAXMIG = 4;    // areas a migrating stock may live on
constexpr int TPB = 512;        // block-size cap of the default variant (one block per SM; the launch picks
                                // blockDim <= the cap so that the batch splits into equally full waves)
constexpr int TPB_BIG = 1024;   // cap of the high-occupancy variant (64 registers per thread), used for plain-double
                                // batches larger than one 512-thread wave: more warps hide more latency
constexpr unsigned WS_STRIDE = 152u * 1024u;    // workspace stride in threads: room for 152 SMs x the largest block

struct KArgs {
    const int32_t* I; const double* R;
    const int32_t* rem_off;      // per cell: offset into rem_slots list of ratio slots with a destination
    const int32_t* rem_slots;    //   [-1 terminated]
    const int32_t* pp_tsid;      // [NPP*maxnarea] total-suitability accumulator of (pp, prey area idx) or -1
    const double* obsprop;       // per component: obs / avoid_zero(total) (sumofsquares), raw obs (multinomial), log(avoid_zero(obs)) (SI)
    const double* obsconst;      // per component/time: multinomial constant term
    const double* par; long long B;
    const int32_t* tangent_idx; int K;
    double* nll; double* comp; double* grad; double* report;
    int NC, NS, NPP, NTS, NCOMP, NX, NSLOT, NPAR, NT, NSTEPS, NNLL, NBOUND, maxnarea, any_const_mat;
    int stock_pp_off[MAXS + 1]; int stock_pp[MAXS * MAXQ];
    int us_flag[MAXS];
    int obs_off[MAXC]; int obsconst_off[MAXC];
    long long rep_dstart[MAXS]; long long rep_model[MAXC]; long long rep_params;
    long long rep_renew[MAXS]; long long rep_pp[32];   // detail section: renewalnum per stock; per pair: suit report, suit, cons
    int w_ndt, w_dt_len[4], w_step_idt[32]; // distinct step lengths (months) and the one each step uses
    const unsigned char* w_pred_active;     // [NT] some landing is non-zero at t
    int w_tr_off[MAXS], w_mv_off[MAXS], w_inc_steps[MAXS];
    int w_ts_eoff[MAXTS], w_ts_estr[MAXTS]; // landings table of each (predator, area) accumulator
    int t_cblk;                             // wide bands: columns per growth sweep block
    int w_ts_kind[MAXTS];                   // G3PRED_* of the accumulator's fleet
    int w_x_kind[MAXX], w_x_unit[MAXX]; unsigned w_x_ppmask[MAXX];
    int w_n_agedst; const int4* w_agedst;   // ageing movement: (destination cell, stash index num, stash index wgt, ratio slot)
    // workspace entries, [entry][thread]
    void* t_ws;
    int t_slot, t_num, t_wgt, t_tn, t_tw, t_mv, t_suit, t_scr, t_ctot;
    int t_mort[MAXS], t_pw[MAXS], t_g[MAXS], t_gr[MAXS], t_gn[MAXS], t_rd[MAXS], t_rw[MAXS], t_ms[MAXC];
    int t_x[MAXX];                          // collect vectors
    int w_suit_nset[32];                    // per pair: sets of suitability parameters (1 unless the selectivity varies with time)
    int t_suitq[32];                        // per predator-prey pair: suitability [set][predator length][prey length]
    int t_pts[MAXS], t_pmc[MAXS];           // per stock predator: totals / consumption factors [area][length]; max consumption [length]
    int t_mig[MAXS], t_mig_steps[MAXS];     // migration matrices [step][dest][src]; steps on which the stock migrates
    int w_grow_nset[MAXS];                  // slot sets of time-varying growth parameters (1 otherwise)
    int w_ncset[MAXS];                      // maturing / staying bands per column (g3a_mature_continuous with an age term), else 1
    int w_mort_nset[MAXS];                  // slot sets of a time-varying M (1 otherwise)
    int t_sps;                              // per stock: the recruitment sum s of a spawning stock at this step (g3a_spawn)
    int t_remf, t_colbase[MAXS], t_cmat[MAXS];      // per global column: unclaimed share; first global column; constant-maturity ratios
    const int4* t_colinc;                   // per global column: (compact index of the source column's first cell or -1, ratio slot, step mask, 0)
};

}  // namespace g3
