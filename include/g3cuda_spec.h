/*
 * g3cuda_spec.h -- layout of the packed model specification ("the IR").
 *
 * A gadget3 model (reference: the action list handed to g3_to_tmb(),
 * /root/reference R/run_tmb.R:709) is lowered on the host into two flat
 * arrays:
 *     int32_t ints[]   structure, index tables, slot programs
 *     double  reals[]  constants, length grids, landings, observations
 * ints[0 .. G3H_LEN) is a header of counts and section offsets; every
 * "*_OFF" value is an offset into ints[] unless its name ends in "_ROFF",
 * in which case it is an offset into reals[].
 *
 * This file is the single source of truth for the layout. It contains only
 * "#define NAME integer" lines so that the Python packer
 * (gadget3_b200/spec.py) can parse it, and it is shared by the CUDA product
 * (gadget3_b200/csrc) and by the CPU oracle (oracle/) -- it is the
 * interface between them, not an implementation.
 *
 * Canonical device layouts (independent of the model's own dimension order,
 * which varies: baseline/introduction-single-stock.cpp:470 has
 * length,area,age; baseline/multiple-substocks.cpp:571 has length,age,area):
 *   stock cell     c = cell_off + (area_idx * nage + age_idx) * L + l
 *   likelihood dest e = (areagroup * n_slices + slice) * n_bins + bin
 *                   slice = stockgroup * n_agegroups + agegroup
 */
#ifndef G3CUDA_SPEC_H
#define G3CUDA_SPEC_H

#define G3CUDA_SPEC_VERSION 8

/* ---- header: ints[0..G3H_LEN) ------------------------------------------ */
#define G3H_VERSION        0
#define G3H_START_YEAR     1   /* g3a_time start_year (R/action_time.R:21) */
#define G3H_END_YEAR       2
#define G3H_NSTEPS         3   /* steps per year, length(step_lengths) */
#define G3H_STEPLEN_OFF    4   /* ints[NSTEPS]: months per step */
#define G3H_NT             5   /* allocated model steps = NSTEPS*(end-start+1) */
#define G3H_NPAR           6   /* length of the full PARAMETER vector */
#define G3H_PAR_RETRO      7   /* index of retro_years in it, or -1 */
#define G3H_PAR_PROJECT    8   /* index of project_years, or -1 */
#define G3H_NSTOCK         9
#define G3H_STOCK_OFF     10   /* NSTOCK records of G3S_LEN ints */
#define G3H_NCELL         11   /* total cells over all stocks */
#define G3H_NSLOT         12   /* scalar slots (parameter formulas) */
#define G3H_SLOT_OFF      13   /* ints[2*NSLOT]: (code_off, code_len) */
#define G3H_NPRED         14   /* fleets + stock predators, run order */
#define G3H_PRED_OFF      15   /* NPRED records of G3P_LEN ints */
#define G3H_NPP           16   /* predator-prey pairs, run order */
#define G3H_PP_OFF        17   /* NPP records of G3PP_LEN ints */
#define G3H_NBOUND        18   /* g3l_bounds_penalty entries, run order */
#define G3H_BOUND_OFF     19   /* ints[NBOUND]: parameter index */
#define G3H_BOUND_ROFF    20   /* reals[2*NBOUND]: (lower, upper) */
#define G3H_BOUND_WS_ROFF 21   /* reals[2]: weight, scale (1, 1e6) */
#define G3H_NCOMP         22   /* g3l_distribution components, run order */
#define G3H_COMP_OFF      23   /* NCOMP records of G3C_LEN ints */
#define G3H_NXBUF         24   /* distinct collect vectors */
#define G3H_XBUF_OFF      25   /* NXBUF records of G3X_LEN ints */
#define G3H_US_N          26   /* g3l_understocking: number of prey stocks (0 = absent) */
#define G3H_US_OFF        27   /* ints[US_N]: stock indices */
#define G3H_US_ROFF       28   /* reals[2]: power, weight */
#define G3H_NAREA         29   /* number of model areas */
#define G3H_MAXL          30   /* max L over stocks */
#define G3H_MAXN          31   /* max maxlengthgroupgrowth over stocks */
#define G3H_NNLL          32   /* number of nll report rows = NCOMP + NRAND + 2: components, random terms, understocking, bounds */
#define G3H_NRAND         33   /* g3l_random_dnorm / g3l_random_walk terms, run order (R/likelihood_random.R:14-109) */
#define G3H_RAND_OFF      34   /* NRAND records of G3RN_LEN ints */
#define G3H_LEN           40

/* ---- stock record ------------------------------------------------------ */
#define G3S_L              0   /* number of length groups */
#define G3S_MINAGE         1
#define G3S_NAGE           2
#define G3S_NAREA          3
#define G3S_AREA_OFF       4   /* ints[NAREA]: model area ids */
#define G3S_CELL_OFF       5
#define G3S_MIDLEN_ROFF    6   /* reals[L] (R/stock.R:105) */
#define G3S_PLUSDL_ROFF    7   /* reals[1] (R/stock.R:72-85) */
#define G3S_INIT_OFF       8   /* ints[NAREA*NAGE*5]: slots mean,sd,factor,walpha,wbeta per column; -1 = no initial conditions */
#define G3S_MORT_OFF       9   /* ints[NAREA*NAGE]: slot of M per column; -1 = no natural mortality */
#define G3S_GROW_N        10   /* maxlengthgroupgrowth; -1 = no growth */
#define G3S_GROW_OFF      11   /* ints[5]: slots linf, kappa, beta, walpha, wbeta */
#define G3S_MAT_KIND      12   /* G3MAT_* */
#define G3S_MAT_OFF       13   /* ints[NSET*4]: slots alpha, l50, beta, a50 (-1 = term absent), one row per set of GROW_TMAP_OFF
                                  (g3a_mature_constant: one set only) */
#define G3S_TRANS_MASK    14   /* bit (cur_step-1) set: transition_f true on that step */
#define G3S_N_OUT         15   /* maturation output stocks */
#define G3S_OUT_OFF       16   /* ints[2*N_OUT]: (stock index, ratio slot) */
#define G3S_AGE_ENABLE    17   /* g3a_age present */
#define G3S_AGE_N_OUT     18
#define G3S_AGE_OUT_OFF   19   /* ints[2*AGE_N_OUT]: (stock index, ratio slot) */
#define G3S_N_RENEW       20
#define G3S_RENEW_OFF     21   /* N_RENEW records of G3R_LEN ints */
#define G3S_MIGRATE_OFF   22   /* ints[NSTEPS*NAREA*NAREA]: slot of migrate_f[step][dest][src] or -1; -1 = no migration */
#define G3S_DIMORDER      23   /* report order: 0 = length,age,area ; 1 = length,area,age */
#define G3S_OUT_MAP_OFF   24   /* ints[N_OUT*ncell]: per output stock, destination global cell of each
                                  local cell (same length, age, area) or -1 (R/action_mature.R:100-112) */
#define G3S_AGE_MAP_OFF   25   /* ints[AGE_N_OUT*NAREA*L]: destination global cell (age maxage+1) of the
                                  oldest column per (area idx, l) or -1 (R/action_age.R:16-20,57-63) */
#define G3S_OTHERFOOD_OFF 26   /* g3a_otherfood (R/action_renewal.R:276-308): ints[NT*NAREA*2]: (num slot, wgt slot) per (step t, area idx),
                                  assigned to every cell of the stock at the start of step t; -1 = not an otherfood stock */
#define G3S_SPAWN_OFF     27   /* g3a_spawn (R/action_spawn.R:86-233): ints[G3SP_LEN] record of the PARENT stock, or -1 */
#define G3S_MORT_TMAP_OFF 28   /* time-varying natural mortality (by_year / by_step tables, g3_timevariable; R/params.R:122-129,
                                  R/timevariable.R:1-31): ints[NT]: which set of MORT_OFF's ints[NSET*NAGE*NAREA] slots step t uses;
                                  -1 = one set for every step */
#define G3S_GROW_TMAP_OFF 29   /* time-varying growth parameters (Linf, K, bbin, walpha, wbeta by year / step, g3_timevariable): ints[NT]:
                                  which set of GROW_OFF's ints[NSET*5] (and MAT_OFF's ints[NSET*4]) slots step t uses; -1 = one set for every step. The reference
                                  recomputes its growth matrices whenever their inputs change (R/action_grow.R:391-410) */
#define G3S_LEN           30

/* ---- spawning record (g3a_spawn, R/action_spawn.R:86-233) ----------------
   At order 6 (after growth, before the maturing fish arrive and before renewal): spawningnum = num * proportion(l);
   s = sum over the whole stock of the recruitment function's summand; parents lose spawningnum * (1 - exp(-mortality(l)));
   r = R(s). At order 8: every output stock gets r * ratio fish at its youngest age, spread over all its areas and length
   groups as exp(-((midlen - mean) / sd)^2 / 2) / avoid_zero(sum of that over the areas and lengths), mean weight
   alpha * midlen^beta, combined with ratio_add_pop. */
#define G3SP_RUN_STEP      0   /* cur_step == this (1-based); 0 = every step */
#define G3SP_RECR_KIND     1   /* G3SPAWN_* */
#define G3SP_RECR_SLOTS    2   /* ints[5]: fecundity p0..p4; simplessb mu; ricker / bevertonholt mu, lambda;
                                  bevertonholt_ss3 h, B0, then the offset of ints[NYEARS] slots of R by year idx; hockeystick r0, blim */
#define G3SP_PROP_KIND     7   /* proportion_f: G3SUIT_* of the parent's length (G3SUIT_CONSTANT: one formula) */
#define G3SP_PROP_SLOTS    8   /* ints[5] */
#define G3SP_MORT_KIND    13   /* mortality_f likewise; -1 = no spawning mortality */
#define G3SP_MORT_SLOTS   14   /* ints[5] */
#define G3SP_N_OUT        19
#define G3SP_OUT_OFF      20   /* ints[N_OUT*6]: (output stock, mean slot, sd slot, alpha slot, beta slot, ratio slot) */
#define G3SP_WLOSS        21   /* spawning weight loss (g3a_weightloss on the spawning part, R/action_spawn.R:166-181, R/action_weightloss.R:19-44):
                                  ints[2]: slots rel_loss, min_weight; after the spawning mortality
                                  wgt = ratio_add_pop(wgt, num - spawningnum, (wgt - min_weight) (1 - rel_loss) + min_weight, spawningnum);
                                  -1 in [0] = none */
#define G3SP_LEN          23

#define G3SPAWN_FECUNDITY        0   /* s = sum midlen^p1 age^p2 spawningnum^p3 wgt^p4, r = p0 s (R/action_spawn.R:1-17) */
#define G3SPAWN_SIMPLESSB        1   /* s = sum wgt spawningnum (this and all below), r = mu s (:19-28) */
#define G3SPAWN_RICKER           2   /* r = mu s exp(-lambda s) (:32-43) */
#define G3SPAWN_BEVERTONHOLT     3   /* r = mu s / (lambda + s) (:45-56) */
#define G3SPAWN_BEVERTONHOLT_SS3 4   /* r = 4 h s R / (B0 (1 - h) + s (5 h - 1)) (:59-74) */
#define G3SPAWN_HOCKEYSTICK      5   /* r = r0 logspace_add(-1000 s / blim, -1000) / -1000 (:76-88) */

#define G3MAT_NONE         0
#define G3MAT_CONSTANT     1   /* g3a_mature_constant (R/action_mature.R:44) */
#define G3MAT_CONTINUOUS   2   /* g3a_mature_continuous (R/action_mature.R:1) */

/* ---- renewal record (g3a_renewal_normalparam, R/action_renewal.R:190) --- */
#define G3R_RUN_STEP       0   /* cur_step == this (1-based); 0 = every step */
#define G3R_AGE_IDX        1   /* age index renewed into; -1 = every age */
#define G3R_MEAN           2   /* slot */
#define G3R_SD             3
#define G3R_WALPHA         4
#define G3R_WBETA          5
#define G3R_FACTOR_OFF     6   /* ints[NYEARS*NAREA]: factor slot per (year idx, area idx); -1 = skip */
#define G3R_LEN            7

/* ---- predator record --------------------------------------------------- */
#define G3P_KIND           0   /* G3PRED_* */
#define G3P_NAREA          1
#define G3P_AREA_OFF       2   /* ints[NAREA]: model area ids */
#define G3P_E_ROFF         3   /* fleets: reals[NT*NAREA] landings per (t, area idx) (g3_timeareadata, R/timedata.R:89) */
#define G3P_N_PP           4
#define G3P_PP_FIRST       5   /* index of first pp record */
#define G3P_STOCK          6   /* stock predators: predator stock index */
#define G3P_SLOT_OFF       7   /* stock predators: ints[6]: slots m0,m1,m2,m3,halffeeding,temperature */
#define G3P_LEN            8

#define G3PRED_TOTALFLEET  0   /* g3a_predate_catchability_totalfleet (R/action_predate.R:1) */
#define G3PRED_NUMBERFLEET 1   /* suit = S * num; cons = E * (suit / total) * wgt (R/action_predate.R:7-11) */
#define G3PRED_STOCK       2   /* g3a_predate_catchability_predator (R/action_predate.R:207) */
#define G3PRED_LINEARFLEET 3   /* cons = suit * E * cur_step_size (R/action_predate.R:13-18) */
#define G3PRED_EFFORTFLEET 4   /* cons = catchability_fs[prey] * E * cur_step_size * suit (R/action_predate.R:117-126) */

/* ---- predator-prey pair record ------------------------------------------ */
#define G3PP_PRED          0
#define G3PP_PREY          1   /* stock index */
#define G3PP_SUIT_KIND     2   /* G3SUIT_* */
#define G3PP_SUIT_OFF      3   /* ints[6]: slots */
#define G3PP_ENERGY        4   /* stock predators: slot energycontent; effort fleets: slot catchability_fs of this prey; else -1 */
#define G3PP_PREF          5   /* stock predators: slot prey_preference, else -1 */
#define G3PP_SUIT_TMAP     6   /* time-varying suitability parameters (selectivity blocks: by_year / by_step tables, g3_timevariable):
                                  ints[NT]: which set of SUIT_OFF's ints[NSET*6] slots step t uses; -1 = one set for every step */
#define G3PP_LEN           7

#define G3SUIT_EXPL50      0   /* alpha, l50 (R/suitability.R:5) */
#define G3SUIT_ANDERSEN    1   /* p0..p4, p5 = predator length (R/suitability.R:15) */
#define G3SUIT_CONSTANT    2   /* value */
#define G3SUIT_ANDERSENFLEET 3 /* p0..p4, p5 = the prey's largest midlen, no avoid_zero (R/suitability.R:32-58) */
#define G3SUIT_GAMMA       4   /* alpha, beta, gamma (R/suitability.R:60-67) */
#define G3SUIT_EXPONENTIAL 5   /* alpha, beta, gamma, delta; uses predator length (R/suitability.R:69-75) */
#define G3SUIT_STRAIGHTLINE 6  /* alpha, beta (R/suitability.R:77-79) */
#define G3SUIT_RICHARDS    7   /* p0..p3 as exponential, ^(1/p4) (R/suitability.R:90-94) */

/* ---- collect vector ("xbuf") record ------------------------------------- */
#define G3X_KIND           0   /* 0 = catch (predprey cons), 1 = abundance */
#define G3X_UNIT           1   /* 0 = number, 1 = weight (R/likelihood_distribution.R:275-287) */
#define G3X_N_PP           2   /* catch: contributing pp records */
#define G3X_PP_OFF         3   /* ints[N_PP] */
#define G3X_ACTIVE_OFF     4   /* ints[NT]: needed at step t */
#define G3X_LEN            5

/* ---- likelihood component record ---------------------------------------- */
#define G3C_FUNC           0   /* G3F_* */
#define G3C_WEIGHT_PAR     1   /* parameter index of <name>_weight */
#define G3C_XBUF           2
#define G3C_MODE           3   /* 0 = gather (CSR rows per dest), 1 = reduce (dest per cell) */
#define G3C_N_DEST         4
#define G3C_N_BINS         5
#define G3C_N_SLICES       6
#define G3C_N_AREAG        7
#define G3C_CSR_PTR_OFF    8   /* gather: ints[N_DEST+1] */
#define G3C_CSR_IDX_OFF    9   /* gather: ints[nnz] cell index */
#define G3C_CSR_COEF_ROFF 10   /* gather: reals[nnz] lgmatrix coefficient (R/step.R:272-282) */
#define G3C_CELL_DEST_OFF 11   /* reduce: ints[NCELL] dest or -1 */
#define G3C_CELL_COEF_ROFF 12  /* reduce: reals[NCELL] */
#define G3C_TIDX_OFF      13   /* ints[NT]: obs/model time index at step t or -1 (R/stock_time.R:46-57) */
#define G3C_CMP_OFF       14   /* ints[NT]: done_aggregating_f at step t (R/likelihood_data.R:219) */
#define G3C_N_TIMES       15
#define G3C_OBS_ROFF      16   /* reals[N_TIMES*N_DEST] observations, canonical dest layout */
#define G3C_PARAM_ROFF    17   /* reals[2]: surveyindices fixed alpha, beta (NaN = estimate); multinomial, sumofsquaredlogratios: epsilon;
                                  sumofsquares: [0] != 0 = stratified (proportions within each length group, over = 'length') */
#define G3C_NAREA_SEL_OFF 18   /* reserved */
#define G3C_LEN           19

#define G3F_SUMOFSQUARES   0   /* R/likelihood_distribution.R:31 */
#define G3F_MULTINOMIAL    1   /* R/likelihood_distribution.R:41 */
#define G3F_SI_LOG         2   /* R/likelihood_distribution.R:82 (fit = 'log') */
#define G3F_SI_LINEAR      3
#define G3F_SUMSQLOGRATIOS 4   /* sum((log(obs + eps) - log(model + eps))^2), R/likelihood_distribution.R:127-136; PARAM[0] = eps */

/* ---- random-effect penalty record (g3l_random_dnorm, g3l_random_walk; R/likelihood_random.R:14-109) ----
   On every step t with RUN[t] != 0:  n = sense * dnorm(slot PARAM[t], slot MEAN[t], slot SIGMA[t], log),
   sense = -1 with LOG (the log density) else +1;  nll += slot WEIGHT * n;  report row NCOMP + r += n.
   A random walk is the same record with MEAN[t] = the PARAM slot of the previous run step (its first run step only
   remembers the value: RUN = 0 there). */
#define G3RN_RUN_OFF       0   /* ints[NT]: 1 = the term is added on step t (period 'step': every step; 'year': the year's last
                                  step); 2 = added on step t if t is the last step of this parameter vector's run, cur_time ==
                                  total_steps (period 'single': it moves with retro_years) */
#define G3RN_PARAM_OFF     1   /* ints[NT]: slot of param_f on step t */
#define G3RN_MEAN_OFF      2   /* ints[NT]: slot of mean_f on step t (walk: of param_f on the previous run step) */
#define G3RN_SIGMA_OFF     3   /* ints[NT]: slot of sigma_f on step t */
#define G3RN_WEIGHT        4   /* slot of the weight formula (<nll_name>_weight) */
#define G3RN_LOG           5   /* log_f */
#define G3RN_LEN           6

/* ---- slot programs: postfix code, one int per op, operands inline ------- */
#define G3OP_CONST         1   /* next int: reals[] offset */
#define G3OP_PARAM         2   /* next int: parameter index */
#define G3OP_ADD           3
#define G3OP_SUB           4
#define G3OP_MUL           5
#define G3OP_DIV           6
#define G3OP_NEG           7
#define G3OP_EXP           8
#define G3OP_LOG           9
#define G3OP_POW          10
#define G3OP_SQRT         11
#define G3OP_AVOID_ZERO   12   /* R/aab_env.R:24 */
#define G3OP_NAN          13   /* missing g3_param_table entry (R/run_tmb.R:1047-1054) */
#define G3_SLOT_STACK     16   /* max evaluation stack depth */

#endif /* G3CUDA_SPEC_H */
