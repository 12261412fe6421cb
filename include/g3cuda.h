/*
 * g3cuda.h -- C ABI of the B200 batched objective evaluator for gadget3 models.
 *
 * This is the drop-in boundary: what a reference-side binding (R .Call shim,
 * see INTEGRATION.md) links against in place of the TMB-compiled objective.
 * Every entry point cites the reference interface it stands in for
 * (paths relative to the gadget3 source tree).
 *
 * Conventions: plain pointers and sizes, no torch/R types. All functions
 * return 0 on success or a negative G3CUDA_E* code; the message is available
 * from g3cuda_last_error(). A NaN objective is a VALUE, not an error (it
 * mirrors assert_msg()/return NaN in R/aab_env.R:98-107, R/action_time.R:78).
 * There is no CPU fallback: without a CUDA device every compute entry point
 * fails with G3CUDA_ENODEVICE.
 */
#ifndef G3CUDA_H
#define G3CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define G3CUDA_OK          0
#define G3CUDA_EINVAL     -1   /* malformed spec / bad argument */
#define G3CUDA_ENODEVICE  -2   /* no usable CUDA device */
#define G3CUDA_ECUDA      -3   /* CUDA runtime error */
#define G3CUDA_EUNSUPP    -4   /* model outside the supported action subset */
#define G3CUDA_ENOMEM     -5

/* Packed model specification; layout in g3cuda_spec.h.
 * Replaces: the generated TMB translation unit + model_data environment
 * returned by g3_to_tmb() (R/run_tmb.R:1076-1091). */
typedef struct g3cuda_spec {
    int32_t        version;   /* G3CUDA_SPEC_VERSION */
    int64_t        n_ints;
    const int32_t* ints;
    int64_t        n_reals;
    const double*  reals;
} g3cuda_spec;

typedef struct g3cuda_model g3cuda_model;

/* Build a model handle: validates the spec, copies it to the device
 * (device = CUDA ordinal), precomputes parameter-independent tables.
 * The caller keeps ownership of `spec`.
 * Replaces: g3_tmb_adfun() -> TMB::compile + dyn.load + TMB::MakeADFun
 * (R/run_tmb.R:1118-1260). */
int g3cuda_model_create(const g3cuda_spec* spec, int device, g3cuda_model** out);

/* Free a handle. Replaces: the ADFun external pointer finalizer. */
void g3cuda_model_free(g3cuda_model* m);

/* Model dimensions, for callers sizing buffers. */
int32_t g3cuda_n_par(const g3cuda_model* m);    /* full PARAMETER vector length */
int32_t g3cuda_n_nll(const g3cuda_model* m);    /* rows of nll_comp: one per g3l_distribution
                                                   component (run order), then understocking, then bounds */
int32_t g3cuda_n_steps(const g3cuda_model* m);  /* allocated model steps (total_steps + 1 at retro_years = 0) */
int64_t g3cuda_report_len(const g3cuda_model* m); /* doubles written by g3cuda_report() */

/* Evaluate the objective for a batch of B independent parameter vectors.
 *   par         B x n_par, row-major, HOST memory; full PARAMETER order
 *               (fixed parameters included), i.e. obj$env$last.par order.
 *   tangent_idx K parameter indices to differentiate with respect to, or NULL
 *   nll         out, B
 *   nll_comp    out, B x n_nll (unweighted per-component sums over time;
 *               understocking = sum of nll_understocking__wgt; bounds = penalty sum), or NULL
 *   grad        out, B x K forward-mode d nll / d par[tangent_idx[k]], or NULL
 * A vector with retro_years < 0 yields NaN (a value, as in the reference: R/action_time.R:78-79); a vector with
 * project_years beyond the years the spec's per-step tables cover (NT / NSTEPS - (END_YEAR - START_YEAR + 1): the
 * max_project_years the model was lowered with, 0 by default) is refused with G3CUDA_EUNSUPP; the device-pointer entry
 * point cannot look at the values and returns NaN for such rows.
 * Replaces: obj$fn(par), obj$gr(par) of the TMB ADFun (R/run_tmb.R:1254-1260,
 * man/run_tmb.Rd:233-242), called once per vector by nlminb/optim. */
int g3cuda_eval_batch(g3cuda_model* m, const double* par, int64_t B,
                      const int32_t* tangent_idx, int32_t K,
                      double* nll, double* nll_comp, double* grad);

/* Same computation with DEVICE pointers, asynchronous on `stream`
 * (a cudaStream_t passed as void*; NULL = default stream). Used by callers
 * that keep parameter batches resident in HBM (bench.py `value`).
 * A handle owns ONE evaluation workspace: launches on different streams are
 * ordered by the library (each waits for the previous one's completion event);
 * calls from different host threads on one handle must be serialised by the caller. */
int g3cuda_eval_batch_device(g3cuda_model* m, const double* d_par, int64_t B,
                             const int32_t* d_tangent_idx, int32_t K,
                             double* d_nll, double* d_nll_comp, double* d_grad,
                             void* stream);

/* One handle over several devices of the box: the spec and observation data are copied to every device in `devices`
 * (CUDA ordinals; one may appear twice), and g3cuda_multi_eval_batch() cuts the B rows into contiguous blocks, one per
 * device, evaluated concurrently (one host thread and one stream set per device); results land in the caller's buffers.
 * The path has no exchange step, so no device talks to another. Same arguments, same error conventions as the
 * single-device calls. This is how ONE R session drives all 8 GPUs: obj <- g3_cuda_adfun(model, devices = 0:7);
 * obj$fn_batch(P). */
typedef struct g3cuda_multi g3cuda_multi;
int g3cuda_multi_create(const g3cuda_spec* spec, const int32_t* devices, int32_t n_devices, g3cuda_multi** out);
void g3cuda_multi_free(g3cuda_multi* mm);
int32_t g3cuda_multi_n_devices(const g3cuda_multi* mm);
g3cuda_model* g3cuda_multi_model(g3cuda_multi* mm, int32_t i);     /* the i-th device's handle (dimensions, report, tuning hooks) */
int g3cuda_multi_eval_batch(g3cuda_multi* mm, const double* par, int64_t B,
                            const int32_t* tangent_idx, int32_t K,
                            double* nll, double* nll_comp, double* grad);

/* Kernel selection, for tests and profiling: 0 = automatic (default), 3 = thread-per-evaluation kernel,
 * 4 = tile kernel, 5 = tile kernel in its full instantiation (by default a model runs on the instantiation that has the
 * features it does not use -- migration, stock predators, constant maturity, the rarer likelihood functions -- compiled out;
 * the same source for what remains; the two compilations agree to rounding) (G3CUDA_EUNSUPP if the model does not fit the requested one;
 * other values: G3CUDA_EINVAL).
 * Automatic: value batches of any size run on the tile kernel (one CTA owns 32 evaluations for the whole time
 * loop, stock state in shared memory; few vectors are spread over the SMs, so a single fn(par) is a ~ms call);
 * gradient batches run on the tile kernel when small and on the thread kernel when large; g3cuda_report always
 * runs on the thread kernel. The kernels agree to rounding, not to the bit.
 * All are CUDA; there is no CPU path to select. */
int g3cuda_set_kernel(g3cuda_model* m, int which);

/* Tuning hook: warps per tile of the tile kernel (the units -- row segments of the stocks' columns -- are dealt to the
 * warps); 0 restores the planner's choice. The partition into units and the order of every sum stay the same, so results
 * do not depend on the warp count at all. */
int g3cuda_set_tile_warps(g3cuda_model* m, int warps);

/* Tuning hook: at most `tiles` tiles (CTAs) of the tile kernel in flight; 0 = every slot of the device. The per-tile slabs
 * of the tiles in flight are the kernel's L2 working set. Results do not depend on it. */
int g3cuda_set_tile_slots(g3cuda_model* m, int tiles);

/* Tuning hook: re-plan the tile kernel with `warps` unit warps (0 = the planner's choice) and row segments of `seg_rows`
 * rows per unit (0 = the planner's choice, -1 = whole columns). The partition only regroups partial sums (per-unit
 * suitable biomass and understocking sums): results agree to rounding between partitions. */
int g3cuda_set_tile_partition(g3cuda_model* m, int warps, int seg_rows);

/* Introspection for tests and bench.py: out[0] = the model fits the tile kernel, out[1] = warps per tile,
 * out[2] = per-evaluation table entries in the CTA's global slab, out[3] = per-evaluation state entries,
 * out[4] = the state of a value run fits shared memory, out[5] = the model fits the thread kernel. */
int32_t g3cuda_tile_info(const g3cuda_model* m, int32_t* out);

/* Number of kernels launched by this handle so far (bench.py gpu_launches). */
int64_t g3cuda_launch_count(const g3cuda_model* m);

/* Milliseconds spent in the last g3cuda_eval_batch[_device] kernel(s),
 * measured with CUDA events on the launching stream (synchronises). */
float g3cuda_last_kernel_ms(g3cuda_model* m);

/* Single-vector report run: fills `out` (g3cuda_report_len doubles, HOST) with
 *   [0]                       nll
 *   [1 .. 1+n_nll*NT)         per-step unweighted nll values, row-major [n_nll][NT]
 *                             (nll_<comp>__num/wgt, nll_understocking__wgt; bounds row: step 0 only)
 *   then per stock, in stock order: dstart num [NT][cells of stock], dstart wgt [NT][cells]
 *   (state at the start of each step, canonical cell layout, as dstart_<stock>__num/wgt
 *   of g3a_report_detail, R/action_report.R:177-213)
 *   then per component: model array [n_times][n_dest] (canonical dest layout)
 *   then per component: 2 doubles (surveyindices alpha, beta; NaN otherwise)
 *   then the rest of g3a_report_detail (R/action_report.R:177-213):
 *     per stock: renewalnum [NT][cells]                       (detail_<stock>__renewalnum)
 *     per predator-prey pair, in packed pair order:
 *       suitability [predator length groups (1 for a fleet)][prey L]   (suit_<prey>_<pred>__report)
 *       suit [NT][prey cells]   (detail_<prey>_<pred>__suit; fleets only, 0 for stock predators)
 *       cons [NT][prey cells]   (detail_<prey>_<pred>__cons == detail_<prey>__predby_<pred>,
 *                                summed over the predator's own dimensions)
 * Replaces: obj$report(par) (R/run_tmb.R:1277-1321). */
int g3cuda_report(g3cuda_model* m, const double* par, double* out);

/* Measured fp64 FMA throughput of `device` in TFLOP/s (DFMA-chain microbenchmark; the roofline
 * denominator bench.py reports against). Negative on failure. */
double g3cuda_fp64_peak_tflops(int device);

/* Device-math probe for the parity tests: evaluates, on `device`, the kernels' own implementation of
 *   which = 0: logspace_add(x[i], y[i])          (R/aab_env.R:16-19)
 *   which = 1: avoid_zero(x[i])                  (R/aab_env.R:24-31; y ignored)
 *   which = 2: dif_pmin(x[i], y[i], 1e3)         (R/env_dif.R:37-47)
 *   which = 3: avoid_zero(x[i]) in the exact-division form the predation step uses
 *   which = 4: dif_pmin(x[i], y[i], 1e3) in the exact-division form
 * for n HOST values into out (HOST). Returns 0 or a negative G3CUDA_E* code. */
int g3cuda_math_probe(int device, int which, const double* x, const double* y, int64_t n, double* out);

/* Last error message of the calling thread. */
const char* g3cuda_last_error(void);

/* Library/ABI version: G3CUDA_SPEC_VERSION the library was built for. */
int32_t g3cuda_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* G3CUDA_H */
