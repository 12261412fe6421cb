/* g3cuda_R.c -- the .Call shim between gadget3 (R) and libg3cuda.so (include/g3cuda.h).
 *
 * Replaces, on the native side, what g3_tmb_adfun() reaches through TMB's MakeADFun external pointer
 * (/root/reference R/run_tmb.R:1229-1260): create a model handle from the packed spec, evaluate batches of parameter
 * vectors, produce the report vector. It only marshals: SEXP -> pointers and sizes, errors -> Rf_error(). No R
 * header is needed by the library itself.
 *
 * Build (on a machine with R):  R CMD SHLIB src/g3cuda_R.c -Iinclude -Lgadget3_b200/csrc -lg3cuda
 * Syntax-checked in this repository against tests/r_stub/ (R is not in the build image): tests/test_r_glue.py.
 */
#include <R.h>
#include <Rinternals.h>

#include "g3cuda.h"
#include "g3cuda_spec.h"

static void g3cuda_R_finalize(SEXP ptr) {
    g3cuda_multi* m = (g3cuda_multi*) R_ExternalPtrAddr(ptr);
    if (m) { g3cuda_multi_free(m); R_ClearExternalPtr(ptr); }
}

static g3cuda_multi* handle_of(SEXP handle) {
    g3cuda_multi* m = (g3cuda_multi*) R_ExternalPtrAddr(handle);
    if (!m) Rf_error("g3cuda: the model handle was freed (or restored from a saved workspace): call g3_cuda_adfun() again");
    return m;
}

/* .Call("g3cuda_R_create", ints, reals, devices) -> external pointer.
 * devices: integer vector of CUDA ordinals; the model is copied to each of them (g3cuda_multi_create). */
SEXP g3cuda_R_create(SEXP ints, SEXP reals, SEXP devices) {
    g3cuda_spec spec;
    g3cuda_multi* m = NULL;
    SEXP ptr;
    spec.version = G3CUDA_SPEC_VERSION;
    spec.n_ints = (int64_t) XLENGTH(ints);   spec.ints = INTEGER(ints);
    spec.n_reals = (int64_t) XLENGTH(reals); spec.reals = REAL(reals);
    if (g3cuda_multi_create(&spec, INTEGER(devices), (int32_t) XLENGTH(devices), &m) != G3CUDA_OK)
        Rf_error("g3_cuda_adfun: %s", g3cuda_last_error());          /* no CPU fallback */
    ptr = PROTECT(R_MakeExternalPtr(m, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, g3cuda_R_finalize, TRUE);
    UNPROTECT(1);
    return ptr;
}

/* .Call("g3cuda_R_dims", handle) -> c(n_par, n_nll, n_steps, report_len, n_devices) */
SEXP g3cuda_R_dims(SEXP handle) {
    g3cuda_multi* mm = handle_of(handle);
    g3cuda_model* m = g3cuda_multi_model(mm, 0);
    SEXP out = PROTECT(Rf_allocVector(REALSXP, 5));
    REAL(out)[0] = g3cuda_n_par(m); REAL(out)[1] = g3cuda_n_nll(m); REAL(out)[2] = g3cuda_n_steps(m);
    REAL(out)[3] = (double) g3cuda_report_len(m); REAL(out)[4] = g3cuda_multi_n_devices(mm);
    UNPROTECT(1);
    return out;
}

/* .Call("g3cuda_R_eval", handle, par, tangent_idx or NULL, want_comp) -> list(nll, nll_comp or NULL, grad or NULL)
 * par: numeric matrix n_par x B, ONE VECTOR PER COLUMN, full PARAMETER order (obj$env$last.par order). R matrices are
 * column-major, so this is exactly the row-major B x n_par block the ABI takes. tangent_idx: 0-based parameter indices.
 * nll_comp comes back n_nll x B, grad K x B (the caller transposes). */
SEXP g3cuda_R_eval(SEXP handle, SEXP par, SEXP tangent_idx, SEXP want_comp) {
    g3cuda_multi* mm = handle_of(handle);
    g3cuda_model* m = g3cuda_multi_model(mm, 0);
    R_xlen_t P = g3cuda_n_par(m), B;
    int K = Rf_isNull(tangent_idx) ? 0 : (int) XLENGTH(tangent_idx), C = g3cuda_n_nll(m), rc;
    SEXP out, nll, comp = R_NilValue, grad = R_NilValue;
    if (!Rf_isReal(par) || XLENGTH(par) % P != 0) Rf_error("g3cuda: par must be a numeric matrix with %d rows", (int) P);
    B = XLENGTH(par) / P;
    out = PROTECT(Rf_allocVector(VECSXP, 3));
    nll = SET_VECTOR_ELT(out, 0, Rf_allocVector(REALSXP, B));
    if (Rf_asLogical(want_comp)) comp = SET_VECTOR_ELT(out, 1, Rf_allocMatrix(REALSXP, C, (int) B));
    if (K > 0) grad = SET_VECTOR_ELT(out, 2, Rf_allocMatrix(REALSXP, K, (int) B));
    rc = g3cuda_multi_eval_batch(mm, REAL(par), (int64_t) B, K ? INTEGER(tangent_idx) : NULL, K, REAL(nll),
                                 comp == R_NilValue ? NULL : REAL(comp), grad == R_NilValue ? NULL : REAL(grad));
    if (rc != G3CUDA_OK) Rf_error("g3cuda_eval_batch: %s", g3cuda_last_error());
    UNPROTECT(1);
    return out;        /* a NaN nll is a value (mirrors assert_msg(): warning + NaN, R/aab_env.R:98-107), not an error */
}

/* .Call("g3cuda_R_report", handle, par) -> numeric vector; layout documented in g3cuda.h, unpacked by g3_cuda_unpack_report() */
SEXP g3cuda_R_report(SEXP handle, SEXP par) {
    g3cuda_model* m = g3cuda_multi_model(handle_of(handle), 0);
    SEXP out;
    if (!Rf_isReal(par) || XLENGTH(par) != g3cuda_n_par(m)) Rf_error("g3cuda: par must hold %d values", (int) g3cuda_n_par(m));
    out = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t) g3cuda_report_len(m)));
    if (g3cuda_report(m, REAL(par), REAL(out)) != G3CUDA_OK) Rf_error("g3cuda_report: %s", g3cuda_last_error());
    UNPROTECT(1);
    return out;
}

static const R_CallMethodDef g3cuda_R_calls[] = {
    {"g3cuda_R_create", (DL_FUNC) &g3cuda_R_create, 3},
    {"g3cuda_R_dims", (DL_FUNC) &g3cuda_R_dims, 1},
    {"g3cuda_R_eval", (DL_FUNC) &g3cuda_R_eval, 4},
    {"g3cuda_R_report", (DL_FUNC) &g3cuda_R_report, 2},
    {NULL, NULL, 0}
};

void R_init_g3cuda_R(DllInfo* dll) {
    R_registerRoutines(dll, NULL, g3cuda_R_calls, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
