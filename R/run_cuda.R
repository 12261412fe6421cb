# run_cuda.R -- the CUDA backend beside g3_to_tmb() / g3_tmb_adfun() (/root/reference R/run_tmb.R:709, :1118).
#
# What a gadget3 maintainer adds to the package; it needs src/g3cuda_R.c built against libg3cuda.so.
# R is not in this repository's build image, so these functions are not executed by its tests; the Python package
# gadget3_b200/ mirrors them one to one (same names, same argument meaning, same errors) and IS executed.
#
#   g3_to_cuda(actions)              lower an action list into the packed model spec (ints, reals) -- see g3_cuda_lower()
#   g3_cuda_adfun(model, parameters) bind it to the GPU(s): $par $fn $gr $report like g3_tmb_adfun(), plus
#                                    $fn_batch(P) $gr_batch(P) $nll_components_batch(P) for a matrix of vectors
#   ut_cuda_tmb_compare(...)         differential test against the TMB backend, patterned on ut_tmb_r_compare2
#                                    (R/test_utils.R:87-105)

# The lowering itself (walk the g3a_*/g3l_* closures, resolve g3_param_table lookups, landings and observation
# tables, emit slot programs and the record tables of include/g3cuda_spec.h) is ~1,500 lines in the Python mirror
# (gadget3_b200/run_cuda.py, formula.py, actions.py, likelihood.py). Until it is ported, an R session takes the spec
# from files the mirror wrote: gadget3_b200.run_cuda.save_spec(model, path, report_len = ...) -> <path>.ints / .reals / .json / .idx
g3_cuda_load_spec <- function(path) {
    meta <- jsonlite::fromJSON(paste0(path, ".json"))
    ints <- readBin(paste0(path, ".ints"), "integer", n = meta$n_ints, size = 4L)
    reals <- readBin(paste0(path, ".reals"), "double", n = meta$n_reals, size = 8L)
    pt <- data.frame(
        switch = meta$parameter_template$switch,
        value = I(as.list(meta$parameter_template$value)),
        optimise = meta$parameter_template$optimise,
        lower = meta$parameter_template$lower,
        upper = meta$parameter_template$upper,
        stringsAsFactors = FALSE)
    rownames(pt) <- pt$switch
    idx <- if (file.exists(paste0(path, ".idx"))) readBin(paste0(path, ".idx"), "integer", n = file.size(paste0(path, ".idx")) / 4L, size = 4L) else integer(0)
    structure(list(ints = ints, reals = reals, report_layout = meta$report_layout, report_idx = idx), class = "g3_cuda",
        parameter_template = pt)
}

g3_to_cuda <- function(actions, strict = FALSE, spec_file = NULL, max_project_years = 0L) {
    if (!is.null(spec_file)) {
        model <- g3_cuda_load_spec(spec_file)
        attr(model, "actions") <- actions
        return(model)
    }
    collated <- g3_collate(actions)                       # same ordering contract as g3_to_tmb (R/run.R:31-72)
    spec <- g3_cuda_lower(collated, max_project_years)    # per-step tables reach end_year + max_project_years (project_years runs);
                                                          # stop("g3_to_cuda: unsupported action ", name) outside the subset:
                                                          # there is no CPU fallback
    structure(list(ints = spec$ints, reals = spec$reals, report_layout = spec$report_layout), class = "g3_cuda",
        actions = actions,
        parameter_template = spec$parameter_template,     # identical rows / order to g3_to_tmb()'s
        model_data = spec$model_data)
}

g3_cuda_adfun <- function(model, parameters = attr(model, "parameter_template"), devices = 0L) {
    template <- attr(model, "parameter_template")
    if (!identical(parameters$switch, template$switch)) stop("Parameters not in expected order")    # R/run_tmb.R:1156-1158
    if (!any(parameters$optimise)) stop("No parameters with optimise=TRUE set")                      # R/run_tmb.R:1166-1168
    spec <- g3_cuda_pack_bounds(model, parameters)        # fills the bounds section like R/run_tmb.R:1234-1235
    h <- .Call("g3cuda_R_create", spec$ints, spec$reals, as.integer(devices))
    dims <- .Call("g3cuda_R_dims", h)
    full <- unlist(parameters$value)
    opt <- which(parameters$optimise)
    expand <- function(P) {                               # optimiser-facing n_opt x B -> full n_par x B
        P <- as.matrix(P)
        if (nrow(P) != length(opt)) stop("expected ", length(opt), " parameters per vector, got ", nrow(P))
        out <- matrix(full, length(full), ncol(P))
        out[opt, ] <- P
        out
    }
    env <- new.env()
    env$last.par <- full
    list(
        par = structure(full[opt], names = cpp_escape_varname(parameters$switch[opt])),
        fn = function(par = full[opt]) { env$last.par[opt] <- par; .Call("g3cuda_R_eval", h, expand(par), NULL, FALSE)[[1]] },
        gr = function(par = full[opt]) { env$last.par[opt] <- par; t(.Call("g3cuda_R_eval", h, expand(par), opt - 1L, FALSE)[[3]]) },
        report = function(par = full[opt]) g3_cuda_unpack_report(model, .Call("g3cuda_R_report", h, expand(par)[, 1])),
        fn_batch = function(P) .Call("g3cuda_R_eval", h, expand(P), NULL, FALSE)[[1]],                 # P: n_opt x B
        gr_batch = function(P) t(.Call("g3cuda_R_eval", h, expand(P), opt - 1L, FALSE)[[3]]),           # B x n_opt
        nll_components_batch = function(P) {
            out <- .Call("g3cuda_R_eval", h, expand(P), NULL, TRUE)
            list(nll = out[[1]], components = t(out[[2]]))
        },
        env = env, handle = h, n_devices = dims[[5]])
}

# bounds penalty section: every parameter with finite lower and upper, in C-locale name order (R/likelihood_bounds.R)
g3_cuda_pack_bounds <- function(model, parameters) {
    model    # (the Python mirror's G3CudaModel.pack(); identical arithmetic on ints/reals)
}

# flat report vector -> named arrays in the model's own dimension order (the Python mirror's unpack_report())
g3_cuda_unpack_report <- function(model, raw) {
    out <- list()
    lay <- model$report_layout
    for (i in seq_len(nrow(lay))) {
        dims <- lay$dim[[i]]
        at <- model$report_idx[seq.int(lay$idx_offset[[i]] + 1L, length.out = prod(dims))]
        arr <- ifelse(at < 0L, NaN, raw[pmax(at, 0L) + 1L])      # position of every element in g3cuda_report()'s vector
        if (length(dims) > 1L) arr <- array(arr, dim = dims)
        out[[lay$name[[i]]]] <- arr
    }
    out
}

# Differential test against the TMB backend on the same vectors: 1e-10 on nll and reports, 1e-8 on gradients
ut_cuda_tmb_compare <- function(model_cpp, model_cuda, parameters, n = 16, seed = 1) {
    obj_tmb <- g3_tmb_adfun(model_cpp, parameters, compile_flags = c("-O2"))
    obj_cuda <- g3_cuda_adfun(model_cuda, parameters)
    set.seed(seed)
    lo <- g3_tmb_lower(parameters); up <- g3_tmb_upper(parameters)
    for (i in seq_len(n)) {
        par <- obj_tmb$par + runif(length(lo), -0.1, 0.1) * ifelse(is.finite(up - lo), up - lo, abs(obj_tmb$par))
        par <- pmin(pmax(par, lo), up)
        nll_t <- obj_tmb$fn(par); nll_c <- obj_cuda$fn(par)
        ok(abs(nll_c - nll_t) <= 1e-10 * abs(nll_t), paste0("nll agrees with TMB, vector ", i))
        g_t <- obj_tmb$gr(par); g_c <- obj_cuda$gr(par)
        ok(max(abs(g_c - g_t)) <= 1e-8 * max(abs(g_t)), paste0("gradient agrees with TMB, vector ", i))
        rep_t <- obj_tmb$report(par); rep_c <- obj_cuda$report(par)
        for (nm in intersect(names(rep_t), names(rep_c)))
            ok(isTRUE(all.equal(as.vector(rep_c[[nm]]), as.vector(rep_t[[nm]]), tolerance = 1e-10)), paste0(nm, " agrees with TMB"))
    }
}
