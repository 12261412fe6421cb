#!/usr/bin/env Rscript
# Dump a TMB "truth bundle" for the CUDA backend to be checked against (SURVEY section 8c).
#
# Needs a machine with R, TMB and gadget3 installed; it is NOT run by this repository's tests
# (there is no R in the build image). It builds the two vignette models the same way
# gadget3_b200/configs.py::build_c1 / build_c2 do, evaluates obj$fn / obj$gr / obj$report() on
# N parameter vectors, and writes everything tools/replay_tmb_bundle.py needs to rebuild the very
# same model (data tables + parameter template) and to compare at 1e-10 (nll) / 1e-8 (gradient):
#
#   Rscript tools/compare_with_tmb.R single out_dir [N]
#   Rscript tools/compare_with_tmb.R multi  out_dir [N]
#   python  tools/replay_tmb_bundle.py out_dir            # on the GPU box
#
# Bundle layout (all CSV, full double precision):
#   model.txt                       "single" | "multi"
#   landings.csv ldist.csv aldist.csv [matp.csv] si.csv     the observation tables handed to gadget3
#   template.csv                    switch,value,optimise,lower,upper  (g3_to_tmb()'s parameter_template order)
#   par.csv                         N x n_par, every PARAMETER in template order (fixed ones included)
#   fn.csv                          N nll values (obj$fn on the optimised subset)
#   gr.csv                          N x n_optimised gradients (obj$gr)
#   report_<i>.csv                  name,index,value for the nll_* arrays of vector i (first 2 vectors)
suppressPackageStartupMessages({ library(gadget3) })

args <- commandArgs(trailingOnly = TRUE)
which_model <- if (length(args) >= 1) args[[1]] else "single"
out_dir <- if (length(args) >= 2) args[[2]] else paste0("tmb_bundle_", which_model)
n_vec <- if (length(args) >= 3) as.integer(args[[3]]) else 32L
dir.create(out_dir, showWarnings = FALSE, recursive = TRUE)
set.seed(123)
multi <- identical(which_model, "multi")
years <- 1990:2023
area_names <- g3_areas(c("IXa", "IXb"))

count_by <- function(df, cols) {
    df$number <- 1
    out <- aggregate(df["number"], df[cols], sum)
    out[do.call(order, out[cols]), ]
}
len_breaks <- if (multi) c(seq(0, 110, 5), Inf) else c(seq(0, 80, 20), Inf)

# ---- observation tables (distributions as in the vignettes; the draws themselves are stored in the bundle)
landings <- data.frame(year = years, step = 2, area = "IXa",
    weight = if (multi) rnorm(length(years), 1e5, 10) else rnorm(length(years), 1000, 100))
n_l <- if (multi) 1000 else 100
raw <- data.frame(year = rep(years, each = n_l), step = 2)
raw$length <- cut(rnorm(nrow(raw), 50, 20), breaks = len_breaks, right = FALSE)
ldist <- count_by(raw[!is.na(raw$length), ], c("year", "step", "length"))
raw <- data.frame(year = rep(years, each = 10000), step = 2)
raw$age <- if (multi) floor(pmin(rnorm(nrow(raw), 6, 1), 10)) else floor(runif(nrow(raw), 1, 5))
raw$len_cm <- rnorm(nrow(raw), 30 + raw$age * (if (multi) 2 else 10), 20)
raw$length <- cut(raw$len_cm, breaks = len_breaks, right = FALSE)
raw <- raw[!is.na(raw$length), ]
aldist <- count_by(raw, c("year", "step", "age", "length"))
if (multi) {
    m <- aldist
    m$stock <- ifelse(m$age > runif(nrow(m), 3, 5), "fish_mat", "fish_imm")
    matp <- aggregate(m["number"], m[c("year", "step", "length", "stock")], sum)
}
si <- data.frame(year = years, step = 3, area = "IXa", weight = runif(length(years), 1e4, 1e5))

# ---- model
f_surv <- g3_fleet("f_surv") |> g3s_livesonareas(area_names["IXa"])
if (!multi) {
    fish <- g3_stock("fish", seq(10, 100, 10)) |> g3s_livesonareas(area_names["IXa"]) |> g3s_age(1, 5)
    stocks <- list(fish)
    stock_actions <- list(
        g3a_growmature(fish, g3a_grow_impl_bbinom(maxlengthgroupgrowth = 4L)),
        g3a_naturalmortality(fish),
        g3a_initialconditions_normalcv(fish),
        g3a_renewal_normalcv(fish, run_step = 2),
        g3a_age(fish))
    suit <- g3_suitability_exponentiall50()
} else {
    st_imm <- g3_stock(c(species = "fish", maturity = "imm"), seq(5, 100, 5)) |> g3s_age(1, 5) |> g3s_livesonareas(area_names["IXa"])
    st_mat <- g3_stock(c(species = "fish", maturity = "mat"), seq(5, 100, 5)) |> g3s_age(3, 10) |> g3s_livesonareas(area_names["IXa"])
    stocks <- list(st_imm, st_mat)
    stock_actions <- list(
        g3a_growmature(st_imm, g3a_grow_impl_bbinom(maxlengthgroupgrowth = 4L),
            maturity_f = g3a_mature_continuous(), output_stocks = list(st_mat), transition_f = ~TRUE),
        g3a_naturalmortality(st_imm),
        g3a_initialconditions_normalcv(st_imm),
        g3a_renewal_normalcv(st_imm),
        g3a_age(st_imm, output_stocks = list(st_mat)),
        g3a_growmature(st_mat, g3a_grow_impl_bbinom(maxlengthgroupgrowth = 4L)),
        g3a_naturalmortality(st_mat),
        g3a_initialconditions_normalcv(st_mat),
        g3a_age(st_mat))
    suit <- g3_suitability_exponentiall50(by_stock = "species")
}
dist <- function(name, data, f) g3l_catchdistribution(name, data, fleets = list(f_surv), stocks = stocks,
    function_f = f, area_group = area_names, report = TRUE, nll_breakdown = TRUE)
actions <- c(list(
    g3a_time(1990L, 2023L, c(3L, 3L, 3L, 3L))),
    stock_actions, list(
    g3l_understocking(stocks, nll_breakdown = TRUE),
    g3a_predate_fleet(f_surv, stocks, suitabilities = suit,
        catchability_f = g3a_predate_catchability_totalfleet(
            g3_timeareadata("landings_f_surv", landings, "weight", areas = area_names))),
    dist("ldist_f_surv", ldist, g3l_distribution_sumofsquares()),
    dist("aldist_f_surv", aldist, g3l_distribution_sumofsquares()),
    if (multi) dist("matp_f_surv", matp, g3l_distribution_sumofsquares()),
    g3l_abundancedistribution("dist_si_cpue", si, stocks = stocks,
        function_f = g3l_distribution_surveyindices_log(alpha = NULL, beta = 1),
        area_group = area_names, report = TRUE, nll_breakdown = TRUE)))
actions <- c(actions, list(g3a_report_detail(actions), g3l_bounds_penalty(actions)))
model_cpp <- g3_to_tmb(actions)

# ---- parameter template: the same g3_init_val() calls as configs.build_c1 / build_c2
midlen <- g3_stock_def(stocks[[1]], "midlen")
est_l50 <- midlen[[length(midlen) / 2]]
est_linf <- if (multi) midlen[[length(midlen) - 2]] else max(midlen)
est_t0 <- g3_stock_def(stocks[[1]], "minage") - 0.8
p <- attr(model_cpp, "parameter_template")
if (!multi) {
    p <- p |>
        g3_init_val("*.rec|init.scalar", 1, optimise = FALSE) |>
        g3_init_val("*.init.#", 10, lower = 0.001, upper = 30) |>
        g3_init_val("*.rec.#", 1e-4, lower = 1e-6, upper = 1e-2) |>
        g3_init_val("*.rec.proj", 0.002) |>
        g3_init_val("init.F", 0.5, lower = 0.1, upper = 1) |>
        g3_init_val("*.M.#", 0.15, lower = 0.001, upper = 10) |>
        g3_init_val("*.Linf", est_linf, spread = 0.2) |>
        g3_init_val("*.K", 0.3, lower = 0.04, upper = 1.2) |>
        g3_init_val("*.t0", est_t0, spread = 2) |>
        g3_init_val("*.walpha", 0.01, optimise = FALSE) |>
        g3_init_val("*.wbeta", 3, optimise = FALSE) |>
        g3_init_val("*.*.alpha", 0.07, lower = 0.01, upper = 0.2) |>
        g3_init_val("*.*.l50", est_l50, spread = 0.25) |>
        g3_init_val("*.bbin", 1, lower = 1e-05, upper = 10)
} else {
    p <- p |>
        g3_init_val("*.rec|init.scalar", 1000, optimise = FALSE) |>
        g3_init_val("*.init.#", 10, lower = 0.001, upper = 10) |>
        g3_init_val("*.rec.#", 100, lower = 1e-6, upper = 1000) |>
        g3_init_val("*.rec.proj", 0.002) |>
        g3_init_val("*.M.#", 0.15, lower = 0.001, upper = 10) |>
        g3_init_val("init.F", 0.5, lower = 0.1, upper = 10) |>
        g3_init_val("*.Linf", est_linf, spread = 2) |>
        g3_init_val("*.K", 0.3, lower = 0.04, upper = 1.2) |>
        g3_init_val("*.t0", est_t0, optimise = FALSE) |>
        g3_init_val("*.walpha", 0.01, optimise = FALSE) |>
        g3_init_val("*.wbeta", 3, optimise = FALSE) |>
        g3_init_val("*.*.alpha", 0.07, lower = 0.01, upper = 1.8) |>
        g3_init_val("*.*.l50", est_l50, spread = 1.5) |>
        g3_init_val("*.mat.alpha", 0.07, lower = 0.0001, upper = 1.8) |>
        g3_init_val("*.mat.l50", est_l50, spread = 3.5) |>
        g3_init_val("*.bbin", 100, lower = 1e-05, upper = 1500)
}
p[["report_detail", "value"]] <- 0L
obj <- g3_tmb_adfun(model_cpp, p)

# ---- N vectors: template +- 10 % of the bounds range on the optimised rows (configs.parameter_batch)
value <- unlist(p$value); lower <- p$lower; upper <- p$upper; opt <- which(p$optimise)
par <- matrix(value, nrow = n_vec, ncol = length(value), byrow = TRUE)
for (i in seq_len(n_vec)[-1]) {
    width <- 0.10 * (upper[opt] - lower[opt])
    par[i, opt] <- pmin(pmax(value[opt] + runif(length(opt), -1, 1) * width, lower[opt]), upper[opt])
}
fn <- numeric(n_vec); gr <- matrix(0, n_vec, length(opt))
for (i in seq_len(n_vec)) {
    fn[i] <- obj$fn(par[i, opt])
    gr[i, ] <- as.vector(obj$gr(par[i, opt]))
    if (i <= 2) {
        r <- obj$report(par[i, opt])
        keep <- grep("^nll", names(r), value = TRUE)
        rows <- do.call(rbind, lapply(keep, function(n) data.frame(
            name = n, index = seq_along(r[[n]]) - 1L, value = sprintf("%.17g", as.vector(r[[n]])))))
        write.csv(rows, file.path(out_dir, sprintf("report_%d.csv", i - 1L)), row.names = FALSE, quote = FALSE)
    }
}

# ---- write the bundle
w <- function(df, name) write.csv(df, file.path(out_dir, name), row.names = FALSE, quote = TRUE)
num <- function(m) { m <- as.data.frame(m); m[] <- lapply(m, function(x) sprintf("%.17g", x)); m }
writeLines(which_model, file.path(out_dir, "model.txt"))
landings$weight <- sprintf("%.17g", landings$weight); si$weight <- sprintf("%.17g", si$weight)
w(landings, "landings.csv"); w(ldist, "ldist.csv"); w(aldist, "aldist.csv"); w(si, "si.csv")
if (multi) w(matp, "matp.csv")
w(data.frame(switch = p$switch, value = sprintf("%.17g", value), optimise = as.integer(p$optimise),
    lower = sprintf("%.17g", lower), upper = sprintf("%.17g", upper)), "template.csv")
write.table(num(par), file.path(out_dir, "par.csv"), sep = ",", row.names = FALSE, col.names = FALSE, quote = FALSE)
write.table(num(matrix(fn)), file.path(out_dir, "fn.csv"), sep = ",", row.names = FALSE, col.names = FALSE, quote = FALSE)
write.table(num(gr), file.path(out_dir, "gr.csv"), sep = ",", row.names = FALSE, col.names = FALSE, quote = FALSE)
cat("bundle written to", out_dir, ":", n_vec, "vectors,", length(value), "parameters,", length(opt), "optimised\n")
