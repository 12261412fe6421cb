/*
 * g3oracle.cpp -- CPU restatement of gadget3's generated objective function.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / reported baseline.
 *
 * What it restates: the body of `objective_function<Type>::operator()` that
 * g3_to_tmb() generates (reference: baseline/multiple-substocks.cpp:692-2161,
 * baseline/introduction-single-stock.cpp:542-1389, inttest/codegeneration/ling.cpp),
 * statement by statement and in the same order, driven by the packed spec of
 * include/g3cuda_spec.h instead of being hard-coded per model. It deliberately keeps
 * the reference's formulation (dense L x L growth matrices, nine-lgamma beta-binomial,
 * every avoid_zero evaluated in full with libm exp/log1p, per-age recomputation) so
 * that it is an independent check of the restructured CUDA path.
 *
 * Parity pin: (1) the known-answer vectors of the reference's unit tests
 * (tests/test_oracle_kat.py <- tests/test-action_grow.R etc.); (2) the reference's own
 * generated C++ (baseline/*.cpp) compiled unmodified against oracle/tmb_shim into
 * oracle/_ref -- whole-model nll and reports agree to ~1e-16..2e-12
 * (tests/test_reference_golden.py, tests/golden/reference_c{1,2}.npz).
 *
 * Scalar type T is double or Dual<K> (forward-mode tangents) so the same
 * restatement is the gradient oracle.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "../include/g3cuda_spec.h"

namespace {

// ------------------------------------------------------------------ dual numbers
template <int K>
struct Dual {
    double v;
    double d[K];
    Dual() : v(0) { for (int i = 0; i < K; i++) d[i] = 0; }
    Dual(double x) : v(x) { for (int i = 0; i < K; i++) d[i] = 0; }
};
template <int K> Dual<K> operator+(const Dual<K>& a, const Dual<K>& b) { Dual<K> r; r.v = a.v + b.v; for (int i = 0; i < K; i++) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int K> Dual<K> operator-(const Dual<K>& a, const Dual<K>& b) { Dual<K> r; r.v = a.v - b.v; for (int i = 0; i < K; i++) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int K> Dual<K> operator-(const Dual<K>& a) { Dual<K> r; r.v = -a.v; for (int i = 0; i < K; i++) r.d[i] = -a.d[i]; return r; }
template <int K> Dual<K> operator*(const Dual<K>& a, const Dual<K>& b) { Dual<K> r; r.v = a.v * b.v; for (int i = 0; i < K; i++) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int K> Dual<K> operator/(const Dual<K>& a, const Dual<K>& b) { Dual<K> r; r.v = a.v / b.v; for (int i = 0; i < K; i++) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v; return r; }
template <int K> Dual<K> operator+(const Dual<K>& a, double b) { return a + Dual<K>(b); }
template <int K> Dual<K> operator+(double a, const Dual<K>& b) { return Dual<K>(a) + b; }
template <int K> Dual<K> operator-(const Dual<K>& a, double b) { return a - Dual<K>(b); }
template <int K> Dual<K> operator-(double a, const Dual<K>& b) { return Dual<K>(a) - b; }
template <int K> Dual<K> operator*(const Dual<K>& a, double b) { Dual<K> r; r.v = a.v * b; for (int i = 0; i < K; i++) r.d[i] = a.d[i] * b; return r; }
template <int K> Dual<K> operator*(double a, const Dual<K>& b) { return b * a; }
template <int K> Dual<K> operator/(const Dual<K>& a, double b) { return a / Dual<K>(b); }
template <int K> Dual<K> operator/(double a, const Dual<K>& b) { return Dual<K>(a) / b; }
template <int K> Dual<K>& operator+=(Dual<K>& a, const Dual<K>& b) { a = a + b; return a; }
template <int K> Dual<K>& operator-=(Dual<K>& a, const Dual<K>& b) { a = a - b; return a; }
template <int K> Dual<K>& operator*=(Dual<K>& a, const Dual<K>& b) { a = a * b; return a; }

inline double val(double x) { return x; }
template <int K> double val(const Dual<K>& x) { return x.v; }

// A direction the argument does not depend on stays exactly zero even where the derivative
// overflows (0 * inf): reverse-mode AD (TMBad, the reference's gradient) never forms that product.
template <int K, class F> Dual<K> chain(const Dual<K>& a, double f, F dfda) {
    Dual<K> r; r.v = f; double g = dfda; for (int i = 0; i < K; i++) r.d[i] = a.d[i] == 0.0 ? 0.0 : g * a.d[i]; return r;
}
inline double xexp(double a) { return std::exp(a); }
inline double xlog(double a) { return std::log(a); }
inline double xlog1p(double a) { return std::log1p(a); }
inline double xsqrt(double a) { return std::sqrt(a); }
inline double xpow(double a, double b) { return std::pow(a, b); }
inline double xlgamma(double a) { return std::lgamma(a); }
template <int K> Dual<K> xexp(const Dual<K>& a) { double f = std::exp(a.v); return chain(a, f, f); }
template <int K> Dual<K> xlog(const Dual<K>& a) { return chain(a, std::log(a.v), 1.0 / a.v); }
template <int K> Dual<K> xlog1p(const Dual<K>& a) { return chain(a, std::log1p(a.v), 1.0 / (1.0 + a.v)); }
template <int K> Dual<K> xsqrt(const Dual<K>& a) { double f = std::sqrt(a.v); return chain(a, f, 0.5 / f); }
template <int K> Dual<K> xpow(const Dual<K>& a, const Dual<K>& b) {
    double f = std::pow(a.v, b.v);
    Dual<K> r; r.v = f;
    double da = (b.v == 0.0) ? 0.0 : b.v * std::pow(a.v, b.v - 1.0);
    bool bconst = true; for (int i = 0; i < K; i++) if (b.d[i] != 0.0) bconst = false;
    double db = bconst ? 0.0 : f * std::log(a.v);
    for (int i = 0; i < K; i++) r.d[i] = da * a.d[i] + (bconst ? 0.0 : db * b.d[i]);
    return r;
}
template <int K> Dual<K> xpow(double a, const Dual<K>& b) { return xpow(Dual<K>(a), b); }
template <int K> Dual<K> xpow(const Dual<K>& a, double b) { return xpow(a, Dual<K>(b)); }
// digamma for the lgamma tangent (asymptotic series after upward recurrence)
inline double digamma(double x) {
    double r = 0;
    if (x <= 0 && std::floor(x) == x) return std::numeric_limits<double>::quiet_NaN();
    if (x < 0) return digamma(1 - x) - M_PI / std::tan(M_PI * x);
    while (x < 10) { r -= 1 / x; x += 1; }
    double f = 1 / (x * x);
    return r + std::log(x) - 0.5 / x -
           f * (1.0 / 12 - f * (1.0 / 120 - f * (1.0 / 252 - f * (1.0 / 240 - f * (1.0 / 132 - f * (691.0 / 32760 - f / 12))))));
}
template <int K> Dual<K> xlgamma(const Dual<K>& a) { return chain(a, std::lgamma(a.v), digamma(a.v)); }

// ------------------------------------------------------------------ g3_env primitives
// logspace_add: R/aab_env.R:16-19 (TMB's logspace_add, Rmath form)
template <class T> T logspace_add(const T& a, const T& b) {
    if (val(a) < val(b)) return b + xlog1p(xexp(a - b));
    return a + xlog1p(xexp(b - a));
}
// dif_pmax: R/env_dif.R:1-35 ; scale < 0 gives the smooth minimum (dif_pmin, R/env_dif.R:37-47)
template <class T> T dif_pmax(const T& a, double b, double scale) {
    return logspace_add(a * scale, T(b * scale)) / scale;
}
template <class T> T dif_pmin(const T& a, double b, double scale) { return dif_pmax(a, b, -scale); }
// avoid_zero: R/aab_env.R:24-31
template <class T> T avoid_zero(const T& a) { return dif_pmax(a, 0.0, 1e3); }
// dif_pminmax: R/env_dif.R:49-67
template <class T> T dif_pminmax(const T& a, double lower, double upper, double scale) {
    T out = dif_pmax(a, upper, -scale);
    return dif_pmax(out, lower, scale);
}
// TMB dnorm (density form used by R/action_renewal.R:66)
template <class T> T dnorm(double x, const T& mean, const T& sd) {
    T resid = (T(x) - mean) / sd;
    T logans = T(-std::log(std::sqrt(2 * M_PI))) - xlog(sd) - 0.5 * resid * resid;
    return xexp(logans);
}
// ratio_add_pop: R/aab_env.R:133-153
template <class T> T ratio_add_pop(const T& ov, const T& oa, const T& nv, const T& na) {
    return (ov * oa + nv * na) / avoid_zero(oa + na);
}

// ------------------------------------------------------------------ spec view
struct Spec {
    const int32_t* I; const double* R;
    int h(int k) const { return I[k]; }
    const int32_t* stock(int s) const { return I + I[G3H_STOCK_OFF] + s * G3S_LEN; }
    const int32_t* pred(int p) const { return I + I[G3H_PRED_OFF] + p * G3P_LEN; }
    const int32_t* pp(int p) const { return I + I[G3H_PP_OFF] + p * G3PP_LEN; }
    const int32_t* comp(int c) const { return I + I[G3H_COMP_OFF] + c * G3C_LEN; }
    const int32_t* xbuf(int x) const { return I + I[G3H_XBUF_OFF] + x * G3X_LEN; }
};

template <class T> struct ParVec {
    const double* p; const int32_t* tidx; int K;   // tangent seeds for Dual
};
inline double mkpar(const ParVec<double>& pv, int i) { return pv.p[i]; }
template <int K> Dual<K> mkpar(const ParVec<Dual<K>>& pv, int i) {
    Dual<K> r(pv.p[i]);
    for (int k = 0; k < pv.K && k < K; k++) if (pv.tidx[k] == i) r.d[k] = 1.0;
    return r;
}

// slot programs (include/g3cuda_spec.h G3OP_*)
template <class T> T eval_slot(const Spec& S, const ParVec<T>& pv, int slot) {
    const int32_t* tab = S.I + S.h(G3H_SLOT_OFF) + 2 * slot;
    const int32_t* c = S.I + tab[0];
    int n = tab[1];
    T st[G3_SLOT_STACK]; int sp = 0;
    for (int i = 0; i < n; i++) {
        switch (c[i]) {
        case G3OP_CONST: st[sp++] = T(S.R[c[++i]]); break;
        case G3OP_PARAM: st[sp++] = mkpar(pv, c[++i]); break;
        case G3OP_NAN:   st[sp++] = T(std::numeric_limits<double>::quiet_NaN()); break;
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
        }
    }
    return st[0];
}

// ------------------------------------------------------------------ growth natives
// g3a_grow_lengthvbsimple: R/action_grow.R:44-55
template <class T> T grow_lengthvbsimple(const T& linf, const T& kappa, double midlen, double cur_step_size) {
    return avoid_zero((linf - midlen) * (1.0 - xexp(-(kappa) * cur_step_size)));
}
// Diagnostic switch (tests only): evaluate growth_bbinom in the lgamma-free product form the CUDA
// kernel uses, to measure how far the two formulations are apart.
int g_bbinom_product = 0;
inline double xabs(double a) { return std::fabs(a); }
template <int K> Dual<K> xabs(const Dual<K>& a) { return chain(a, std::fabs(a.v), a.v < 0 ? -1.0 : 1.0); }

// growth_bbinom: R/action_grow.R:155-213. out[l + L*x], column-major L x (n+1)
template <class T> void growth_bbinom(const std::vector<T>& delt_l, int binn, const T& beta, std::vector<T>& out) {
    int na = (int)delt_l.size(), n = binn;
    out.assign((size_t)na * (n + 1), T(0.0));
    if (g_bbinom_product) {
        // C(n,x) prod_{i<x}|alpha+i| prod_{j<n-x}(beta+j) / prod_{k<n}|alpha+beta+k|
        for (int l = 0; l < na; l++) {
            T alpha = (beta * delt_l[l]) / (double(binn) - delt_l[l]);
            T g(1.0);
            for (int j = 0; j < n; j++) g = g * ((beta + double(j)) / xabs(alpha + beta + double(j)));
            for (int x = 0; x <= n; x++) {
                out[l + na * x] = g;
                if (x < n) g = g * (double(n - x) / double(x + 1)) * (xabs(alpha + double(x)) / (beta + double(n - x - 1)));
            }
        }
        return;
    }
    for (int i = 0; i < na * (n + 1); i++) {
        T alpha = (beta * delt_l[i % na]) / (double(binn) - delt_l[i % na]);
        double x = double(i / na);
        out[i] = xexp(T(std::lgamma(double(n) + 1)) + xlgamma(alpha + beta) + xlgamma(T(n - x) + beta) + xlgamma(T(x) + alpha) -
                      T(std::lgamma(n - x + 1)) - T(std::lgamma(x + 1)) - xlgamma(T(double(n)) + alpha + beta) - xlgamma(beta) - xlgamma(alpha));
    }
}
// g3a_grow_matrix_len: R/action_grow.R:233-271. delta L x D (col-major) -> dense L x L (row lg, col dest)
template <class T> void grow_matrix_len(const std::vector<T>& delta, int L, int D, std::vector<T>& M) {
    M.assign((size_t)L * L, T(0.0));
    for (int lg = 0; lg < L; lg++) {
        if (lg == L - 1) {
            T s(0.0); for (int x = 0; x < D; x++) s = s + delta[lg + L * x];
            M[lg + L * lg] = s;
        } else if (lg + D > L) {
            for (int x = 0; x < L - lg; x++) M[lg + L * (lg + x)] = delta[lg + L * x];
            T s(0.0); for (int x = L - lg - 1; x < D; x++) s = s + delta[lg + L * x];   // tail(total_deltas - (L - lg) + 1)
            M[lg + L * (L - 1)] = s;
        } else {
            for (int x = 0; x < D; x++) M[lg + L * (lg + x)] = delta[lg + L * x];
        }
    }
}
// g3a_grow_matrix_wgt: R/action_grow.R:273-307
template <class T> void grow_matrix_wgt(const std::vector<T>& delta, int L, int D, std::vector<T>& M) {
    M.assign((size_t)L * L, T(0.0));
    for (int lg = 0; lg < L; lg++) {
        int w = (lg == L - 1 || lg + D > L) ? L - lg : D;
        for (int x = 0; x < w; x++) M[lg + L * (lg + x)] = delta[lg + L * x];
    }
}
// g3a_grow_apply: R/action_grow.R:309-335
template <class T> void grow_apply(const std::vector<T>& G, const std::vector<T>& W, int L, const T* num, const T* wgt,
                                   std::vector<T>& out_num, std::vector<T>& out_wgt) {
    std::vector<T> g((size_t)L * L), w((size_t)L * L);
    for (int j = 0; j < L; j++) for (int l = 0; l < L; l++) {
        g[l + L * j] = G[l + L * j] * num[l];
        w[l + L * j] = (W[l + L * j] + wgt[l]) * g[l + L * j];
    }
    out_num.assign(L, T(0.0)); out_wgt.assign(L, T(0.0));
    for (int j = 0; j < L; j++) {
        T sn(0.0), sw(0.0);
        for (int l = 0; l < L; l++) { sn = sn + g[l + L * j]; sw = sw + w[l + L * j]; }
        out_num[j] = sn;
        out_wgt[j] = sw / avoid_zero(sn);
    }
}

// g3_suitability_* of prey length `ml` for predator length `pl` (R/suitability.R:5-30)
template <class T, class F> T suitability(int kind, const int32_t* ss, F& slot, double ml, double pl, bool& ok) {
    if (kind == G3SUIT_EXPL50) return 1.0 / (1.0 + xexp(-slot(ss[0]) * (T(ml) - slot(ss[1]))));
    if (kind == G3SUIT_CONSTANT) return slot(ss[0]);
    if (kind == G3SUIT_ANDERSEN) {
        T p0 = slot(ss[0]), p1 = slot(ss[1]), p2 = slot(ss[2]), p3 = slot(ss[3]), p4 = slot(ss[4]);
        T x = xlog(avoid_zero(T(pl / ml)));
        T d = x - p1;
        // bounded_vec(y, 0, 1) = 0 + (1 - 0) / (1 + exp(y))  (R/aab_env.R:51-55)
        T lo = 1.0 / (1.0 + xexp(1000.0 * (p1 - x)));
        T hi = 1.0 / (1.0 + xexp(1000.0 * (x - p1)));
        return p0 + avoid_zero(p2) * xexp(-(d * d) / avoid_zero(p3)) * lo + avoid_zero(p2) * xexp(-(d * d) / avoid_zero(p4)) * hi;
    }
    if (kind == G3SUIT_ANDERSENFLEET) {         // pl = the prey's largest midlen (p5 = stock__maxmidlen); R/suitability.R:32-58
        T p0 = slot(ss[0]), p1 = slot(ss[1]), p2 = slot(ss[2]), p3 = slot(ss[3]), p4 = slot(ss[4]);
        T x = xlog(T(pl / ml));
        T d = x - p1;
        T lo = 1.0 / (1.0 + xexp(1000.0 * (p1 - x)));
        T hi = 1.0 / (1.0 + xexp(1000.0 * (x - p1)));
        return p0 + p2 * xexp(-(d * d) / p3) * lo + p2 * xexp(-(d * d) / p4) * hi;
    }
    if (kind == G3SUIT_GAMMA) {                 // R/suitability.R:60-67
        T a = slot(ss[0]), b = slot(ss[1]), g = slot(ss[2]);
        return xpow(T(ml) / ((a - 1.0) * b * g), a - 1.0) * xexp(a - 1.0 - T(ml) / (b * g));
    }
    if (kind == G3SUIT_EXPONENTIAL || kind == G3SUIT_RICHARDS) {       // R/suitability.R:69-75,90-94
        T a = slot(ss[0]), b = slot(ss[1]), g = slot(ss[2]), dl = slot(ss[3]);
        T e = dl / (1.0 + xexp(-a - b * ml - g * pl));
        return kind == G3SUIT_RICHARDS ? xpow(e, 1.0 / slot(ss[4])) : e;
    }
    if (kind == G3SUIT_STRAIGHTLINE) return slot(ss[0]) + slot(ss[1]) * ml;        // R/suitability.R:77-79
    ok = false;
    return T(0.0);
}
// g3a_migrate_normalize: R/action_migrate.R:2-15 (row_total = 1)
template <class T> void migrate_normalize(std::vector<T>& vec, int src_idx, double row_total) {
    vec[src_idx] = T(0.0);
    T tot(0.0);
    for (auto& v : vec) { v = v * v; tot = tot + v; }
    vec[src_idx] = row_total - tot;
    for (auto& v : vec) v = v / row_total;
}

// ------------------------------------------------------------------ the model run
struct ReportSink {
    double* nll;          // 1
    double* steps;        // [n_nll][NT]
    std::vector<double*> dstart_num, dstart_wgt;   // per stock [NT][cells]
    std::vector<double*> model;                    // per comp [n_times][n_dest]
    std::vector<double*> params;                   // per comp [2]
    std::vector<double*> renewalnum;               // per stock [NT][cells]   (detail_<stock>__renewalnum)
    std::vector<double*> suit_report;              // per predator-prey pair [predator lengths (1 for a fleet)][L]
    std::vector<double*> dsuit, dcons;             // per pair [NT][prey cells] (detail_<prey>_<pred>__suit/__cons)
};

template <class T>
T run_model(const Spec& S, const ParVec<T>& pv, double* comp_out /*n_nll or null*/, ReportSink* rep) {
    const double NaN = std::numeric_limits<double>::quiet_NaN();
    const int NS = S.h(G3H_NSTOCK), NC = S.h(G3H_NCELL), NP = S.h(G3H_NPRED), NPP = S.h(G3H_NPP);
    const int NCOMP = S.h(G3H_NCOMP), NX = S.h(G3H_NXBUF), NSTEP = S.h(G3H_NSTEPS), NT = S.h(G3H_NT);
    const int NNLL = S.h(G3H_NNLL);
    const int32_t* steplen = S.I + S.h(G3H_STEPLEN_OFF);
    auto slot = [&](int id) { return eval_slot<T>(S, pv, id); };

    // g3a_time (R/action_time.R:75-91)
    double retro = S.h(G3H_PAR_RETRO) >= 0 ? pv.p[S.h(G3H_PAR_RETRO)] : 0.0;
    double proj = S.h(G3H_PAR_PROJECT) >= 0 ? pv.p[S.h(G3H_PAR_PROJECT)] : 0.0;
    if (!(retro >= 0) || !(proj >= 0)) return T(NaN);
    int total_steps = NSTEP * (S.h(G3H_END_YEAR) - (int)std::floor(retro) - S.h(G3H_START_YEAR) + (int)std::floor(proj)) + NSTEP - 1;
    if (total_steps + 1 > NT) return T(NaN);   // projections need tables beyond the packed range

    std::vector<T> num(NC, T(0.0)), wgt(NC, T(1.0));
    std::vector<T> totalpredate(NC), consratio(NC), consconv(NC);
    std::vector<std::vector<T>> suit(NPP, std::vector<T>(NC, T(0.0))), cons(NPP, std::vector<T>(NC, T(0.0)));
    std::vector<T> overcons(NS, T(0.0));
    std::vector<T> renewalnum(NC, T(0.0));       // <stock>__renewalnum: only ever overwritten by a renewal
    std::vector<std::vector<T>> trans_num(NS), trans_wgt(NS), growth_l(NS), growth_w(NS);
    std::vector<int> lastcalc(NS, -1);
    std::vector<std::vector<T>> mv_num(NS), mv_wgt(NS);
    std::vector<std::vector<T>> model(NCOMP);
    std::vector<double> nllsteps((size_t)NNLL * NT, 0.0);
    std::vector<T> comp_tot(NNLL, T(0.0));
    for (int c = 0; c < NCOMP; c++) {
        const int32_t* C = S.comp(c);
        bool si = C[G3C_FUNC] == G3F_SI_LOG || C[G3C_FUNC] == G3F_SI_LINEAR;
        model[c].assign((size_t)C[G3C_N_DEST] * (si || rep ? C[G3C_N_TIMES] : 1), T(0.0));
    }
    std::vector<std::vector<T>> si_params(NCOMP, std::vector<T>(2, T(0.0)));
    T nll(0.0);

    for (int cur_time = 0;; cur_time++) {
        int cur_year = S.h(G3H_START_YEAR) + cur_time / NSTEP;
        int cur_step = cur_time % NSTEP + 1;
        double cur_step_size = steplen[cur_step - 1] / 12.0;
        bool cur_step_final = cur_step == NSTEP;
        int yidx = cur_year - S.h(G3H_START_YEAR);

        // ---- g3a_initialconditions (R/action_renewal.R:83-103), cur_time == 0
        if (cur_time == 0) for (int s = 0; s < NS; s++) {
            const int32_t* St = S.stock(s);
            if (St[G3S_INIT_OFF] < 0) continue;
            int L = St[G3S_L]; const double* midlen = S.R + St[G3S_MIDLEN_ROFF];
            for (int col = 0; col < St[G3S_NAREA] * St[G3S_NAGE]; col++) {
                const int32_t* sl = S.I + St[G3S_INIT_OFF] + col * 5;
                T mean = slot(sl[0]), sd = avoid_zero(slot(sl[1])), factor = slot(sl[2]), wa = slot(sl[3]), wb = slot(sl[4]);
                std::vector<T> d(L); T tot(0.0);
                for (int l = 0; l < L; l++) { d[l] = dnorm(midlen[l], mean, sd); tot = tot + d[l]; }
                T den = avoid_zero(tot);                                    // normalize_vec, R/aab_env.R:37-41
                for (int l = 0; l < L; l++) {
                    int c = St[G3S_CELL_OFF] + col * L + l;
                    num[c] = d[l] / den * 10000.0 * factor;
                    wgt[c] = wa * xpow(midlen[l], wb);
                }
            }
        }
        // ---- g3a_otherfood (R/action_renewal.R:276-308): number and mean weight assigned from formulas, every step
        if (cur_time <= total_steps) for (int s = 0; s < NS; s++) {
            const int32_t* St = S.stock(s);
            if (St[G3S_OTHERFOOD_OFF] < 0) continue;
            int L = St[G3S_L], A = St[G3S_NAGE], NR = St[G3S_NAREA];
            for (int r = 0; r < NR; r++) {
                const int32_t* sl = S.I + St[G3S_OTHERFOOD_OFF] + ((size_t)cur_time * NR + r) * 2;
                T n = slot(sl[0]), w = slot(sl[1]);
                for (int a = 0; a < A; a++) for (int l = 0; l < L; l++) {
                    int c = St[G3S_CELL_OFF] + (r * A + a) * L + l;
                    num[c] = n; wgt[c] = w;
                }
            }
        }
        // ---- report dstart_* (R/action_report.R:177-213)
        if (rep && cur_time <= total_steps) for (int s = 0; s < NS; s++) {
            const int32_t* St = S.stock(s);
            int n = St[G3S_L] * St[G3S_NAGE] * St[G3S_NAREA];
            for (int i = 0; i < n; i++) {
                rep->dstart_num[s][(size_t)cur_time * n + i] = val(num[St[G3S_CELL_OFF] + i]);
                rep->dstart_wgt[s][(size_t)cur_time * n + i] = val(wgt[St[G3S_CELL_OFF] + i]);
            }
        }
        if (cur_time > total_steps) break;

        // ---- g3a_migrate (R/action_migrate.R:17-82), action order 2
        for (int s = 0; s < NS; s++) {
            const int32_t* St = S.stock(s);
            if (St[G3S_MIGRATE_OFF] < 0) continue;
            int L = St[G3S_L], A = St[G3S_NAGE], NR = St[G3S_NAREA], c0 = St[G3S_CELL_OFF];
            const int32_t* mig = S.I + St[G3S_MIGRATE_OFF] + (cur_step - 1) * NR * NR;       // [dest][src] slots
            bool runs = false;
            for (int i = 0; i < NR * NR; i++) if (mig[i] >= 0) runs = true;
            if (!runs) continue;        // "Copy not-migrating stocks"
            std::vector<T> pn((size_t)L * A * NR, T(0.0)), pw((size_t)L * A * NR, T(0.0));
            for (int src = 0; src < NR; src++) {
                std::vector<T> vec(NR);
                for (int d = 0; d < NR; d++) vec[d] = mig[d * NR + src] >= 0 ? slot(mig[d * NR + src]) : T(0.0);
                migrate_normalize(vec, src, 1.0);
                for (int a = 0; a < A; a++) for (int d = 0; d < NR; d++) for (int l = 0; l < L; l++) {
                    int cs = (src * A + a) * L + l, cd = (d * A + a) * L + l;
                    T moved = num[c0 + cs] * vec[d];
                    pw[cd] = ratio_add_pop(pw[cd], pn[cd], wgt[c0 + cs], moved);       // stock_combine_subpop
                    pn[cd] = pn[cd] + moved;
                }
            }
            for (size_t i = 0; i < pn.size(); i++) { num[c0 + i] = pn[i]; wgt[c0 + i] = pw[i]; }
        }

        // ---- g3a_predate (R/action_predate.R:240-420)
        std::fill(totalpredate.begin(), totalpredate.end(), T(0.0));
        std::vector<std::vector<T>> totalsuit(NP);
        for (int p = 0; p < NP; p++) totalsuit[p].assign(S.pred(p)[G3P_NAREA], T(0.0));
        std::vector<char> is_prey(NS, 0);
        // stock predators: total suitability per (predator area idx, predator length), R/action_predate.R:207-238
        std::vector<std::vector<T>> ptotsuit(NP);
        for (int p = 0; p < NP; p++)
            if (S.pred(p)[G3P_KIND] == G3PRED_STOCK) {
                const int32_t* Sp = S.stock(S.pred(p)[G3P_STOCK]);
                ptotsuit[p].assign((size_t)Sp[G3S_NAREA] * Sp[G3S_L], T(0.0));
            }
        for (int q = 0; q < NPP; q++) {     // step 1: collect suitable biomass
            const int32_t* Q = S.pp(q); const int32_t* P = S.pred(Q[G3PP_PRED]); const int32_t* St = S.stock(Q[G3PP_PREY]);
            is_prey[Q[G3PP_PREY]] = 1;
            int L = St[G3S_L]; const double* midlen = S.R + St[G3S_MIDLEN_ROFF];
            const int32_t* ss = S.I + Q[G3PP_SUIT_OFF] + 6 * (Q[G3PP_SUIT_TMAP] >= 0 ? S.I[Q[G3PP_SUIT_TMAP] + cur_time] : 0);       // (time-varying selectivity: this step's set)
            bool ok = true;
            if (P[G3P_KIND] == G3PRED_STOCK) {
                const int32_t* Sp = S.stock(P[G3P_STOCK]);
                int PL = Sp[G3S_L]; const double* pmid = S.R + Sp[G3S_MIDLEN_ROFF];
                T energy = slot(Q[G3PP_ENERGY]);
                for (int r = 0; r < St[G3S_NAREA]; r++) {
                    int area = S.I[St[G3S_AREA_OFF] + r], pr = -1;
                    for (int k = 0; k < Sp[G3S_NAREA]; k++) if (S.I[Sp[G3S_AREA_OFF] + k] == area) pr = k;
                    if (pr < 0) continue;
                    for (int pl = 0; pl < PL; pl++)
                        for (int a = 0; a < St[G3S_NAGE]; a++) {
                            T colsum(0.0);
                            for (int l = 0; l < L; l++) {
                                int c = St[G3S_CELL_OFF] + (r * St[G3S_NAGE] + a) * L + l;
                                T su = suitability<T>(Q[G3PP_SUIT_KIND], ss, slot, midlen[l], Q[G3PP_SUIT_KIND] == G3SUIT_ANDERSENFLEET ? midlen[L - 1] : pmid[pl], ok) * energy * num[c] * wgt[c];
                                if (Q[G3PP_PREF] >= 0) su = xpow(su, slot(Q[G3PP_PREF]));        // (...)^prey_preference, R/action_predate.R:220-225
                                colsum = colsum + su;
                            }
                            ptotsuit[Q[G3PP_PRED]][pr * PL + pl] = ptotsuit[Q[G3PP_PRED]][pr * PL + pl] + colsum;
                        }
                }
                if (!ok) return T(NaN);
                continue;
            }
            std::vector<T> suitab(L);
            for (int l = 0; l < L; l++) suitab[l] = suitability<T>(Q[G3PP_SUIT_KIND], ss, slot, midlen[l], midlen[L - 1], ok);
            // (suitabilities that read predator_length need a predator with lengths)
            if (!ok || Q[G3PP_SUIT_KIND] == G3SUIT_ANDERSEN || Q[G3PP_SUIT_KIND] == G3SUIT_EXPONENTIAL || Q[G3PP_SUIT_KIND] == G3SUIT_RICHARDS) return T(NaN);
            std::fill(suit[q].begin(), suit[q].end(), T(0.0));
            for (int r = 0; r < St[G3S_NAREA]; r++) {
                int area = S.I[St[G3S_AREA_OFF] + r], pr = -1;
                for (int k = 0; k < P[G3P_NAREA]; k++) if (S.I[P[G3P_AREA_OFF] + k] == area) pr = k;
                if (pr < 0) continue;
                for (int a = 0; a < St[G3S_NAGE]; a++) {
                    T colsum(0.0);
                    for (int l = 0; l < L; l++) {
                        int c = St[G3S_CELL_OFF] + (r * St[G3S_NAGE] + a) * L + l;
                        suit[q][c] = P[G3P_KIND] == G3PRED_NUMBERFLEET ? suitab[l] * num[c] : suitab[l] * num[c] * wgt[c];
                        colsum = colsum + suit[q][c];
                    }
                    totalsuit[Q[G3PP_PRED]][pr] = totalsuit[Q[G3PP_PRED]][pr] + colsum;
                }
            }
        }
        for (int q = 0; q < NPP; q++) {     // step 2: scale by total expected catch
            const int32_t* Q = S.pp(q); const int32_t* P = S.pred(Q[G3PP_PRED]); const int32_t* St = S.stock(Q[G3PP_PREY]);
            int L = St[G3S_L];
            std::fill(cons[q].begin(), cons[q].end(), T(0.0));
            if (P[G3P_KIND] == G3PRED_STOCK) {
                // cons = predN * max_consumption * feeding_level * suit / (energycontent * total_predsuit)
                const int32_t* Sp = S.stock(P[G3P_STOCK]);
                int PL = Sp[G3S_L], PA = Sp[G3S_NAGE]; const double* pmid = S.R + Sp[G3S_MIDLEN_ROFF];
                const double* midlen = S.R + St[G3S_MIDLEN_ROFF];
                const int32_t* ss = S.I + Q[G3PP_SUIT_OFF] + 6 * (Q[G3PP_SUIT_TMAP] >= 0 ? S.I[Q[G3PP_SUIT_TMAP] + cur_time] : 0);
                const int32_t* ps = S.I + P[G3P_SLOT_OFF];
                T energy = slot(Q[G3PP_ENERGY]);
                T m0 = slot(ps[0]), m1 = slot(ps[1]), m2 = slot(ps[2]), m3 = slot(ps[3]), hf = slot(ps[4]), temp = slot(ps[5]);
                bool ok = true;
                for (int r = 0; r < St[G3S_NAREA]; r++) {
                    int area = S.I[St[G3S_AREA_OFF] + r], pr = -1;
                    for (int k = 0; k < Sp[G3S_NAREA]; k++) if (S.I[Sp[G3S_AREA_OFF] + k] == area) pr = k;
                    if (pr < 0) continue;
                    for (int pl = 0; pl < PL; pl++) {
                        T ts = ptotsuit[Q[G3PP_PRED]][pr * PL + pl];
                        T maxcons = m0 * cur_step_size * xexp(m1 * temp - m2 * temp * temp * temp) * xpow(T(pmid[pl]), m3);     // R/action_predate.R:190-205
                        T feeding = ts / (hf * cur_step_size + ts);
                        for (int pa = 0; pa < PA; pa++) {
                            T predn = num[Sp[G3S_CELL_OFF] + (pr * PA + pa) * PL + pl];
                            for (int a = 0; a < St[G3S_NAGE]; a++) for (int l = 0; l < L; l++) {
                                int c = St[G3S_CELL_OFF] + (r * St[G3S_NAGE] + a) * L + l;
                                T su = suitability<T>(Q[G3PP_SUIT_KIND], ss, slot, midlen[l], Q[G3PP_SUIT_KIND] == G3SUIT_ANDERSENFLEET ? midlen[L - 1] : pmid[pl], ok) * energy * num[c] * wgt[c];
                                if (Q[G3PP_PREF] >= 0) su = xpow(su, slot(Q[G3PP_PREF]));
                                T cn = (predn * maxcons * feeding * su) / (energy * ts);
                                cons[q][c] = cons[q][c] + cn;
                                totalpredate[c] = totalpredate[c] + cn;
                            }
                        }
                    }
                }
                continue;
            }
            for (int r = 0; r < St[G3S_NAREA]; r++) {
                int area = S.I[St[G3S_AREA_OFF] + r], pr = -1;
                for (int k = 0; k < P[G3P_NAREA]; k++) if (S.I[P[G3P_AREA_OFF] + k] == area) pr = k;
                if (pr < 0) continue;
                double E = S.R[P[G3P_E_ROFF] + (size_t)cur_time * P[G3P_NAREA] + pr];
                for (int a = 0; a < St[G3S_NAGE]; a++) for (int l = 0; l < L; l++) {
                    int c = St[G3S_CELL_OFF] + (r * St[G3S_NAGE] + a) * L + l;
                    // g3a_predate_catchability_project / _effortfleet (R/action_predate.R:1-126)
                    if (P[G3P_KIND] == G3PRED_TOTALFLEET) cons[q][c] = E * (suit[q][c] / totalsuit[Q[G3PP_PRED]][pr]);
                    else if (P[G3P_KIND] == G3PRED_NUMBERFLEET) cons[q][c] = E * (suit[q][c] / totalsuit[Q[G3PP_PRED]][pr]) * wgt[c];
                    else if (P[G3P_KIND] == G3PRED_LINEARFLEET) cons[q][c] = suit[q][c] * E * cur_step_size;
                    else cons[q][c] = slot(Q[G3PP_ENERGY]) * E * cur_step_size * suit[q][c];
                    totalpredate[c] = totalpredate[c] + cons[q][c];
                }
            }
        }
        for (int s = 0; s < NS; s++) {      // step 4: overconsumption (R/action_predate.R:381-406)
            if (!is_prey[s]) continue;
            const int32_t* St = S.stock(s);
            int c0 = St[G3S_CELL_OFF], n = St[G3S_L] * St[G3S_NAGE] * St[G3S_NAREA];
            T o(0.0), o2(0.0);
            for (int c = c0; c < c0 + n; c++) {
                consratio[c] = totalpredate[c] / avoid_zero(num[c] * wgt[c]);
                consratio[c] = dif_pmin(consratio[c], 0.95, 1e3);
            }
            for (int c = c0; c < c0 + n; c++) o = o + totalpredate[c];
            for (int c = c0; c < c0 + n; c++) {
                consconv[c] = 1.0 / avoid_zero(totalpredate[c]);
                totalpredate[c] = (num[c] * wgt[c]) * consratio[c];
            }
            for (int c = c0; c < c0 + n; c++) o2 = o2 + totalpredate[c];
            overcons[s] = o - o2;
            for (int c = c0; c < c0 + n; c++) {
                consconv[c] = consconv[c] * totalpredate[c];
                num[c] = num[c] * (1.0 - consratio[c]);
            }
        }
        for (int q = 0; q < NPP; q++) {     // step 5
            const int32_t* St = S.stock(S.pp(q)[G3PP_PREY]);
            int c0 = St[G3S_CELL_OFF], n = St[G3S_L] * St[G3S_NAGE] * St[G3S_NAREA];
            for (int c = c0; c < c0 + n; c++) cons[q][c] = cons[q][c] * consconv[c];
        }

        // ---- g3a_naturalmortality (R/action_naturalmortality.R:26)
        for (int s = 0; s < NS; s++) {
            const int32_t* St = S.stock(s);
            if (St[G3S_MORT_OFF] < 0) continue;
            int L = St[G3S_L], ncol = St[G3S_NAREA] * St[G3S_NAGE];
            // time-varying M (by_year / by_step tables, g3_timevariable): the slot set of this step
            int mset = St[G3S_MORT_TMAP_OFF] >= 0 ? S.I[St[G3S_MORT_TMAP_OFF] + cur_time] : 0;
            for (int col = 0; col < ncol; col++) {
                T f = xexp(-slot(S.I[St[G3S_MORT_OFF] + mset * ncol + col]) * cur_step_size);
                for (int l = 0; l < L; l++) { int c = St[G3S_CELL_OFF] + col * L + l; num[c] = num[c] * f; }
            }
        }

        // ---- g3a_growmature (R/action_grow.R:340-497)
        for (int s = 0; s < NS; s++) {
            const int32_t* St = S.stock(s);
            int n = St[G3S_GROW_N];
            if (n < 0) continue;
            int L = St[G3S_L], D = n + 1, ncell = L * St[G3S_NAGE] * St[G3S_NAREA], c0 = St[G3S_CELL_OFF];
            const double* midlen = S.R + St[G3S_MIDLEN_ROFF];
            double plusdl = S.R[St[G3S_PLUSDL_ROFF]];
            bool transition = St[G3S_N_OUT] > 0 && ((St[G3S_TRANS_MASK] >> (cur_step - 1)) & 1);
            if (transition) {       // "Reset transitioning arrays" (R/action_grow.R:376-381)
                trans_num[s].assign(ncell, T(0.0));
                trans_wgt[s].assign(wgt.begin() + c0, wgt.begin() + c0 + ncell);
            }
            // time-varying growth parameters: the slot set of this step; the matrices are recomputed when the step length
            // OR the set changes (the reference keys its cache on what the formulas depend on, R/action_grow.R:391-410)
            const int gset = St[G3S_GROW_TMAP_OFF] >= 0 ? S.I[St[G3S_GROW_TMAP_OFF] + cur_time] : 0;
            const int32_t* gs = S.I + St[G3S_GROW_OFF] + 5 * gset;
            int calc = (int)std::floor(cur_step_size * 12) + 100 * gset;
            if (lastcalc[s] != calc) {
                T linf = slot(gs[0]), kappa = slot(gs[1]), beta = slot(gs[2]), wa = slot(gs[3]), wb = slot(gs[4]);
                std::vector<T> dl(L);
                for (int l = 0; l < L; l++)     // g3a_grow_lengthvbsimple + bbinom wrapper (R/action_grow.R:54,220)
                    dl[l] = avoid_zero(grow_lengthvbsimple(linf, kappa, midlen[l], cur_step_size) / plusdl);
                growth_bbinom(dl, n, avoid_zero(beta), growth_l[s]);
                growth_w[s].assign((size_t)L * D, T(0.0));
                std::vector<T> pw(L);
                for (int l = 0; l < L; l++) pw[l] = xpow(midlen[l], wb);
                for (int l = 0; l < L; l++) for (int x = 0; x < D; x++)     // vec_rotate - vec_extrude (R/action_grow.R:1-41,67-70)
                    growth_w[s][l + L * x] = (pw[(l + x < L) ? l + x : L - 1] - pw[l]) * wa;
            }
            std::vector<T> Wm; grow_matrix_wgt(growth_w[s], L, D, Wm);
            const int32_t* ms = St[G3S_MAT_KIND] != G3MAT_NONE ? S.I + St[G3S_MAT_OFF] + (St[G3S_MAT_KIND] == G3MAT_CONTINUOUS ? 4 * gset : 0) : nullptr;
            std::vector<T> on, ow;
            for (int col = 0; col < St[G3S_NAREA] * St[G3S_NAGE]; col++) {
                int age = St[G3S_MINAGE] + col % St[G3S_NAGE];
                T* cn = &num[c0 + col * L]; T* cw = &wgt[c0 + col * L];
                if (transition && St[G3S_MAT_KIND] == G3MAT_CONTINUOUS) {
                    // maturity_ratio depends on growth_delta_l: two applies (R/action_grow.R:420-437)
                    T alpha = slot(ms[0]), l50 = slot(ms[1]);
                    T beta = ms[2] >= 0 ? slot(ms[2]) : T(0.0);
                    std::vector<T> gm((size_t)L * D), gi((size_t)L * D);
                    for (int l = 0; l < L; l++) {
                        T inner = T(0.0) - alpha * (T(midlen[l]) - l50);         // g3a_mature_constant (R/action_mature.R:46-72)
                        if (ms[2] >= 0) inner = inner - beta * (T(double(age)) - slot(ms[3]));
                        T m0 = 1.0 / (1.0 + xexp(inner));
                        for (int x = 0; x < D; x++) {
                            T ratio = dif_pminmax(m0 * (alpha * (plusdl * x) + beta * cur_step_size), 0.0, 1.0, 1e5);
                            gm[l + L * x] = growth_l[s][l + L * x] * ratio;
                            gi[l + L * x] = growth_l[s][l + L * x] * (1.0 - ratio);
                        }
                    }
                    std::vector<T> Gm, Gi; grow_matrix_len(gm, L, D, Gm); grow_matrix_len(gi, L, D, Gi);
                    std::vector<T> tn, tw;
                    grow_apply(Gm, Wm, L, cn, cw, tn, tw);
                    grow_apply(Gi, Wm, L, cn, cw, on, ow);
                    for (int l = 0; l < L; l++) { trans_num[s][col * L + l] = tn[l]; trans_wgt[s][col * L + l] = tw[l]; cn[l] = on[l]; cw[l] = ow[l]; }
                } else if (transition && St[G3S_MAT_KIND] == G3MAT_CONSTANT) {
                    // num * ratio / num * (1 - ratio) with one growth matrix (R/action_grow.R:438-457)
                    std::vector<T> G; grow_matrix_len(growth_l[s], L, D, G);
                    std::vector<T> nm(L), ni(L);
                    for (int l = 0; l < L; l++) {
                        T inner(0.0);
                        if (ms[0] >= 0) inner = inner - slot(ms[0]) * (T(midlen[l]) - slot(ms[1]));
                        if (ms[2] >= 0) inner = inner - slot(ms[2]) * (T(double(age)) - slot(ms[3]));
                        T ratio = 1.0 / (1.0 + xexp(inner));
                        nm[l] = cn[l] * ratio; ni[l] = cn[l] * (1.0 - ratio);
                    }
                    std::vector<T> tn, tw;
                    grow_apply(G, Wm, L, nm.data(), cw, tn, tw);
                    grow_apply(G, Wm, L, ni.data(), cw, on, ow);
                    for (int l = 0; l < L; l++) { trans_num[s][col * L + l] = tn[l]; trans_wgt[s][col * L + l] = tw[l]; cn[l] = on[l]; cw[l] = ow[l]; }
                } else {
                    std::vector<T> G; grow_matrix_len(growth_l[s], L, D, G);
                    grow_apply(G, Wm, L, cn, cw, on, ow);
                    for (int l = 0; l < L; l++) { cn[l] = on[l]; cw[l] = ow[l]; }
                }
            }
            lastcalc[s] = calc;
        }
        // ---- g3a_spawn, order 6 (R/action_spawn.R:121-168): spawning parents, total offspring, spawning mortality
        std::vector<T> spawn_total(NS, T(0.0));
        std::vector<char> spawn_now(NS, 0);
        for (int s = 0; s < NS; s++) {
            const int32_t* St = S.stock(s);
            if (St[G3S_SPAWN_OFF] < 0) continue;
            const int32_t* Sp = S.I + St[G3S_SPAWN_OFF];
            if (Sp[G3SP_RUN_STEP] != 0 && Sp[G3SP_RUN_STEP] != cur_step) continue;
            spawn_now[s] = 1;
            int L = St[G3S_L], A = St[G3S_NAGE], NR = St[G3S_NAREA], ncell = L * A * NR;
            const double* midlen = S.R + St[G3S_MIDLEN_ROFF];
            const int32_t* rs = Sp + G3SP_RECR_SLOTS;
            const int kind = Sp[G3SP_RECR_KIND];
            bool ok = true;
            T sum(0.0);
            for (int r = 0; r < NR; r++) for (int a = 0; a < A; a++) for (int l = 0; l < L; l++) {
                int c = St[G3S_CELL_OFF] + (r * A + a) * L + l;
                T prop = suitability<T>(Sp[G3SP_PROP_KIND], Sp + G3SP_PROP_SLOTS, slot, midlen[l], midlen[l], ok);
                T spn = num[c] * prop;          // stock__spawningnum
                if (kind == G3SPAWN_FECUNDITY)
                    sum = sum + xpow(T(midlen[l]), slot(rs[1])) * xpow(T((double)(St[G3S_MINAGE] + a)), slot(rs[2])) * xpow(spn, slot(rs[3])) * xpow(wgt[c], slot(rs[4]));
                else
                    sum = sum + wgt[c] * spn;
                if (Sp[G3SP_MORT_KIND] >= 0) {  // g3a_naturalmortality_exp(mortality_f, action_step_size_f = 1)
                    T m = suitability<T>(Sp[G3SP_MORT_KIND], Sp + G3SP_MORT_SLOTS, slot, midlen[l], midlen[l], ok);
                    num[c] = num[c] - spn * (1.0 - xexp(-m));
                }
                if (Sp[G3SP_WLOSS] >= 0) {      // g3a_weightloss on the spawning part (R/action_spawn.R:166-181, R/action_weightloss.R:19-44)
                    T mw = slot(Sp[G3SP_WLOSS + 1]);
                    T lost = ((wgt[c] - mw) * (1.0 - slot(Sp[G3SP_WLOSS]))) + mw;
                    wgt[c] = ratio_add_pop(wgt[c], num[c] - spn, lost, spn);
                }
            }
            if (!ok) return T(NaN);
            T r(0.0);
            if (kind == G3SPAWN_FECUNDITY || kind == G3SPAWN_SIMPLESSB) r = slot(rs[0]) * sum;
            else if (kind == G3SPAWN_RICKER) r = slot(rs[0]) * sum * xexp(-slot(rs[1]) * sum);
            else if (kind == G3SPAWN_BEVERTONHOLT) r = slot(rs[0]) * sum / (slot(rs[1]) + sum);
            else if (kind == G3SPAWN_BEVERTONHOLT_SS3) {
                T h = slot(rs[0]), B0 = slot(rs[1]), Ry = slot(S.I[rs[2] + yidx]);
                r = 4.0 * h * sum * Ry / (B0 * (1.0 - h) + sum * (5.0 * h - 1.0));
            } else if (kind == G3SPAWN_HOCKEYSTICK) r = slot(rs[0]) * logspace_add(-1000.0 * sum / slot(rs[1]), T(-1000.0)) / -1000.0;
            else return T(NaN);
            // stock__offspringnum = r * 1 / avoid_zero(sum(1)) in every cell; the offspring step sums it again
            T per = r * 1.0 / avoid_zero(T((double)ncell));
            for (int i = 0; i < ncell; i++) spawn_total[s] = spawn_total[s] + per;
        }

        // ---- g3a_step_transition after growth (R/action_mature.R:78-134)
        for (int s = 0; s < NS; s++) {
            const int32_t* St = S.stock(s);
            if (St[G3S_GROW_N] < 0 || St[G3S_N_OUT] == 0) continue;
            if (!((St[G3S_TRANS_MASK] >> (cur_step - 1)) & 1)) continue;
            int ncell = St[G3S_L] * St[G3S_NAGE] * St[G3S_NAREA], c0 = St[G3S_CELL_OFF];
            std::vector<T> rem(trans_num[s]);
            for (int o = 0; o < St[G3S_N_OUT]; o++) {
                T ratio = slot(S.I[St[G3S_OUT_OFF] + 2 * o + 1]);
                const int32_t* map = S.I + St[G3S_OUT_MAP_OFF] + o * ncell;
                for (int i = 0; i < ncell; i++) {
                    int d = map[i]; if (d < 0) continue;
                    T moved = trans_num[s][i] * ratio;
                    wgt[d] = ratio_add_pop(wgt[d], num[d], trans_wgt[s][i], moved);
                    num[d] = num[d] + moved;
                    rem[i] = rem[i] - moved;
                }
            }
            for (int i = 0; i < ncell; i++) num[c0 + i] = num[c0 + i] + rem[i];
        }

        // ---- g3a_renewal (R/action_renewal.R:166-188)
        for (int s = 0; s < NS; s++) {
            const int32_t* St = S.stock(s);
            int L = St[G3S_L]; const double* midlen = S.R + St[G3S_MIDLEN_ROFF];
            for (int k = 0; k < St[G3S_N_RENEW]; k++) {
                const int32_t* Rn = S.I + St[G3S_RENEW_OFF] + k * G3R_LEN;
                if (Rn[G3R_RUN_STEP] != 0 && Rn[G3R_RUN_STEP] != cur_step) continue;
                for (int r = 0; r < St[G3S_NAREA]; r++) {
                    int fs = S.I[Rn[G3R_FACTOR_OFF] + yidx * St[G3S_NAREA] + r];
                    if (fs < 0) continue;
                    T factor = slot(fs), mean = slot(Rn[G3R_MEAN]), sd = avoid_zero(slot(Rn[G3R_SD]));
                    T wa = slot(Rn[G3R_WALPHA]), wb = slot(Rn[G3R_WBETA]);
                    std::vector<T> d(L); T tot(0.0);
                    for (int l = 0; l < L; l++) { d[l] = dnorm(midlen[l], mean, sd); tot = tot + d[l]; }
                    T den = avoid_zero(tot);
                    for (int l = 0; l < L; l++) {
                        int c = St[G3S_CELL_OFF] + (r * St[G3S_NAGE] + Rn[G3R_AGE_IDX]) * L + l;
                        T rn = d[l] / den * 10000.0 * factor, rw = wa * xpow(midlen[l], wb);
                        renewalnum[c] = rn;
                        wgt[c] = ratio_add_pop(wgt[c], num[c], rw, rn);
                        num[c] = num[c] + rn;
                    }
                }
            }
        }

        // ---- g3a_spawn, order 8 (R/action_spawn.R:183-218): the offspring arrive in the youngest age of the output stocks
        for (int s = 0; s < NS; s++) {
            if (!spawn_now[s]) continue;
            const int32_t* Sp = S.I + S.stock(s)[G3S_SPAWN_OFF];
            for (int o = 0; o < Sp[G3SP_N_OUT]; o++) {
                const int32_t* Or = S.I + Sp[G3SP_OUT_OFF] + 6 * o;
                const int32_t* So = S.stock(Or[0]);
                int L = So[G3S_L], A = So[G3S_NAGE], NR = So[G3S_NAREA];
                const double* midlen = S.R + So[G3S_MIDLEN_ROFF];
                T mean = slot(Or[1]), sd = slot(Or[2]), wa = slot(Or[3]), wb = slot(Or[4]), ratio = slot(Or[5]);
                std::vector<T> shape(L); T tot(0.0);
                for (int l = 0; l < L; l++) { T z = (T(midlen[l]) - mean) * (1.0 / sd); shape[l] = xexp(-(z * z) * 0.5); }
                for (int r = 0; r < NR; r++) for (int l = 0; l < L; l++) tot = tot + shape[l];
                T den = avoid_zero(tot);
                for (int r = 0; r < NR; r++) for (int l = 0; l < L; l++) {
                    int c = So[G3S_CELL_OFF] + (r * A + 0) * L + l;
                    T spawned = (shape[l] / den) * spawn_total[s] * ratio;
                    wgt[c] = ratio_add_pop(wgt[c], num[c], wa * xpow(T(midlen[l]), wb), spawned);
                    num[c] = num[c] + spawned;
                }
            }
        }

        // ---- g3a_report_detail (R/action_report.R:177-213): this step's detail_* columns
        if (rep) {
            for (int s = 0; s < NS; s++) {
                const int32_t* St = S.stock(s);
                int c0 = St[G3S_CELL_OFF], n = St[G3S_L] * St[G3S_NAGE] * St[G3S_NAREA];
                for (int i = 0; i < n; i++) rep->renewalnum[s][(size_t)cur_time * n + i] = val(renewalnum[c0 + i]);
            }
            for (int q = 0; q < NPP; q++) {
                const int32_t* St = S.stock(S.pp(q)[G3PP_PREY]);
                int c0 = St[G3S_CELL_OFF], n = St[G3S_L] * St[G3S_NAGE] * St[G3S_NAREA];
                for (int i = 0; i < n; i++) {
                    rep->dsuit[q][(size_t)cur_time * n + i] = val(suit[q][c0 + i]);
                    rep->dcons[q][(size_t)cur_time * n + i] = val(cons[q][c0 + i]);
                }
            }
        }

        // ---- g3l_bounds_penalty (R/likelihood_bounds.R:13-16)
        if (cur_time == 0) {
            double bw = S.R[S.h(G3H_BOUND_WS_ROFF)], bs = S.R[S.h(G3H_BOUND_WS_ROFF) + 1];
            for (int i = 0; i < S.h(G3H_NBOUND); i++) {
                T p = mkpar(pv, S.I[S.h(G3H_BOUND_OFF) + i]);
                double lo = S.R[S.h(G3H_BOUND_ROFF) + 2 * i], up = S.R[S.h(G3H_BOUND_ROFF) + 2 * i + 1];
                T a = logspace_add(bs * (p - up) / (up - lo), T(0.0)) + logspace_add(bs * (lo - p) / (up - lo), T(0.0));
                T pen = a * a;
                nll = nll + bw * pen;
                comp_tot[NNLL - 1] = comp_tot[NNLL - 1] + pen;
                nllsteps[(size_t)(NNLL - 1) * NT + 0] += val(pen);
            }
        }
        // ---- g3l_distribution (R/likelihood_distribution.R:275-394)
        std::vector<std::vector<T>> xb(NX);
        for (int c = 0; c < NCOMP; c++) {
            const int32_t* C = S.comp(c);
            if (!(pv.p[C[G3C_WEIGHT_PAR]] > 0)) continue;
            T weight = mkpar(pv, C[G3C_WEIGHT_PAR]);
            int ti = S.I[C[G3C_TIDX_OFF] + cur_time];
            if (ti < 0) continue;
            int x = C[G3C_XBUF];
            if (xb[x].empty()) {            // collect vector: R/likelihood_distribution.R:275-287
                const int32_t* X = S.xbuf(x);
                xb[x].assign(NC, T(0.0));
                if (X[G3X_KIND] == 0) {
                    for (int k = 0; k < X[G3X_N_PP]; k++) {
                        int q = S.I[X[G3X_PP_OFF] + k];
                        const int32_t* St = S.stock(S.pp(q)[G3PP_PREY]);
                        int c0 = St[G3S_CELL_OFF], n = St[G3S_L] * St[G3S_NAGE] * St[G3S_NAREA];
                        for (int i = c0; i < c0 + n; i++)
                            xb[x][i] = xb[x][i] + (X[G3X_UNIT] == 0 ? cons[q][i] / avoid_zero(wgt[i]) : cons[q][i]);
                    }
                } else {
                    for (int i = 0; i < NC; i++) xb[x][i] = X[G3X_UNIT] == 0 ? num[i] : num[i] * wgt[i];
                }
            }
            bool si = C[G3C_FUNC] == G3F_SI_LOG || C[G3C_FUNC] == G3F_SI_LINEAR;
            int nd = C[G3C_N_DEST], nt = C[G3C_N_TIMES];
            T* m = model[c].data() + ((si || rep) ? (size_t)ti * nd : 0);
            if (C[G3C_MODE] == 0) {
                const int32_t* ptr = S.I + C[G3C_CSR_PTR_OFF]; const int32_t* idx = S.I + C[G3C_CSR_IDX_OFF];
                const double* cf = S.R + C[G3C_CSR_COEF_ROFF];
                for (int e = 0; e < nd; e++) for (int j = ptr[e]; j < ptr[e + 1]; j++) m[e] = m[e] + cf[j] * xb[x][idx[j]];
            } else {
                const int32_t* cd = S.I + C[G3C_CELL_DEST_OFF]; const double* cf = S.R + C[G3C_CELL_COEF_ROFF];
                for (int i = 0; i < NC; i++) if (cd[i] >= 0) m[cd[i]] = m[cd[i]] + cf[i] * xb[x][i];
            }
            if (!S.I[C[G3C_CMP_OFF] + cur_time]) continue;
            const double* obs = S.R + C[G3C_OBS_ROFF] + (size_t)ti * nd;
            int nb = C[G3C_N_BINS], nsl = C[G3C_N_SLICES], nag = C[G3C_N_AREAG];
            T cnll(0.0);
            if (C[G3C_FUNC] == G3F_SUMOFSQUARES && S.R[C[G3C_PARAM_ROFF]] != 0.0) {
                // stratified (over includes 'length'): x / avoid_zero(rowSums(x))[length], R/likelihood_distribution.R:8-16
                for (int ag = 0; ag < nag; ag++)
                    for (int b = 0; b < nb; b++) {
                        T tm(0.0); double to = 0.0;
                        for (int sl = 0; sl < nsl; sl++) { tm = tm + m[(ag * nsl + sl) * nb + b]; to += obs[(ag * nsl + sl) * nb + b]; }
                        T atm = avoid_zero(tm); double ato = avoid_zero(to);
                        for (int sl = 0; sl < nsl; sl++) {
                            int e = (ag * nsl + sl) * nb + b;
                            T dlt = m[e] / atm - obs[e] / ato;
                            cnll = cnll + dlt * dlt;
                        }
                    }
            } else if (C[G3C_FUNC] == G3F_SUMOFSQUARES) {   // R/likelihood_distribution.R:3-39
                for (int ag = 0; ag < nag; ag++) {
                    T tm(0.0); double to = 0.0;
                    for (int i = 0; i < nsl * nb; i++) { tm = tm + m[ag * nsl * nb + i]; to += obs[ag * nsl * nb + i]; }
                    T atm = avoid_zero(tm); double ato = avoid_zero(to);
                    for (int sl = 0; sl < nsl; sl++) {
                        T v(0.0);
                        for (int b = 0; b < nb; b++) {
                            int e = (ag * nsl + sl) * nb + b;
                            T dlt = m[e] / atm - obs[e] / ato;
                            v = v + dlt * dlt;
                        }
                        cnll = cnll + v;
                    }
                }
            } else if (C[G3C_FUNC] == G3F_MULTINOMIAL) {    // R/likelihood_distribution.R:41-60
                double eps = S.R[C[G3C_PARAM_ROFF]];
                for (int ag = 0; ag < nag; ag++) for (int sl = 0; sl < nsl; sl++) {
                    const T* mm = m + (ag * nsl + sl) * nb; const double* oo = obs + (ag * nsl + sl) * nb;
                    T tm(0.0); double so = 0.0, slg = 0.0;
                    for (int b = 0; b < nb; b++) { tm = tm + mm[b]; so += oo[b]; slg += std::lgamma(1 + oo[b]); }
                    T atm = avoid_zero(tm);
                    T likely(0.0);
                    for (int b = 0; b < nb; b++) likely = likely + oo[b] * xlog(dif_pmax(mm[b] / atm, 1.0 / (nb * eps), 1e4));
                    cnll = cnll + 2.0 * (-likely + (slg - std::lgamma(1 + so)));
                }
            } else if (C[G3C_FUNC] == G3F_SUMSQLOGRATIOS) {  // R/likelihood_distribution.R:127-136
                double eps = S.R[C[G3C_PARAM_ROFF]];
                for (int e = 0; e < nd; e++) {
                    T dlt = std::log(obs[e] + eps) - xlog(m[e] + eps);
                    cnll = cnll + dlt * dlt;
                }
            } else if (ti == nt - 1) {                      // surveyindices: R/likelihood_distribution.R:82-125
                double fa = S.R[C[G3C_PARAM_ROFF]], fb = S.R[C[G3C_PARAM_ROFF] + 1];
                bool lg = C[G3C_FUNC] == G3F_SI_LOG;
                for (int e = 0; e < nd; e++) {
                    std::vector<T> N(nt), I(nt); T mN(0.0), mI(0.0);
                    for (int t = 0; t < nt; t++) {
                        T mv = model[c][(size_t)t * nd + e]; double ov = S.R[C[G3C_OBS_ROFF] + (size_t)t * nd + e];
                        N[t] = lg ? xlog(avoid_zero(mv)) : mv;
                        I[t] = lg ? T(std::log(avoid_zero(ov))) : T(ov);
                        mN = mN + N[t]; mI = mI + I[t];
                    }
                    mN = mN / double(nt); mI = mI / double(nt);
                    T beta, alpha;
                    if (std::isnan(fb)) {
                        T sxy(0.0), sxx(0.0);
                        for (int t = 0; t < nt; t++) { sxy = sxy + (I[t] - mI) * (N[t] - mN); sxx = sxx + (N[t] - mN) * (N[t] - mN); }
                        beta = sxy / avoid_zero(sxx);
                    } else beta = T(fb);
                    alpha = std::isnan(fa) ? mI - beta * mN : T(fa);
                    si_params[c][0] = alpha; si_params[c][1] = beta;
                    T v(0.0);
                    for (int t = 0; t < nt; t++) { T r = alpha + beta * N[t] - I[t]; v = v + r * r; }
                    cnll = cnll + v;
                }
            }
            nll = nll + weight * cnll;
            comp_tot[c] = comp_tot[c] + cnll;
            nllsteps[(size_t)c * NT + cur_time] += val(cnll);
            if (!si && !rep) std::fill(model[c].begin(), model[c].end(), T(0.0));
        }
        // ---- g3l_random_dnorm / g3l_random_walk (R/likelihood_random.R:14-109): n = sense * dnorm(param, mean, sigma, log)
        for (int r = 0; r < S.h(G3H_NRAND); r++) {
            const int32_t* Rn = S.I + S.h(G3H_RAND_OFF) + r * G3RN_LEN;
            const int run = S.I[Rn[G3RN_RUN_OFF] + cur_time];
            if (!run || (run == 2 && cur_time != total_steps)) continue;
            T x = slot(S.I[Rn[G3RN_PARAM_OFF] + cur_time]), mu = slot(S.I[Rn[G3RN_MEAN_OFF] + cur_time]);
            T sd = slot(S.I[Rn[G3RN_SIGMA_OFF] + cur_time]);
            T resid = (x - mu) / sd;
            T logans = T(-std::log(std::sqrt(2 * M_PI))) - xlog(sd) - 0.5 * resid * resid;      // TMB dnorm
            T n = Rn[G3RN_LOG] ? T(0.0) - logans : xexp(logans);
            nll = nll + slot(Rn[G3RN_WEIGHT]) * n;
            comp_tot[NCOMP + r] = comp_tot[NCOMP + r] + n;
            nllsteps[(size_t)(NCOMP + r) * NT + cur_time] += val(n);
        }
        // ---- g3l_understocking (R/likelihood_understocking.R:36-46)
        if (S.h(G3H_US_N) > 0) {
            T tot(0.0);
            for (int i = 0; i < S.h(G3H_US_N); i++) tot = tot + overcons[S.I[S.h(G3H_US_OFF) + i]];
            double power = S.R[S.h(G3H_US_ROFF)], w = S.R[S.h(G3H_US_ROFF) + 1];
            T u = (power == 2.0) ? tot * tot : xpow(tot, power);
            nll = nll + w * u;
            comp_tot[NNLL - 2] = comp_tot[NNLL - 2] + u;
            nllsteps[(size_t)(NNLL - 2) * NT + cur_time] += val(u);
        }

        // ---- g3a_age (R/action_age.R:51-85), then movement hand-off (R/action_mature.R:78-134, move_remainder = FALSE)
        if (cur_step_final) {
            for (int s = 0; s < NS; s++) {
                const int32_t* St = S.stock(s);
                if (!St[G3S_AGE_ENABLE]) continue;
                int L = St[G3S_L], A = St[G3S_NAGE], NR = St[G3S_NAREA], c0 = St[G3S_CELL_OFF];
                mv_num[s].assign((size_t)NR * L, T(0.0)); mv_wgt[s].assign((size_t)NR * L, T(0.0));
                if (A == 1) {
                    if (St[G3S_AGE_N_OUT] == 0) continue;
                    for (int r = 0; r < NR; r++) for (int l = 0; l < L; l++) {
                        int c = c0 + r * L + l;
                        mv_num[s][r * L + l] = num[c]; mv_wgt[s][r * L + l] = wgt[c]; num[c] = T(0.0);
                    }
                    continue;
                }
                for (int r = 0; r < NR; r++) for (int a = A - 1; a >= 0; a--) for (int l = 0; l < L; l++) {
                    int c = c0 + (r * A + a) * L + l, cp = c - L;
                    if (a == A - 1) {
                        if (St[G3S_AGE_N_OUT] > 0) {
                            mv_num[s][r * L + l] = num[c]; mv_wgt[s][r * L + l] = wgt[c];
                            num[c] = num[cp]; wgt[c] = wgt[cp];
                        } else {
                            wgt[c] = ratio_add_pop(wgt[c], num[c], wgt[cp], num[cp]);
                            num[c] = num[c] + num[cp];
                        }
                    } else if (a == 0) {
                        num[c] = T(0.0);
                    } else {
                        num[c] = num[cp]; wgt[c] = wgt[cp];
                    }
                }
            }
            for (int s = 0; s < NS; s++) {
                const int32_t* St = S.stock(s);
                if (!St[G3S_AGE_ENABLE] || St[G3S_AGE_N_OUT] == 0) continue;
                int L = St[G3S_L], NR = St[G3S_NAREA];
                for (int o = 0; o < St[G3S_AGE_N_OUT]; o++) {
                    T ratio = slot(S.I[St[G3S_AGE_OUT_OFF] + 2 * o + 1]);
                    const int32_t* map = S.I + St[G3S_AGE_MAP_OFF] + o * NR * L;
                    for (int i = 0; i < NR * L; i++) {
                        int d = map[i]; if (d < 0) continue;
                        T moved = mv_num[s][i] * ratio;
                        wgt[d] = ratio_add_pop(wgt[d], num[d], mv_wgt[s][i], moved);
                        num[d] = num[d] + moved;
                    }
                }
            }
        }
    }

    if (comp_out) for (int i = 0; i < NNLL; i++) comp_out[i] = val(comp_tot[i]);
    if (rep) {
        rep->nll[0] = val(nll);
        std::memcpy(rep->steps, nllsteps.data(), sizeof(double) * nllsteps.size());
        for (int c = 0; c < NCOMP; c++) {
            for (size_t i = 0; i < model[c].size(); i++) rep->model[c][i] = val(model[c][i]);
            const int32_t* C = S.comp(c);
            bool si = C[G3C_FUNC] == G3F_SI_LOG || C[G3C_FUNC] == G3F_SI_LINEAR;
            rep->params[c][0] = si ? val(si_params[c][0]) : NaN;
            rep->params[c][1] = si ? val(si_params[c][1]) : NaN;
        }
        for (int q = 0; q < NPP; q++) {         // suit_<prey>_<pred>__report (R/action_predate.R:292-299)
            const int32_t* Q = S.pp(q); const int32_t* P = S.pred(Q[G3PP_PRED]); const int32_t* St = S.stock(Q[G3PP_PREY]);
            const bool sp = P[G3P_KIND] == G3PRED_STOCK;
            const int32_t* Sp = sp ? S.stock(P[G3P_STOCK]) : nullptr;
            int L = St[G3S_L], PL = sp ? Sp[G3S_L] : 1;
            bool ok = true;
            for (int pl = 0; pl < PL; pl++) for (int l = 0; l < L; l++)
                rep->suit_report[q][pl * L + l] = val(suitability<T>(Q[G3PP_SUIT_KIND], S.I + Q[G3PP_SUIT_OFF] + 6 * (Q[G3PP_SUIT_TMAP] >= 0 ? S.I[Q[G3PP_SUIT_TMAP] + std::min(total_steps, S.h(G3H_NT) - 1)] : 0), slot,      // (the last step's set)
                    (S.R + St[G3S_MIDLEN_ROFF])[l],
                    sp && Q[G3PP_SUIT_KIND] != G3SUIT_ANDERSENFLEET ? (S.R + Sp[G3S_MIDLEN_ROFF])[pl] : (S.R + St[G3S_MIDLEN_ROFF])[L - 1], ok));
        }
    }
    return nll;
}

int64_t report_len(const Spec& S) {
    int64_t n = 1 + (int64_t)S.h(G3H_NNLL) * S.h(G3H_NT);
    for (int s = 0; s < S.h(G3H_NSTOCK); s++) {
        const int32_t* St = S.stock(s);
        n += 2LL * S.h(G3H_NT) * St[G3S_L] * St[G3S_NAGE] * St[G3S_NAREA];
    }
    for (int c = 0; c < S.h(G3H_NCOMP); c++) n += (int64_t)S.comp(c)[G3C_N_TIMES] * S.comp(c)[G3C_N_DEST] + 2;
    for (int s = 0; s < S.h(G3H_NSTOCK); s++) {
        const int32_t* St = S.stock(s);
        n += (int64_t)S.h(G3H_NT) * St[G3S_L] * St[G3S_NAGE] * St[G3S_NAREA];
    }
    for (int q = 0; q < S.h(G3H_NPP); q++) {
        const int32_t* Q = S.pp(q); const int32_t* P = S.pred(Q[G3PP_PRED]); const int32_t* St = S.stock(Q[G3PP_PREY]);
        int PL = P[G3P_KIND] == G3PRED_STOCK ? S.stock(P[G3P_STOCK])[G3S_L] : 1;
        n += (int64_t)PL * St[G3S_L] + 2LL * S.h(G3H_NT) * St[G3S_L] * St[G3S_NAGE] * St[G3S_NAREA];
    }
    return n;
}

constexpr int KT = 8;   // tangents per Dual pass

void eval_range(const Spec& S, const double* par, int64_t b0, int64_t b1, const int32_t* tidx, int K,
                double* nll, double* comp, double* grad) {
    int P = S.h(G3H_NPAR), NN = S.h(G3H_NNLL);
    for (int64_t b = b0; b < b1; b++) {
        const double* p = par + b * P;
        ParVec<double> pv{p, nullptr, 0};
        double v = run_model<double>(S, pv, comp ? comp + b * NN : nullptr, nullptr);
        if (nll) nll[b] = v;
        if (grad && K > 0) {
            for (int k0 = 0; k0 < K; k0 += KT) {
                ParVec<Dual<KT>> dv{p, tidx + k0, std::min(KT, K - k0)};
                Dual<KT> r = run_model<Dual<KT>>(S, dv, nullptr, nullptr);
                for (int k = 0; k < std::min(KT, K - k0); k++) grad[b * K + k0 + k] = r.d[k];
            }
        }
    }
}

}  // namespace

extern "C" {

int64_t g3oracle_report_len(const int32_t* ints, const double* reals) {
    Spec S{ints, reals};
    return report_len(S);
}

// Batched evaluation on `nthreads` host threads (one std::thread per contiguous block).
int g3oracle_eval(const int32_t* ints, const double* reals, const double* par, int64_t B,
                  const int32_t* tangent_idx, int32_t K, double* nll, double* nll_comp, double* grad,
                  int32_t nthreads) {
    Spec S{ints, reals};
    if (S.h(G3H_VERSION) != G3CUDA_SPEC_VERSION) return -1;
    if (nthreads < 1) nthreads = 1;
    if (nthreads == 1 || B < 2) { eval_range(S, par, 0, B, tangent_idx, K, nll, nll_comp, grad); return 0; }
    std::vector<std::thread> th;
    int64_t per = (B + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        int64_t b0 = t * per, b1 = std::min<int64_t>(B, b0 + per);
        if (b0 >= b1) break;
        th.emplace_back([=, &S] { eval_range(S, par, b0, b1, tangent_idx, K, nll, nll_comp, grad); });
    }
    for (auto& t : th) t.join();
    return 0;
}

// Single-vector report, same buffer layout as g3cuda_report() (include/g3cuda.h).
int g3oracle_report(const int32_t* ints, const double* reals, const double* par, double* out) {
    Spec S{ints, reals};
    if (S.h(G3H_VERSION) != G3CUDA_SPEC_VERSION) return -1;
    int NT = S.h(G3H_NT);
    ReportSink rep;
    double* p = out;
    rep.nll = p; p += 1;
    rep.steps = p; p += (size_t)S.h(G3H_NNLL) * NT;
    for (int s = 0; s < S.h(G3H_NSTOCK); s++) {
        const int32_t* St = S.stock(s);
        size_t n = (size_t)NT * St[G3S_L] * St[G3S_NAGE] * St[G3S_NAREA];
        rep.dstart_num.push_back(p); p += n;
        rep.dstart_wgt.push_back(p); p += n;
    }
    for (int c = 0; c < S.h(G3H_NCOMP); c++) { rep.model.push_back(p); p += (size_t)S.comp(c)[G3C_N_TIMES] * S.comp(c)[G3C_N_DEST]; }
    for (int c = 0; c < S.h(G3H_NCOMP); c++) { rep.params.push_back(p); p += 2; }
    for (int s = 0; s < S.h(G3H_NSTOCK); s++) {
        const int32_t* St = S.stock(s);
        rep.renewalnum.push_back(p); p += (size_t)NT * St[G3S_L] * St[G3S_NAGE] * St[G3S_NAREA];
    }
    for (int q = 0; q < S.h(G3H_NPP); q++) {
        const int32_t* Q = S.pp(q); const int32_t* P = S.pred(Q[G3PP_PRED]); const int32_t* St = S.stock(Q[G3PP_PREY]);
        size_t PL = P[G3P_KIND] == G3PRED_STOCK ? S.stock(P[G3P_STOCK])[G3S_L] : 1;
        size_t n = (size_t)NT * St[G3S_L] * St[G3S_NAGE] * St[G3S_NAREA];
        rep.suit_report.push_back(p); p += PL * St[G3S_L];
        rep.dsuit.push_back(p); p += n;
        rep.dcons.push_back(p); p += n;
    }
    std::fill(out, p, 0.0);
    ParVec<double> pv{par, nullptr, 0};
    run_model<double>(S, pv, nullptr, &rep);
    return 0;
}

// ---- primitives exposed for the known-answer tests (tests/test_oracle_kat.py) ----
void g3oracle_growth_bbinom(const double* delt_l, int32_t na, int32_t binn, double beta, double* out) {
    std::vector<double> d(delt_l, delt_l + na), o;
    growth_bbinom<double>(d, binn, beta, o);
    std::copy(o.begin(), o.end(), out);
}
void g3oracle_grow_matrix_len(const double* delta, int32_t L, int32_t D, double* out) {
    std::vector<double> d(delta, delta + (size_t)L * D), o;
    grow_matrix_len<double>(d, L, D, o);
    std::copy(o.begin(), o.end(), out);
}
void g3oracle_grow_matrix_wgt(const double* delta, int32_t L, int32_t D, double* out) {
    std::vector<double> d(delta, delta + (size_t)L * D), o;
    grow_matrix_wgt<double>(d, L, D, o);
    std::copy(o.begin(), o.end(), out);
}
void g3oracle_grow_apply(const double* G, const double* W, int32_t L, const double* num, const double* wgt,
                         double* out_num, double* out_wgt) {
    std::vector<double> g(G, G + (size_t)L * L), w(W, W + (size_t)L * L), on, ow;
    grow_apply<double>(g, w, L, num, wgt, on, ow);
    std::copy(on.begin(), on.end(), out_num);
    std::copy(ow.begin(), ow.end(), out_wgt);
}
double g3oracle_grow_lengthvbsimple(double linf, double kappa, double midlen, double dt) {
    return grow_lengthvbsimple<double>(linf, kappa, midlen, dt);
}
void g3oracle_set_bbinom_product(int32_t on) { g_bbinom_product = on; }
double g3oracle_suitability(int32_t kind, const double* par5, double midlen, double pred_len) {
    int32_t ss[5] = {0, 1, 2, 3, 4};
    auto slot = [&](int i) { return par5[i]; };
    bool ok = true;
    double v = suitability<double>(kind, ss, slot, midlen, pred_len, ok);
    return ok ? v : std::numeric_limits<double>::quiet_NaN();
}
double g3oracle_suit_andersen(double p0, double p1, double p2, double p3, double p4, double pred_len, double midlen) {
    double par[5] = {p0, p1, p2, p3, p4};
    int32_t ss[5] = {0, 1, 2, 3, 4};
    auto slot = [&](int i) { return par[i]; };
    bool ok = true;
    return suitability<double>(G3SUIT_ANDERSEN, ss, slot, midlen, pred_len, ok);
}
void g3oracle_migrate_normalize(double* vec, int32_t n, int32_t src_idx, double row_total) {
    std::vector<double> v(vec, vec + n);
    migrate_normalize<double>(v, src_idx, row_total);
    std::copy(v.begin(), v.end(), vec);
}
double g3oracle_avoid_zero(double a) { return avoid_zero<double>(a); }
double g3oracle_logspace_add(double a, double b) { return logspace_add<double>(a, b); }
double g3oracle_dif_pmin(double a, double b, double s) { return dif_pmin<double>(a, b, s); }
double g3oracle_dif_pminmax(double a, double lo, double hi, double s) { return dif_pminmax<double>(a, lo, hi, s); }
double g3oracle_dnorm(double x, double m, double s) { return dnorm<double>(x, m, s); }
double g3oracle_ratio_add_pop(double ov, double oa, double nv, double na) { return ratio_add_pop<double>(ov, oa, nv, na); }
double g3oracle_digamma(double x) { return digamma(x); }

}  // extern "C"
