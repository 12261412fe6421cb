/*
 * TMB.hpp -- a minimal stand-in for the TMB / Eigen / R headers, just large enough to compile the
 * reference's checked-in GENERATED objectives (baseline/introduction-single-stock.cpp,
 * baseline/multiple-substocks.cpp) unmodified, where they lie, for Type = double.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/g3oracle.cpp). It exists so that the CPU oracle can be
 * validated against the reference's own code path: the generated translation unit is the
 * reference; this file only supplies the container types (eagerly evaluated, column-major, with
 * tmbutils::array's col()/view semantics), the handful of R-math functions it calls
 * (logspace_add, dnorm, lgamma as documented for TMB/Rmath) and the DATA_/PARAMETER/REPORT macros.
 * It is NOT TMB: no AD, no Eigen expression templates, no R.
 */
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <map>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <vector>

#define TMBAD_FRAMEWORK
namespace TMBad { namespace global { struct ad_aug {}; } }
namespace CppAD { template <class T> T abs(T x) { return std::fabs(x); } }
#ifndef TRUE
#define TRUE true
#define FALSE false
#endif
inline int& g3shim_warnings() { static int n = 0; return n; }
inline void Rf_warning(const char*, ...) { g3shim_warnings()++; }
inline double asDouble(double x) { return x; }
using std::exp; using std::floor; using std::log; using std::pow; using std::sqrt;

// TMB's logspace_add (R's logspace_add): max + log1p(exp(-|dx|))
inline double logspace_add(double lx, double ly) {
    if (lx == -INFINITY) return ly;
    if (ly == -INFINITY) return lx;
    return lx < ly ? ly + std::log1p(std::exp(lx - ly)) : lx + std::log1p(std::exp(ly - lx));
}
// TMB's dnorm (tmb_core density): exp(-log(sqrt(2 pi)) - log(sd) - .5 resid^2)
inline double dnorm(double x, double mean, double sd, int give_log = 0) {
    double resid = (x - mean) / sd;
    double logans = -std::log(std::sqrt(2 * M_PI)) - std::log(sd) - 0.5 * resid * resid;
    return give_log ? logans : std::exp(logans);
}

namespace Eigen {
template <class Derived> struct DenseBase {
    const Derived& derived() const { return *static_cast<const Derived*>(this); }
    long size() const { return derived().size(); }
    auto operator[](long i) const { return derived()[i]; }
    auto operator()(long i) const { return derived()[i]; }
    auto replicate(int a, int b) const { return derived().replicate(a, b); }
};
template <class X> class Map;       // only named in never-instantiated overloads
}

template <class T> struct Mat;
template <class T> struct MArr;

template <class T> struct Flat : Eigen::DenseBase<Flat<T>> {
    typedef T value_type;
    std::shared_ptr<std::vector<T>> buf;
    T* p = nullptr;
    std::vector<int> dim;
    long n = 0;
    bool view = false;

    Flat() {}
    void alloc(std::vector<int> d) {
        dim = d; n = 1; for (int x : d) n *= x;
        buf = std::make_shared<std::vector<T>>((size_t)std::max<long>(n, 1)); p = buf->data(); view = false;
    }
    explicit Flat(int a) { alloc({a}); }
    Flat(int a, int b) { alloc({a, b}); }
    Flat(int a, int b, int c) { alloc({a, b, c}); }
    Flat(int a, int b, int c, int d) { alloc({a, b, c, d}); }
    Flat(int a, int b, int c, int d, int e) { alloc({a, b, c, d, e}); }
    Flat(T* p_, std::vector<int> d, std::shared_ptr<std::vector<T>> keep) : buf(keep), p(p_), dim(d), view(true) { n = 1; for (int x : d) n *= x; }
    Flat(const Flat& o) { alloc(o.dim); std::copy(o.p, o.p + n, p); }            // copies are deep (tmbutils::array)
    Flat(Flat&& o) noexcept : buf(std::move(o.buf)), p(o.p), dim(std::move(o.dim)), n(o.n), view(o.view) {}
    template <class U, class = std::enable_if_t<!std::is_same<U, T>::value>> Flat(const Flat<U>& o) { alloc(o.dim); for (long i = 0; i < n; i++) p[i] = (T)o.p[i]; }
    Flat(const Mat<T>& m);
    Flat(const MArr<T>& m);
    void assign(const Flat& o) {
        if (p && n == o.n) { if (p != o.p) std::copy(o.p, o.p + n, p); return; }
        if (view) throw std::runtime_error("g3shim: size mismatch assigning into an array view");
        alloc(o.dim); std::copy(o.p, o.p + n, p);
    }
    Flat& operator=(const Flat& o) { assign(o); return *this; }
    Flat& operator=(Flat&& o) {
        if (!p) { buf = std::move(o.buf); p = o.p; dim = std::move(o.dim); n = o.n; view = o.view; if (view) { Flat t(*this); *this = Flat(); alloc(t.dim); std::copy(t.p, t.p + n, p); } return *this; }
        assign(o); return *this;
    }
    Flat& operator=(const MArr<T>& m);
    Flat& operator=(const Mat<T>& m);
    template <class S, class = std::enable_if_t<std::is_arithmetic<S>::value>> Flat& operator=(S s) { for (long i = 0; i < n; i++) p[i] = (T)s; return *this; }

    long size() const { return n; }
    int rows() const { return dim.empty() ? 0 : dim[0]; }
    int cols() const { return dim.empty() ? 0 : dim.back(); }
    T& operator[](long i) { return p[i]; }
    const T& operator[](long i) const { return p[i]; }
    T& operator()(long i) { return p[i]; }
    const T& operator()(long i) const { return p[i]; }
    T& operator()(long i, long j) { return p[i + (long)dim[0] * j]; }
    const T& operator()(long i, long j) const { return p[i + (long)dim[0] * j]; }
    T& operator()(long i, long j, long k) { return p[i + (long)dim[0] * (j + (long)dim[1] * k)]; }
    T& operator()(long i, long j, long k, long l) { return p[i + (long)dim[0] * (j + (long)dim[1] * (k + (long)dim[2] * l))]; }
    Flat col(long i) const {            // tmbutils::array::col: index of the LAST dimension, returns a view
        std::vector<int> nd(dim.begin(), dim.end() - 1);
        if (nd.empty()) nd.push_back(1);
        long stride = n / dim.back();
        return Flat(p + i * stride, nd, buf);
    }
    void setZero() { for (long i = 0; i < n; i++) p[i] = T(0); }
    void setConstant(T v) { for (long i = 0; i < n; i++) p[i] = v; }
    T sum() const { T s = T(0); for (long i = 0; i < n; i++) s += p[i]; return s; }
    T mean() const { return sum() / (T)n; }
    Flat vec() const { Flat r((int)n); std::copy(p, p + n, r.p); return r; }
    Flat array() const { return Flat(*this); }
    Flat transpose() const { return Flat(*this); }
    Flat pow(T e) const { Flat r(*this); for (long i = 0; i < n; i++) r.p[i] = std::pow(p[i], e); return r; }
    Flat exp() const { Flat r(*this); for (long i = 0; i < n; i++) r.p[i] = std::exp(p[i]); return r; }
    Flat replicate(int a, int b) const {
        assert(b == 1); (void)b;
        Flat r((int)(n * a));
        for (int k = 0; k < a; k++) std::copy(p, p + n, r.p + k * n);
        return r;
    }
    Mat<T> matrix() const;
    struct Rowwise { const Flat& f; Flat operator/(const Flat& d) const { Flat r(f); for (long i = 0; i < r.n; i++) r.p[i] = f.p[i] / d.p[i % d.n]; return r; } };
    Rowwise rowwise() const { return Rowwise{*this}; }

#define G3SHIM_CASSIGN(OP)                                                                              \
    Flat& operator OP(const Flat& o) { for (long i = 0; i < n; i++) p[i] OP o.p[o.n == n ? i : i % o.n]; return *this; } \
    template <class S, class = std::enable_if_t<std::is_arithmetic<S>::value>> Flat& operator OP(S s) { for (long i = 0; i < n; i++) p[i] OP (T)s; return *this; }
    G3SHIM_CASSIGN(+=) G3SHIM_CASSIGN(-=) G3SHIM_CASSIGN(*=) G3SHIM_CASSIGN(/=)
#undef G3SHIM_CASSIGN
};

template <class T> using array = Flat<T>;
template <class T> using vector = Flat<T>;

#define G3SHIM_BINOP(OP)                                                                                 \
    template <class T> Flat<T> operator OP(const Flat<T>& a, const Flat<T>& b) {                          \
        if (a.n != b.n) throw std::runtime_error("g3shim: size mismatch in elementwise " #OP);            \
        Flat<T> r(a); for (long i = 0; i < r.n; i++) r.p[i] = a.p[i] OP b.p[i]; return r; }               \
    template <class T, class S, class = std::enable_if_t<std::is_arithmetic<S>::value>>                   \
    Flat<T> operator OP(const Flat<T>& a, S s) { Flat<T> r(a); for (long i = 0; i < r.n; i++) r.p[i] = a.p[i] OP (T)s; return r; } \
    template <class T, class S, class = std::enable_if_t<std::is_arithmetic<S>::value>>                   \
    Flat<T> operator OP(S s, const Flat<T>& a) { Flat<T> r(a); for (long i = 0; i < r.n; i++) r.p[i] = (T)s OP a.p[i]; return r; }
G3SHIM_BINOP(+) G3SHIM_BINOP(-) G3SHIM_BINOP(*) G3SHIM_BINOP(/)
#undef G3SHIM_BINOP
template <class T> Flat<T> operator-(const Flat<T>& a) { Flat<T> r(a); for (long i = 0; i < r.n; i++) r.p[i] = -a.p[i]; return r; }

#define G3SHIM_UNARY(NAME, EXPR) \
    template <class T> Flat<T> NAME(const Flat<T>& a) { Flat<T> r(a); for (long i = 0; i < r.n; i++) { T x = a.p[i]; r.p[i] = EXPR; } return r; }
G3SHIM_UNARY(exp, std::exp(x)) G3SHIM_UNARY(log, std::log(x)) G3SHIM_UNARY(lgamma, ::lgamma(x)) G3SHIM_UNARY(sqrt, std::sqrt(x))
#undef G3SHIM_UNARY
template <class T, class S> Flat<T> pow(const Flat<T>& a, S e) { return a.pow((T)e); }
template <class T> Flat<T> dnorm(const Flat<T>& x, T mean, T sd, int give_log = 0) {
    Flat<T> r(x); for (long i = 0; i < r.n; i++) r.p[i] = dnorm(x.p[i], mean, sd, give_log); return r;
}

// ------------------------------------------------------------------ 2-D types
template <class T> struct MArr {        // elementwise view of a matrix (".array()")
    int r = 0, c = 0; std::vector<T> d;
    struct Colwise {
        const MArr& m;
        MArr operator*(const Flat<T>& v) const { MArr o(m); for (int j = 0; j < m.c; j++) for (int i = 0; i < m.r; i++) o.d[i + m.r * j] = m.d[i + m.r * j] * v.p[i]; return o; }
        MArr operator+(const Flat<T>& v) const { MArr o(m); for (int j = 0; j < m.c; j++) for (int i = 0; i < m.r; i++) o.d[i + m.r * j] = m.d[i + m.r * j] + v.p[i]; return o; }
    };
    Colwise colwise() const { return Colwise{*this}; }
    MArr operator*(const MArr& o) const { MArr x(*this); for (size_t i = 0; i < d.size(); i++) x.d[i] = d[i] * o.d[i]; return x; }
    MArr operator+(const MArr& o) const { MArr x(*this); for (size_t i = 0; i < d.size(); i++) x.d[i] = d[i] + o.d[i]; return x; }
};

template <class T> struct Mat {
    int r = 0, c = 0; std::vector<T> d;
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_), d((size_t)r_ * c_) {}
    Mat(const MArr<T>& a) : r(a.r), c(a.c), d(a.d) {}
    Mat& operator=(const MArr<T>& a) { r = a.r; c = a.c; d = a.d; return *this; }
    int rows() const { return r; }
    int cols() const { return c; }
    void setZero() { std::fill(d.begin(), d.end(), T(0)); }
    T& operator()(int i, int j) { return d[i + (size_t)r * j]; }
    const T& operator()(int i, int j) const { return d[i + (size_t)r * j]; }
    struct Block {
        Mat* m; const Mat* cm; int i, j, p, q;
        T get(int a, int b) const { return (*cm)(i + a, j + b); }
        Block& operator=(const Block& o) { for (int b = 0; b < q; b++) for (int a = 0; a < p; a++) (*m)(i + a, j + b) = o.get(a, b); return *this; }
        T sum() const { T s = T(0); for (int b = 0; b < q; b++) for (int a = 0; a < p; a++) s += get(a, b); return s; }
        Block tail(int k) const { return Block{m, cm, i, j + q - k, p, k}; }      // (row vectors only)
    };
    Block block(int i, int j, int p, int q) { return Block{this, this, i, j, p, q}; }
    Block block(int i, int j, int p, int q) const { return Block{nullptr, this, i, j, p, q}; }
    Block row(int i) { return Block{this, this, i, 0, 1, c}; }
    Block row(int i) const { return Block{nullptr, this, i, 0, 1, c}; }
    struct Colwise { const Mat& m; Flat<T> sum() const { Flat<T> o(m.c); for (int j = 0; j < m.c; j++) { T s = T(0); for (int i = 0; i < m.r; i++) s += m(i, j); o.p[j] = s; } return o; } };
    Colwise colwise() const { return Colwise{*this}; }
    MArr<T> array() const { MArr<T> a; a.r = r; a.c = c; a.d = d; return a; }
    Mat transpose() const { Mat t(c, r); for (int i = 0; i < r; i++) for (int j = 0; j < c; j++) t(j, i) = (*this)(i, j); return t; }
    Mat operator*(const Mat& o) const {
        if (c != o.r) throw std::runtime_error("g3shim: matrix product size mismatch");
        Mat x(r, o.c);
        for (int j = 0; j < o.c; j++) for (int i = 0; i < r; i++) { T s = T(0); for (int k = 0; k < c; k++) s += (*this)(i, k) * o(k, j); x(i, j) = s; }
        return x;
    }
    Flat<T> vec() const { Flat<T> o((int)d.size()); std::copy(d.begin(), d.end(), o.p); return o; }
};
template <class T> using matrix = Mat<T>;
template <class T> Flat<T>::Flat(const Mat<T>& m) { alloc({m.r, m.c}); std::copy(m.d.begin(), m.d.end(), p); }
template <class T> Flat<T>::Flat(const MArr<T>& m) { alloc({m.r, m.c}); std::copy(m.d.begin(), m.d.end(), p); }
template <class T> Flat<T>& Flat<T>::operator=(const MArr<T>& m) { Flat t(m); assign(t); return *this; }
template <class T> Flat<T>& Flat<T>::operator=(const Mat<T>& m) { Flat t(m); assign(t); return *this; }
template <class T> Mat<T> Flat<T>::matrix() const {
    int r = dim.size() >= 2 ? dim[0] : (int)n, c = dim.size() >= 2 ? (int)(n / dim[0]) : 1;
    Mat<T> m(r, c); std::copy(p, p + n, m.d.begin()); return m;
}

// ------------------------------------------------------------------ objective_function + macros
struct G3ShimStore {
    std::map<std::string, double> scalars, params;
    std::map<std::string, std::pair<std::vector<double>, std::vector<int>>> arrays;
    std::map<std::string, std::pair<std::vector<double>, std::vector<int>>> reports;
    std::vector<std::string> missing;
};
template <class Type> struct objective_function {
    G3ShimStore* S;
    explicit objective_function(G3ShimStore* s) : S(s) {}
    Type operator()();
    double get_scalar(const char* n) { auto it = S->scalars.find(n); if (it == S->scalars.end()) { S->missing.push_back(n); return NAN; } return it->second; }
    double get_param(const char* n) { auto it = S->params.find(n); if (it == S->params.end()) { S->missing.push_back(n); return NAN; } return it->second; }
    template <class T> Flat<T> get_array(const char* n) {
        auto it = S->arrays.find(n);
        if (it == S->arrays.end()) { S->missing.push_back(n); return Flat<T>(0); }
        Flat<T> a; a.alloc(it->second.second);
        for (long i = 0; i < a.n; i++) a.p[i] = (T)it->second.first[i];
        return a;
    }
    void put(const char* n, double v) { S->reports[n] = {{v}, {1}}; }
    template <class T> void put(const char* n, const Flat<T>& a) { std::vector<double> v(a.n); for (long i = 0; i < a.n; i++) v[i] = (double)a.p[i]; S->reports[n] = {v, a.dim}; }
};
#define DATA_SCALAR(x) Type x = this->get_scalar(#x);
#define DATA_UPDATE(x)
#define DATA_INTEGER(x) int x = (int)this->get_scalar(#x);
#define PARAMETER(x) Type x = this->get_param(#x);
#define PARAMETER_VECTOR(x) vector<Type> x = this->template get_array<Type>(#x);
#define DATA_VECTOR(x) vector<Type> x = this->template get_array<Type>(#x);
#define DATA_ARRAY(x) array<Type> x = this->template get_array<Type>(#x);
#define DATA_IVECTOR(x) vector<int> x = this->template get_array<int>(#x);
#define REPORT(x) this->put(#x, x);
#define ADREPORT(x)

