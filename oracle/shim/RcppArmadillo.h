// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
//
// Stand-in for <RcppArmadillo.h> so that the reference engine headers
// (/root/reference/inst/include/*.h) and model sources (/root/reference/src/*.cpp)
// compile UNMODIFIED in a container that has no R, Rcpp, Armadillo or Rmath.
// Only the subset of the Armadillo / Rcpp / R-nmath surface that those files
// touch is provided.  Nothing under rcppsmc_b200/ may
// include this file.
//
// Third-party arithmetic restated here is PARITY-UNPINNED (the real libraries
// are not vendored in the reference and no reference test pins them):
//   * arma::sum  -> two interleaved accumulators (Armadillo accu() order, recollection)
//   * arma::cumsum -> sequential left-to-right fp64
//   * R::runif(a,b) = a + (b-a)*unif_rand();  R::rnorm(m,s) = m + s*norm_rand()
//   * rmultinom  -> sequential conditional binomials (R nmath control flow)
// Randomness is injectable: queues of uniforms / standard normals / multinomial
// count vectors are consumed first; when a queue is empty a std::mt19937_64
// stream is used (NOT R's Mersenne-Twister + inversion).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <deque>
#include <initializer_list>
#include <iostream>
#include <numeric>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif

// ---------------------------------------------------------------------------
// shim state: RNG + injection queues (one instance per process)
// ---------------------------------------------------------------------------
namespace shim {
struct State {
    std::mt19937_64 eng{20251019ull};
    std::normal_distribution<double> nd{0.0, 1.0};
    std::deque<double> uq;                 // uniforms in [0,1)
    std::deque<double> zq;                 // standard normals
    std::deque<std::vector<int>> mq;       // multinomial count vectors
    const double* exp_override = nullptr;  // one-shot replacement for an upcoming arma::exp(vec) of matching length
    long exp_override_n = 0;
    int exp_override_skip = 0;             // matching calls to let through first (e.g. the one inside stableLogSumWeights)
    long n_unif = 0, n_norm = 0, n_multinom = 0;  // consumption counters
};
inline State& st() { static State s; return s; }
inline double unif01() {
    State& s = st(); ++s.n_unif;
    if (!s.uq.empty()) { double u = s.uq.front(); s.uq.pop_front(); return u; }
    return std::generate_canonical<double, 53>(s.eng);
}
inline double norm01() {
    State& s = st(); ++s.n_norm;
    if (!s.zq.empty()) { double z = s.zq.front(); s.zq.pop_front(); return z; }
    return s.nd(s.eng);
}
}  // namespace shim

// ---------------------------------------------------------------------------
// Armadillo subset
// ---------------------------------------------------------------------------
namespace arma {
typedef unsigned long long uword;
namespace fill { struct fill_zeros {}; static const fill_zeros zeros{}; }

template <class T> struct Mat;

template <class T> struct Col {
    std::vector<T> d;
    uword n_rows = 0, n_elem = 0;
    static const uword n_cols = 1;
    void fix() { n_rows = n_elem = d.size(); }
    Col() {}
    explicit Col(uword n) : d(n) { fix(); }
    Col(uword n, fill::fill_zeros) : d(n, T(0)) { fix(); }
    Col(std::initializer_list<T> l) : d(l) { fix(); }
    Col(const std::vector<T>& v) : d(v) { fix(); }
    Col(const Col& o) : d(o.d) { fix(); }
    Col& operator=(const Col& o) { d = o.d; fix(); return *this; }
    T& operator()(uword i) { return d[i]; }
    const T& operator()(uword i) const { return d[i]; }
    T& operator[](uword i) { return d[i]; }
    const T& operator[](uword i) const { return d[i]; }
    T& at(uword i) { return d[i]; }
    const T& at(uword i) const { return d[i]; }
    T* memptr() { return d.data(); }
    const T* memptr() const { return d.data(); }
    uword size() const { return d.size(); }
    bool is_empty() const { return d.empty(); }
    void reset() { d.clear(); fix(); }
    void fill(T v) { std::fill(d.begin(), d.end(), v); }
    void zeros() { fill(T(0)); }
    void zeros(uword n) { d.assign(n, T(0)); fix(); }
    void set_size(uword n) { d.resize(n); fix(); }
    void resize(uword n) { d.resize(n); fix(); }
    T* begin() { return d.data(); }
    T* end() { return d.data() + d.size(); }
    const T* begin() const { return d.data(); }
    const T* end() const { return d.data() + d.size(); }
    Col head(uword n) const { Col r(n); std::copy(d.begin(), d.begin() + n, r.d.begin()); return r; }
    Col tail(uword n) const { Col r(n); std::copy(d.end() - n, d.end(), r.d.begin()); return r; }
    Col rows(uword a, uword b) const { return subvec(a, b); }
    Col subvec(uword a, uword b) const { Col r(b - a + 1); std::copy(d.begin() + a, d.begin() + b + 1, r.d.begin()); return r; }
    struct ElemRef {   // x.elem(indices).fill(v)
        Col& c; std::vector<uword> ix;
        void fill(T v) { for (uword i : ix) c.d[i] = v; }
    };
    template <class U> ElemRef elem(const Col<U>& ix) { ElemRef r{*this, {}}; for (auto v : ix.d) r.ix.push_back((uword)v); return r; }
    template <class U> Col elem(const Col<U>& ix) const {
        Col r(ix.d.size());
        for (size_t i = 0; i < ix.d.size(); ++i) r.d[i] = d[ix.d[i]];
        return r;
    }
    template <class U> Col& operator+=(const Col<U>& o) { for (size_t i = 0; i < d.size(); ++i) d[i] += o.d[i]; return *this; }
    template <class U> Col& operator-=(const Col<U>& o) { for (size_t i = 0; i < d.size(); ++i) d[i] -= o.d[i]; return *this; }
    Col& operator+=(T b) { for (auto& v : d) v += b; return *this; }
    Col& operator-=(T b) { for (auto& v : d) v -= b; return *this; }
    Col& operator*=(T b) { for (auto& v : d) v *= b; return *this; }
    Col& operator/=(T b) { for (auto& v : d) v /= b; return *this; }
    Mat<T> t() const;
};
typedef Col<double> vec;
typedef Col<double> colvec;
typedef Col<uword> uvec;
typedef Col<int> ivec;

// element-wise vec arithmetic (left-to-right, one rounding per op like Armadillo's eOp loops)
#define SHIM_VV(op)                                                                         \
    template <class U> inline vec operator op(const vec& a, const Col<U>& b) {              \
        vec r(a.d.size());                                                                   \
        for (size_t i = 0; i < a.d.size(); ++i) r.d[i] = a.d[i] op (double)b.d[i];           \
        return r;                                                                            \
    }
SHIM_VV(+) SHIM_VV(-) SHIM_VV(/)
#undef SHIM_VV
inline vec operator%(const vec& a, const vec& b) { vec r(a.d.size()); for (size_t i = 0; i < a.d.size(); ++i) r.d[i] = a.d[i] * b.d[i]; return r; }
#define SHIM_VS(op)                                                                         \
    inline vec operator op(const vec& a, double b) { vec r(a.d.size()); for (size_t i = 0; i < a.d.size(); ++i) r.d[i] = a.d[i] op b; return r; } \
    inline vec operator op(double b, const vec& a) { vec r(a.d.size()); for (size_t i = 0; i < a.d.size(); ++i) r.d[i] = b op a.d[i]; return r; }
SHIM_VS(+) SHIM_VS(-) SHIM_VS(*) SHIM_VS(/)
#undef SHIM_VS
inline vec operator-(const vec& a) { vec r(a.d.size()); for (size_t i = 0; i < a.d.size(); ++i) r.d[i] = -a.d[i]; return r; }

inline vec exp(const vec& a) {
    shim::State& s = shim::st();
    if (s.exp_override && (long)a.d.size() == s.exp_override_n && s.exp_override_skip-- <= 0) {
        vec r(a.d.size());
        std::copy(s.exp_override, s.exp_override + s.exp_override_n, r.d.begin());
        s.exp_override = nullptr; s.exp_override_n = 0;
        return r;
    }
    vec r(a.d.size());
    for (size_t i = 0; i < a.d.size(); ++i) r.d[i] = std::exp(a.d[i]);
    return r;
}
inline vec log(const vec& a) { vec r(a.d.size()); for (size_t i = 0; i < a.d.size(); ++i) r.d[i] = std::log(a.d[i]); return r; }
inline vec sqrt(const vec& a) { vec r(a.d.size()); for (size_t i = 0; i < a.d.size(); ++i) r.d[i] = std::sqrt(a.d[i]); return r; }
inline vec pow(const vec& a, double p) { vec r(a.d.size()); for (size_t i = 0; i < a.d.size(); ++i) r.d[i] = std::pow(a.d[i], p); return r; }
inline vec square(const vec& a) { vec r(a.d.size()); for (size_t i = 0; i < a.d.size(); ++i) r.d[i] = a.d[i] * a.d[i]; return r; }
template <class T> inline T max(const Col<T>& a) { return *std::max_element(a.d.begin(), a.d.end()); }
template <class T> inline T min(const Col<T>& a) { return *std::min_element(a.d.begin(), a.d.end()); }
// Armadillo accu(): two accumulators over even/odd positions, added at the end (recollection)
template <class T> inline T sum(const Col<T>& a) {
    T v1 = 0, v2 = 0; size_t i, j, n = a.d.size();
    for (i = 0, j = 1; j < n; i += 2, j += 2) { v1 += a.d[i]; v2 += a.d[j]; }
    if (i < n) v1 += a.d[i];
    return v1 + v2;
}
template <class T> inline T accu(const Col<T>& a) { return sum(a); }
template <class T> inline Col<T> cumsum(const Col<T>& a) {
    Col<T> r(a.d.size()); T acc = 0;
    for (size_t i = 0; i < a.d.size(); ++i) { acc += a.d[i]; r.d[i] = acc; }
    return r;
}
template <class T> inline double mean(const Col<T>& a) { return (double)sum(a) / (double)a.d.size(); }
template <class T> inline T as_scalar(const Col<T>& a) { return a.d.at(0); }
inline double as_scalar(double v) { return v; }
template <class C> inline C zeros(uword n) { C r(n); std::fill(r.d.begin(), r.d.end(), 0); return r; }
inline vec zeros(uword n) { return zeros<vec>(n); }
template <class C> inline C ones(uword n) { C r(n); std::fill(r.d.begin(), r.d.end(), 1); return r; }
inline vec ones(uword n) { return ones<vec>(n); }
template <class C> inline C linspace(double a, double b, uword n) {
    C r(n);
    for (uword i = 0; i < n; ++i) r.d[i] = (n > 1) ? (a + (b - a) * (double)i / (double)(n - 1)) : b;
    return r;
}
inline vec linspace(double a, double b, uword n) { return linspace<vec>(a, b, n); }
template <class C> struct conv_to {
    template <class U> static C from(const Col<U>& o) { C r(o.d.size()); for (size_t i = 0; i < o.d.size(); ++i) r.d[i] = o.d[i]; return r; }
};
// arma::find on a boolean-like column: indices of non-zero entries
template <class T> inline uvec find(const Col<T>& a, uword k, const char*) { uvec r; for (size_t i = 0; i < a.d.size() && r.d.size() < k; ++i) if (a.d[i]) r.d.push_back(i); r.fix(); return r; }
template <class T> inline uvec find(const Col<T>& a) { uvec r; for (size_t i = 0; i < a.d.size(); ++i) if (a.d[i]) r.d.push_back(i); r.fix(); return r; }
#define SHIM_CMP(op)                                                                        \
    template <class T> inline Col<int> operator op(const Col<T>& a, double b) { Col<int> r(a.d.size()); for (size_t i = 0; i < a.d.size(); ++i) r.d[i] = (a.d[i] op b); return r; }
SHIM_CMP(>) SHIM_CMP(<) SHIM_CMP(>=) SHIM_CMP(<=) SHIM_CMP(==)
#undef SHIM_CMP

// ---- dense column-major matrix -------------------------------------------------
template <class T> struct Mat {
    std::vector<T> d; uword n_rows = 0, n_cols = 0, n_elem = 0;
    Mat() {}
    Mat(uword r, uword c) : d(r * c), n_rows(r), n_cols(c), n_elem(r * c) {}
    Mat(uword r, uword c, fill::fill_zeros) : d(r * c, T(0)), n_rows(r), n_cols(c), n_elem(r * c) {}
    Mat(const Col<T>& v) : d(v.d), n_rows(v.d.size()), n_cols(1), n_elem(v.d.size()) {}
    T& operator()(uword i, uword j) { return d[i + j * n_rows]; }
    const T& operator()(uword i, uword j) const { return d[i + j * n_rows]; }
    T& at(uword i, uword j) { return d[i + j * n_rows]; }
    const T& at(uword i, uword j) const { return d[i + j * n_rows]; }
    T* memptr() { return d.data(); }
    const T* memptr() const { return d.data(); }
    void fill(T v) { std::fill(d.begin(), d.end(), v); }
    void zeros() { fill(T(0)); }
    Col<T> col(uword j) const { Col<T> r(n_rows); for (uword i = 0; i < n_rows; ++i) r.d[i] = (*this)(i, j); return r; }
    struct ColRef {   // writable column view
        Mat& m; uword j;
        ColRef& operator=(const Col<T>& o) { for (uword i = 0; i < m.n_rows; ++i) m(i, j) = o.d[i]; return *this; }
        operator Col<T>() const { Col<T> r(m.n_rows); for (uword i = 0; i < m.n_rows; ++i) r.d[i] = m(i, j); return r; }
    };
    ColRef col(uword j) { return ColRef{*this, j}; }
    // "a b c; d e f" initialiser (arma::mat from a string)
    explicit Mat(const char* txt) {
        std::vector<std::vector<T>> rows(1);
        std::istringstream in(txt); std::string tok;
        while (in >> tok) {
            bool end = !tok.empty() && tok.back() == ';';
            if (end) tok.pop_back();
            if (!tok.empty()) rows.back().push_back((T)std::stod(tok));
            if (end) rows.emplace_back();
        }
        if (rows.back().empty()) rows.pop_back();
        n_rows = rows.size(); n_cols = rows[0].size(); n_elem = n_rows * n_cols; d.resize(n_elem);
        for (uword i = 0; i < n_rows; ++i) for (uword j = 0; j < n_cols; ++j) (*this)(i, j) = rows[i][j];
    }
    Mat t() const { Mat r(n_cols, n_rows); for (uword i = 0; i < n_rows; ++i) for (uword j = 0; j < n_cols; ++j) r(j, i) = (*this)(i, j); return r; }
    // row access: a tiny proxy supporting read (as 1 x n Mat) and assignment from Col/Mat
    struct RowRef {
        Mat& m; uword i;
        RowRef& operator=(const Mat& o) { for (uword j = 0; j < m.n_cols; ++j) m(i, j) = o.d[j]; return *this; }
        RowRef& operator=(const Col<T>& o) { for (uword j = 0; j < m.n_cols; ++j) m(i, j) = o.d[j]; return *this; }
        operator Mat() const { Mat r(1, m.n_cols); for (uword j = 0; j < m.n_cols; ++j) r(0, j) = m(i, j); return r; }
        Mat t() const { Mat r(m.n_cols, 1); for (uword j = 0; j < m.n_cols; ++j) r(j, 0) = m(i, j); return r; }
    };
    RowRef row(uword i) { return RowRef{*this, i}; }
    Mat row(uword i) const { Mat r(1, n_cols); for (uword j = 0; j < n_cols; ++j) r(0, j) = (*this)(i, j); return r; }
    Mat rows(uword a, uword b) const { Mat r(b - a + 1, n_cols); for (uword i = a; i <= b; ++i) for (uword j = 0; j < n_cols; ++j) r(i - a, j) = (*this)(i, j); return r; }
};
typedef Mat<double> mat;
template <class T> struct Cube {
    std::vector<Mat<T>> s; uword n_rows = 0, n_cols = 0, n_slices = 0;
    Cube() {}
    Cube(uword r, uword c, uword k) : s(k, Mat<T>(r, c)), n_rows(r), n_cols(c), n_slices(k) {}
    Mat<T>& slice(uword k) { return s[k]; }
    const Mat<T>& slice(uword k) const { return s[k]; }
};
typedef Cube<double> cube;
template <class T> inline Mat<T> Col<T>::t() const { Mat<T> r(1, d.size()); for (size_t j = 0; j < d.size(); ++j) r(0, j) = d[j]; return r; }

struct DiagView { const vec& v; };
inline DiagView diagmat(const vec& v) { return DiagView{v}; }
inline mat ones(uword r, uword c) { mat m(r, c); m.fill(1.0); return m; }
inline mat zeros(uword r, uword c) { mat m(r, c); m.fill(0.0); return m; }
inline mat mean(const mat& a, int dim) {
    if (dim == 0) { mat m(1, a.n_cols); for (uword j = 0; j < a.n_cols; ++j) { double s = 0; for (uword i = 0; i < a.n_rows; ++i) s += a(i, j); m(0, j) = s / (double)a.n_rows; } return m; }
    mat m(a.n_rows, 1); for (uword i = 0; i < a.n_rows; ++i) { double s = 0; for (uword j = 0; j < a.n_cols; ++j) s += a(i, j); m(i, 0) = s / (double)a.n_cols; } return m;
}
inline mat operator*(const mat& a, const mat& b) {
    mat r(a.n_rows, b.n_cols);
    for (uword i = 0; i < a.n_rows; ++i) for (uword j = 0; j < b.n_cols; ++j) { double s = 0; for (uword k = 0; k < a.n_cols; ++k) s += a(i, k) * b(k, j); r(i, j) = s; }
    return r;
}
inline mat operator*(const mat& a, const DiagView& dg) { mat r = a; for (uword j = 0; j < a.n_cols; ++j) for (uword i = 0; i < a.n_rows; ++i) r(i, j) *= dg.v.d[j]; return r; }
inline vec operator*(const mat& a, const vec& x) { vec r(a.n_rows); for (uword i = 0; i < a.n_rows; ++i) { double s = 0; for (uword k = 0; k < a.n_cols; ++k) s += a(i, k) * x.d[k]; r.d[i] = s; } return r; }
inline mat operator-(const mat& a, const mat& b) { mat r = a; for (size_t i = 0; i < r.d.size(); ++i) r.d[i] -= b.d[i]; return r; }
inline mat operator+(const mat& a, const mat& b) { mat r = a; for (size_t i = 0; i < r.d.size(); ++i) r.d[i] += b.d[i]; return r; }
inline mat operator*(double s, const mat& a) { mat r = a; for (auto& v : r.d) v *= s; return r; }
// upper-triangular Cholesky factor U with U^T U = A (arma::chol default)
inline mat chol(const mat& A) {
    uword n = A.n_rows; mat U(n, n); U.fill(0.0);
    for (uword j = 0; j < n; ++j)
        for (uword i = 0; i <= j; ++i) {
            double s = A(i, j);
            for (uword k = 0; k < i; ++k) s -= U(k, i) * U(k, j);
            U(i, j) = (i == j) ? std::sqrt(s) : s / U(i, i);
        }
    return U;
}
}  // namespace arma

// ---------------------------------------------------------------------------
// R nmath subset
// ---------------------------------------------------------------------------
inline double unif_rand() { return shim::unif01(); }
inline double norm_rand() { return shim::norm01(); }
namespace R {
inline double runif(double a, double b) { return a + (b - a) * shim::unif01(); }
inline double rnorm(double m, double s) { return m + s * shim::norm01(); }
inline double dnorm(double x, double mu, double sigma, int give_log) {
    const double M_LN_SQRT_2PI_ = 0.918938533204672741780329736406;
    double z = (x - mu) / sigma;
    if (give_log) return -(M_LN_SQRT_2PI_ + 0.5 * z * z + std::log(sigma));
    return std::exp(-0.5 * z * z) / (sigma * 2.50662827463100050241576528481);
}
// Gamma(shape, scale).  When uniforms are injected and the shape is a small integer k the draw is the sum of k
// exponentials, -scale * sum log(u_i) (k uniforms consumed); otherwise std::gamma_distribution.  Parity-unpinned
// like the rest of the RNG surface (R's rgamma is Marsaglia-Tsang / Ahrens-Dieter).
inline double rgamma(double shape, double scale) {
    shim::State& s = shim::st();
    int k = (int)shape;
    if (!s.uq.empty() && (double)k == shape && k >= 1 && k <= 8) {
        double acc = 0.0;
        for (int i = 0; i < k; ++i) acc += -std::log(shim::unif01());
        return scale * acc;
    }
    std::gamma_distribution<double> g(shape, scale);
    ++s.n_unif;
    return g(s.eng);
}
}  // namespace R

// R nmath rmultinom control flow: K-1 sequential conditional binomials, remainder in the last cell.
inline void rmultinom(int n, double* prob, int K, int* rN) {
    shim::State& s = shim::st(); ++s.n_multinom;
    if (!s.mq.empty()) {
        std::vector<int> c = s.mq.front(); s.mq.pop_front();
        for (int k = 0; k < K; ++k) rN[k] = c[k];
        return;
    }
    long double p_tot = 0;
    for (int k = 0; k < K; ++k) { p_tot += prob[k]; rN[k] = 0; }
    if (n == 0) return;
    for (int k = 0; k < K - 1; ++k) {
        if (prob[k]) {
            double pp = (double)(prob[k] / p_tot);
            rN[k] = (pp < 1.) ? (int)std::binomial_distribution<int>(n, pp)(s.eng) : n;
            n -= rN[k];
        } else rN[k] = 0;
        if (n <= 0) return;
        p_tot -= prob[k];
    }
    rN[K - 1] = n;
}

// ---------------------------------------------------------------------------
// Rcpp subset (containers are thin host vectors; List/DataFrame keep named columns)
// ---------------------------------------------------------------------------
struct SEXPREC_shim;
#define R_NilValue (Rcpp::nil())
namespace Rcpp {
inline void stop(const char* m) { throw std::runtime_error(m); }
inline void stop(const std::string& m) { throw std::runtime_error(m); }
// Rcout / Rcerr: private streams that discard their output.  (Not std::cout: inside a Python process a second libstdc++
// may own the global stream objects, and formatting numbers through them crashes.)
struct NullBuf : std::streambuf { int overflow(int c) override { return c; } };
inline std::ostream& null_stream() { static NullBuf nb; static std::ostream os(&nb); return os; }
static std::ostream& Rcout = null_stream();
static std::ostream& Rcerr = null_stream();
struct Nil {};
inline Nil nil() { return Nil{}; }

template <class T> struct Vector_ {
    std::vector<T> v;
    Vector_() {}
    Vector_(long n) : v(n) {}
    Vector_(int n) : v(n) {}
    Vector_(unsigned long n) : v(n) {}
    Vector_(const std::vector<T>& o) : v(o) {}
    Vector_(const T* first, const T* last) : v(first, last) {}   // iterator-range constructor (glue syntax check)
    Vector_(long n, T fill) : v(n, fill) {}
    Vector_(unsigned long n, T fill) : v(n, fill) {}
    template <class U> Vector_(const arma::Col<U>& o) : v(o.d.begin(), o.d.end()) {}
    T& operator()(long i) { return v[i]; }
    const T& operator()(long i) const { return v[i]; }
    T& operator[](long i) { return v[i]; }
    const T& operator[](long i) const { return v[i]; }
    long size() const { return (long)v.size(); }
    long length() const { return (long)v.size(); }
    T* begin() { return v.data(); }
    T* end() { return v.data() + v.size(); }
};
typedef Vector_<double> NumericVector;
typedef Vector_<int> IntegerVector;

struct Column { std::string name; std::vector<double> val; double scalar = 0; bool is_scalar = false; };
struct NamedPH {
    std::string n;
    Column operator=(const NumericVector& x) const { return Column{n, x.v}; }
    Column operator=(const IntegerVector& x) const { return Column{n, std::vector<double>(x.v.begin(), x.v.end())}; }
    Column operator=(const arma::vec& x) const { return Column{n, x.d}; }
    Column operator=(const arma::mat& x) const { return Column{n, x.d}; }
    Column operator=(const arma::cube& x) const { std::vector<double> v; for (auto& m : x.s) v.insert(v.end(), m.d.begin(), m.d.end()); return Column{n, v}; }
    Column operator=(const std::vector<double>& x) const { return Column{n, x}; }
    Column operator=(double x) const { Column c{n, {x}}; c.scalar = x; c.is_scalar = true; return c; }
};
inline NamedPH Named(const char* s) { return NamedPH{s}; }
struct Underscore { NamedPH operator[](const char* s) const { return NamedPH{s}; } };
static const Underscore _{};
struct ListSlot;
struct List {
    std::vector<Column> items; bool is_nil = false;
    ListSlot operator[](const char* name);
    List() {}
    List(Nil) : is_nil(true) {}
    template <class... A> static List create(A... a) { List l; (l.items.push_back(a), ...); return l; }
    const Column* find(const std::string& n) const { for (auto& c : items) if (c.name == n) return &c; return nullptr; }
};
struct ListSlot {   // list["name"] as an rvalue scalar or an assignment target
    List& l; std::string name;
    operator double() const { const Column* c = l.find(name); if (!c) throw std::runtime_error("list element missing: " + name); return c->is_scalar ? c->scalar : c->val.at(0); }
    ListSlot& operator=(const arma::mat& m) { l.items.push_back(Column{name, m.d}); return *this; }
    ListSlot& operator=(double v) { Column c{name, {v}}; c.scalar = v; c.is_scalar = true; l.items.push_back(c); return *this; }
};
inline ListSlot List::operator[](const char* name) { return ListSlot{*this, name}; }
struct DataFrame : List {
    DataFrame() {}
    DataFrame(Nil n) : List(n) {}
    DataFrame(const List& l) : List(l) {}
    template <class... A> static DataFrame create(A... a) { DataFrame l; (l.items.push_back(a), ...); return l; }
};
// an R closure: the shim can only hold a C++ callback taking the two running-mean vectors
struct Function {
    void (*cb)(const NumericVector&, const NumericVector&) = nullptr;
    Function() {}
    Function(void (*f)(const NumericVector&, const NumericVector&)) : cb(f) {}
    void operator()(const NumericVector& a, const NumericVector& b) const { if (cb) cb(a, b); }
};
inline NumericVector rnorm(long n, double m = 0.0, double s = 1.0) { NumericVector r(n); for (long i = 0; i < n; ++i) r.v[i] = m + s * shim::norm01(); return r; }
inline NumericVector runif(long n, double a = 0.0, double b = 1.0) { NumericVector r(n); for (long i = 0; i < n; ++i) r.v[i] = a + (b - a) * shim::unif01(); return r; }
template <class T> inline Vector_<T> operator-(const Vector_<T>& a, int b) { Vector_<T> r(a); for (auto& v : r.v) v -= b; return r; }
inline NumericVector wrap(const arma::vec& x) { return NumericVector(x.d); }
// Rcpp::sample(n, size, replace, probs): 1-based draws by sequential inverse CDF on the normalised probabilities, one
// unif_rand() per draw.  (R sorts the probabilities first and switches to Walker's alias method for n >= 200; this
// stand-in keeps the natural order -- same distribution, parity-unpinned like the rest of the RNG surface.)
inline IntegerVector sample(int n, int size, bool /*replace*/, const NumericVector& probs) {
    double tot = 0.0; for (int i = 0; i < n; ++i) tot += probs.v[i];
    IntegerVector r(size);
    for (int s = 0; s < size; ++s) {
        double u = shim::unif01(), c = 0.0; int j = 0;
        for (; j < n - 1; ++j) { c += probs.v[j] / tot; if (u <= c) break; }
        r.v[s] = j + 1;
    }
    return r;
}
template <class T> inline T as(const NumericVector& x);
template <> inline arma::vec as<arma::vec>(const NumericVector& x) { return arma::vec(x.v); }
template <> inline NumericVector as<NumericVector>(const NumericVector& x) { return x; }
}  // namespace Rcpp
