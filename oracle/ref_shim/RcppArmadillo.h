// RcppArmadillo.h -- MINIMAL STAND-IN (test infrastructure, written for this repo) for the
// subset of Rcpp / Armadillo that the reference's hot-path sources use, so that the
// reference's OWN, UNMODIFIED files
//     /root/reference/src/collapsed_gibbs_z_VB.cpp, collapsed_gibbs_z.cpp, commons.cpp,
//     sample_gibbs.cpp
// can be compiled where they lie (oracle/Makefile target `ref` -> oracle/_ref/libnbmf_ref.so)
// and used to pin the oracle.  Nothing here is copied from Armadillo or Rcpp: containers are
// plain column-major arrays, every expression is evaluated eagerly and element-wise in the
// order it is written, accu() is a left-to-right sum.  R's RNG (rmultinom / rbeta / runif) and
// arma::randg are replaced by the same xoshiro256** stream the oracle uses, so sampling code
// can be compared draw for draw.  boost::math::digamma / lgamma are provided by
// boost/math/special_functions/*.hpp next to this file.
#ifndef NBMF_REF_SHIM_RCPPARMADILLO_H
#define NBMF_REF_SHIM_RCPPARMADILLO_H

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <iterator>
#include <limits>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#define NA_INTEGER INT_MIN

// ---------------------------------------------------------------------------------------
// RNG shared with oracle/nbmf_oracle.c (same algorithm, same draw order)
// ---------------------------------------------------------------------------------------
namespace shim {
struct Rng {
    uint64_t s[4];
    static uint64_t splitmix(uint64_t &x)
    {
        uint64_t z = (x += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    void seed(uint64_t v) { for (int i = 0; i < 4; i++) s[i] = splitmix(v); }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next()
    {
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double unif() { return ((double)(next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
    double norm() { double u = unif(), v = unif(); return std::sqrt(-2.0 * std::log(u)) * std::cos(2.0 * M_PI * v); }
    double gamma(double a)
    {
        if (a < 1.0) { double u = unif(); return gamma(a + 1.0) * std::pow(u, 1.0 / a); }
        double d = a - 1.0 / 3.0, c = 1.0 / std::sqrt(9.0 * d);
        for (;;) {
            double x = norm(), v = 1.0 + c * x;
            if (v <= 0.0) continue;
            v = v * v * v;
            double u = unif();
            if (u < 1.0 - 0.0331 * x * x * x * x) return d * v;
            if (std::log(u) < 0.5 * x * x + d * (1.0 - v + std::log(v))) return d * v;
        }
    }
};
inline Rng &rng() { static Rng g = [] { Rng r; r.seed(1); return r; }(); return g; }
inline int &warning_count() { static int c = 0; return c; }
} // namespace shim

// ---------------------------------------------------------------------------------------
// arma
// ---------------------------------------------------------------------------------------
namespace arma {
typedef unsigned long long uword;
namespace datum { static const double nan = std::numeric_limits<double>::quiet_NaN(); }

template <class T> class Vec;
template <class T> class Mat;

// strided view of one row of a column-major matrix
template <class T> class subview_row {
public:
    T *p; uword stride, n_elem;
    subview_row(T *p_, uword s, uword n) : p(p_), stride(s), n_elem(n) {}
    T &operator[](uword i) const { return p[i * stride]; }
    T &operator()(uword i) const { return p[i * stride]; }
    operator Vec<T>() const;
    const subview_row &operator=(const Vec<T> &v) const { for (uword i = 0; i < n_elem; i++) p[i * stride] = v[i]; return *this; }
    const subview_row &operator=(const subview_row &o) const { Vec<T> t = o; return *this = t; }
    const subview_row &operator+=(const Vec<T> &v) const { for (uword i = 0; i < n_elem; i++) p[i * stride] += v[i]; return *this; }
    const subview_row &operator-=(const Vec<T> &v) const { for (uword i = 0; i < n_elem; i++) p[i * stride] -= v[i]; return *this; }
    const subview_row &operator/=(T s) const { for (uword i = 0; i < n_elem; i++) p[i * stride] /= s; return *this; }
    const subview_row &zeros() const { for (uword i = 0; i < n_elem; i++) p[i * stride] = T(0); return *this; }
    Vec<T> t() const;
    struct iterator {
        typedef std::forward_iterator_tag iterator_category; typedef T value_type; typedef std::ptrdiff_t difference_type;
        typedef T *pointer; typedef T &reference;
        T *p; uword stride;
        T &operator*() const { return *p; }
        iterator &operator++() { p += stride; return *this; }
        iterator operator++(int) { iterator t = *this; p += stride; return t; }
        bool operator==(const iterator &o) const { return p == o.p; }
        bool operator!=(const iterator &o) const { return p != o.p; }
    };
    iterator begin() const { return iterator{p, stride}; }
    iterator end() const { return iterator{p + n_elem * stride, stride}; }
};

template <class T> class Vec {
public:
    typedef T elem_type;
    std::vector<T> d; uword n_elem;
    Vec() : n_elem(0) {}
    explicit Vec(uword n) : d(n, T(0)), n_elem(n) {}
    T &operator[](uword i) { return d[i]; }
    const T &operator[](uword i) const { return d[i]; }
    T &operator()(uword i) { return d[i]; }
    const T &operator()(uword i) const { return d[i]; }
    T *begin() { return d.data(); }
    const T *begin() const { return d.data(); }
    T *end() { return d.data() + n_elem; }
    T *memptr() { return d.data(); }
    const T *memptr() const { return d.data(); }
    Vec &zeros() { std::fill(d.begin(), d.end(), T(0)); return *this; }
    Vec &fill(T v) { std::fill(d.begin(), d.end(), v); return *this; }
    void set_size(uword n) { d.assign(n, T(0)); n_elem = n; }
    Vec t() const { return *this; }
    Vec &operator+=(const Vec &o) { for (uword i = 0; i < n_elem; i++) d[i] += o[i]; return *this; }
    Vec &operator-=(const Vec &o) { for (uword i = 0; i < n_elem; i++) d[i] -= o[i]; return *this; }
    Vec &operator/=(T s) { for (uword i = 0; i < n_elem; i++) d[i] /= s; return *this; }
    void print(const char * = "") const {}
};
template <class T> subview_row<T>::operator Vec<T>() const { Vec<T> v(n_elem); for (uword i = 0; i < n_elem; i++) v[i] = p[i * stride]; return v; }
template <class T> Vec<T> subview_row<T>::t() const { return Vec<T>(*this); }

typedef Vec<double> vec; typedef Vec<double> rowvec; typedef Vec<double> colvec;
typedef Vec<int> ivec; typedef Vec<int> irowvec;
typedef Vec<int> ubvec;

template <class T> class Mat {
public:
    typedef T elem_type;
    std::vector<T> own; T *mem; uword n_rows, n_cols, n_elem;
    Mat() : mem(nullptr), n_rows(0), n_cols(0), n_elem(0) {}
    Mat(uword r, uword c) : own(r * c, T(0)), mem(nullptr), n_rows(r), n_cols(c), n_elem(r * c) { mem = own.data(); }
    Mat(T *ext, uword r, uword c) : mem(ext), n_rows(r), n_cols(c), n_elem(r * c) {} // view
    Mat(const Mat &o) : own(o.mem, o.mem + o.n_elem), mem(nullptr), n_rows(o.n_rows), n_cols(o.n_cols), n_elem(o.n_elem) { mem = own.data(); }
    Mat &operator=(const Mat &o)
    {
        if (this == &o) return *this;
        if (own.empty() && mem != nullptr) { // assigning into a view (cube.slice(j) = Z): copy elements
            std::copy(o.mem, o.mem + std::min(n_elem, o.n_elem), mem);
            return *this;
        }
        own.assign(o.mem, o.mem + o.n_elem); mem = own.data(); n_rows = o.n_rows; n_cols = o.n_cols; n_elem = o.n_elem;
        return *this;
    }
    void set_size(uword r, uword c) { own.assign(r * c, T(0)); mem = own.data(); n_rows = r; n_cols = c; n_elem = r * c; }
    T &operator()(uword i, uword j) { return mem[i + n_rows * j]; }
    const T &operator()(uword i, uword j) const { return mem[i + n_rows * j]; }
    T *memptr() { return mem; }
    const T *memptr() const { return mem; }
    subview_row<T> row(uword i) { return subview_row<T>(mem + i, n_rows, n_cols); }
    subview_row<T> row(uword i) const { return subview_row<T>(const_cast<T *>(mem) + i, n_rows, n_cols); }
    subview_row<T> col(uword j) const { return subview_row<T>(const_cast<T *>(mem) + n_rows * j, 1, n_rows); }
    Mat &zeros() { std::fill(mem, mem + n_elem, T(0)); return *this; }
    Mat &fill(T v) { std::fill(mem, mem + n_elem, v); return *this; }
    Mat t() const { Mat r(n_cols, n_rows); for (uword i = 0; i < n_rows; i++) for (uword j = 0; j < n_cols; j++) r(j, i) = (*this)(i, j); return r; }
    T max() const { return *std::max_element(mem, mem + n_elem); }
    void print(const char * = "") const {}
};
typedef Mat<double> mat; typedef Mat<int> imat;

template <class T> class Cube {
public:
    typedef T elem_type;
    std::vector<T> d; uword n_rows, n_cols, n_slices, n_elem;
    Cube() : n_rows(0), n_cols(0), n_slices(0), n_elem(0) {}
    Cube(uword r, uword c, uword s) : d(r * c * s, T(0)), n_rows(r), n_cols(c), n_slices(s), n_elem(r * c * s) {}
    void set_size(uword r, uword c, uword s) { d.assign(r * c * s, T(0)); n_rows = r; n_cols = c; n_slices = s; n_elem = r * c * s; }
    T &operator()(uword i, uword j, uword k) { return d[i + n_rows * j + n_rows * n_cols * k]; }
    const T &operator()(uword i, uword j, uword k) const { return d[i + n_rows * j + n_rows * n_cols * k]; }
    Mat<T> slice(uword k) { return Mat<T>(d.data() + n_rows * n_cols * k, n_rows, n_cols); }
    const Mat<T> slice(uword k) const { return Mat<T>(const_cast<T *>(d.data()) + n_rows * n_cols * k, n_rows, n_cols); }
    Cube &zeros() { std::fill(d.begin(), d.end(), T(0)); return *this; }
    T max() const { return *std::max_element(d.begin(), d.end()); }
    T *memptr() { return d.data(); }
    const T *memptr() const { return d.data(); }
    Cube &operator+=(const Cube &o) { for (uword i = 0; i < n_elem; i++) d[i] += o.d[i]; return *this; }
    template <class Fn> Cube &transform(Fn f) { for (uword i = 0; i < n_elem; i++) d[i] = f(d[i]); return *this; }
};
typedef Cube<double> cube; typedef Cube<int> icube;
template <class T> inline Cube<T> operator/(const Cube<T> &a, double s) { Cube<T> r(a.n_rows, a.n_cols, a.n_slices); for (uword i = 0; i < a.n_elem; i++) r.d[i] = a.d[i] / s; return r; }

// ---- element-wise vector algebra on Vec<double> (subview_row<double> converts implicitly) ----
#define SHIM_VV(op, name) \
    inline vec name(const vec &a, const vec &b) { vec r(a.n_elem); for (uword i = 0; i < a.n_elem; i++) r[i] = a[i] op b[i]; return r; }
SHIM_VV(+, operator+) SHIM_VV(-, operator-) SHIM_VV(*, operator%) SHIM_VV(/, operator/)
#undef SHIM_VV
#define SHIM_VS(op) \
    inline vec operator op(const vec &a, double s) { vec r(a.n_elem); for (uword i = 0; i < a.n_elem; i++) r[i] = a[i] op s; return r; } \
    inline vec operator op(double s, const vec &a) { vec r(a.n_elem); for (uword i = 0; i < a.n_elem; i++) r[i] = s op a[i]; return r; }
SHIM_VS(+) SHIM_VS(-) SHIM_VS(*) SHIM_VS(/)
#undef SHIM_VS
inline ivec operator+(const ivec &a, const ivec &b) { ivec r(a.n_elem); for (uword i = 0; i < a.n_elem; i++) r[i] = a[i] + b[i]; return r; }
inline ivec operator<(const vec &a, double s) { ivec r(a.n_elem); for (uword i = 0; i < a.n_elem; i++) r[i] = a[i] < s; return r; }
inline bool any(const ivec &a) { for (uword i = 0; i < a.n_elem; i++) if (a[i]) return true; return false; }
inline vec vectorise(const vec &v) { return v; }
inline vec vectorise(const mat &m) { vec r(m.n_elem); for (uword i = 0; i < m.n_elem; i++) r[i] = m.mem[i]; return r; }
inline double accu(const vec &a) { double s = 0.0; for (uword i = 0; i < a.n_elem; i++) s += a[i]; return s; }
inline double accu(const mat &a) { double s = 0.0; for (uword i = 0; i < a.n_elem; i++) s += a.mem[i]; return s; }
inline double sum(const vec &a) { return accu(a); }
inline vec exp(const vec &a) { vec r(a.n_elem); for (uword i = 0; i < a.n_elem; i++) r[i] = std::exp(a[i]); return r; }
inline uword index_max(const ivec &a) { uword b = 0; for (uword i = 1; i < a.n_elem; i++) if (a[i] > a[b]) b = i; return b; }

inline mat operator+(const mat &a, const mat &b) { mat r(a.n_rows, a.n_cols); for (uword i = 0; i < a.n_elem; i++) r.mem[i] = a.mem[i] + b.mem[i]; return r; }
inline mat operator/(const mat &a, double s) { mat r(a.n_rows, a.n_cols); for (uword i = 0; i < a.n_elem; i++) r.mem[i] = a.mem[i] / s; return r; }
inline mat operator*(const mat &a, const mat &b)
{
    mat r(a.n_rows, b.n_cols);
    for (uword j = 0; j < b.n_cols; j++) for (uword k = 0; k < a.n_cols; k++) { double bk = b(k, j); for (uword i = 0; i < a.n_rows; i++) r(i, j) += a(i, k) * bk; }
    return r;
}
// sum over a dimension: matrices (dim 0 = over rows -> one value per column)
inline ivec sum(const imat &m, int) { ivec r(m.n_cols); for (uword j = 0; j < m.n_cols; j++) for (uword i = 0; i < m.n_rows; i++) r[j] += m(i, j); return r; }
// cubes: dim 2 -> rows x cols (sum over slices); dim 0 -> cols x slices (sum over rows)
inline mat sum(const cube &c, int dim)
{
    if (dim == 2) { mat r(c.n_rows, c.n_cols); for (uword s = 0; s < c.n_slices; s++) for (uword j = 0; j < c.n_cols; j++) for (uword i = 0; i < c.n_rows; i++) r(i, j) += c(i, j, s); return r; }
    mat r(c.n_cols, c.n_slices);
    for (uword s = 0; s < c.n_slices; s++) for (uword j = 0; j < c.n_cols; j++) for (uword i = 0; i < c.n_rows; i++) r(j, s) += c(i, j, s);
    return r;
}

template <class Out> struct conv_to {
    typedef typename Out::elem_type E;
    template <class U> static Out from(const subview_row<U> &v) { Out r(v.n_elem); for (uword i = 0; i < v.n_elem; i++) r[i] = (E)v[i]; return r; }
    template <class U> static Out from(const Vec<U> &v) { Out r(v.n_elem); for (uword i = 0; i < v.n_elem; i++) r[i] = (E)v[i]; return r; }
    template <class U> static Out from(const Mat<U> &m) { Out r(m.n_rows, m.n_cols); for (uword i = 0; i < m.n_elem; i++) r.mem[i] = (E)m.mem[i]; return r; }
    template <class U> static Out from(const Cube<U> &c) { Out r(c.n_rows, c.n_cols, c.n_slices); for (uword i = 0; i < c.n_elem; i++) r.d[i] = (E)c.d[i]; return r; }
};

struct distr_param { double a, b; distr_param(double a_, double b_) : a(a_), b(b_) {} };
inline vec randg(uword n, const distr_param &p) { vec r(n); for (uword i = 0; i < n; i++) r[i] = shim::rng().gamma(p.a) * p.b; return r; }

template <class T> inline std::ostream &operator<<(std::ostream &o, const Mat<T> &) { return o; }
template <class T> inline std::ostream &operator<<(std::ostream &o, const Vec<T> &) { return o; }
} // namespace arma

// ---------------------------------------------------------------------------------------
// Rcpp
// ---------------------------------------------------------------------------------------
namespace shim {
struct Obj { // what a SEXP points to in this shim
    int kind = 0; // 0 nil, 1 real array, 2 int array, 3 list
    std::vector<unsigned long long> dim;
    std::vector<double> real;
    std::vector<int> integer;
    std::vector<std::pair<std::string, Obj *>> items;
    ~Obj() { for (auto &it : items) delete it.second; }
    Obj *get(const std::string &n) const { for (auto &it : items) if (it.first == n) return it.second; return nullptr; }
};
struct NullBuf : std::streambuf { int overflow(int c) override { return c; } };
} // namespace shim
typedef shim::Obj *SEXP;
#define R_NilValue ((SEXP) nullptr)

namespace Rcpp {
inline std::ostream &rcout_stream() { static shim::NullBuf nb; static std::ostream os(&nb); return os; }
#define Rcout rcout_stream()
inline void stop(const char *msg) { throw std::runtime_error(msg); }
inline void stop(const std::string &msg) { throw std::runtime_error(msg); }
inline void warning(const char *) { shim::warning_count()++; }

inline SEXP wrap(const arma::mat &m) { SEXP o = new shim::Obj; o->kind = 1; o->dim = {m.n_rows, m.n_cols}; o->real.assign(m.mem, m.mem + m.n_elem); return o; }
inline SEXP wrap(const arma::imat &m) { SEXP o = new shim::Obj; o->kind = 2; o->dim = {m.n_rows, m.n_cols}; o->integer.assign(m.mem, m.mem + m.n_elem); return o; }
inline SEXP wrap(const arma::vec &v) { SEXP o = new shim::Obj; o->kind = 1; o->dim = {v.n_elem}; o->real = v.d; return o; }
inline SEXP wrap(const arma::cube &c) { SEXP o = new shim::Obj; o->kind = 1; o->dim = {c.n_rows, c.n_cols, c.n_slices}; o->real = c.d; return o; }
inline SEXP wrap(const arma::icube &c) { SEXP o = new shim::Obj; o->kind = 2; o->dim = {c.n_rows, c.n_cols, c.n_slices}; o->integer = c.d; return o; }
inline SEXP wrap(double x) { SEXP o = new shim::Obj; o->kind = 1; o->dim = {1}; o->real = {x}; return o; }

struct NamedArg { std::string name; SEXP value; };
struct Named {
    std::string name;
    explicit Named(const char *n) : name(n) {}
    NamedArg operator=(SEXP v) const { return NamedArg{name, v}; }
};
struct List {
    static void add(SEXP) {}
    template <class... Rest> static void add(SEXP l, const NamedArg &a, const Rest &...rest) { l->items.emplace_back(a.name, a.value); add(l, rest...); }
    template <class... Args> static SEXP create(const Args &...args) { SEXP l = new shim::Obj; l->kind = 3; add(l, args...); return l; }
};
struct NumericVector { double v; };
inline NumericVector rbeta(int, double a, double b) { double x = shim::rng().gamma(a), y = shim::rng().gamma(b); return NumericVector{x / (x + y)}; }
inline NumericVector runif(int) { return NumericVector{shim::rng().unif()}; }
template <class T> inline T as(const NumericVector &v) { return (T)v.v; }
} // namespace Rcpp

// R API: one categorical draw as a one-hot vector (rmultinom(1, prob, K, ans)), inverse CDF on
// the shim RNG -- identical to orng_categorical() in oracle/nbmf_oracle.c
inline void rmultinom(int, double *prob, int K, int *ans)
{
    double u = shim::rng().unif(), c = 0.0;
    int k = -1;
    for (int i = 0; i < K; i++) { c += prob[i]; if (u < c) { k = i; break; } }
    if (k < 0) { for (int i = K - 1; i >= 0; i--) if (prob[i] > 0.0) { k = i; break; } if (k < 0) k = K - 1; }
    for (int i = 0; i < K; i++) ans[i] = 0;
    ans[k] = 1;
}
#endif
