// glue_driver.cpp -- extern "C" access to integration/nbmf_glue.cpp compiled against oracle/ref_shim
// (test infrastructure: the shim stands in for Rcpp/Armadillo, which the image does not have).
#include <RcppArmadillo.h>

SEXP VB_Aspect_Rcpp_cuda(const arma::imat &V, arma::imat Z, int K, double alpha, double beta, double gamma, int iter, bool checklb);
SEXP VB_DirBer_Rcpp_cuda(const arma::imat &V, arma::imat Z, int K, double alpha, double beta, double gamma, int iter);
SEXP Gibbs_DirBer_finite_Rcpp_cuda(const arma::imat &V, arma::imat Z, int K, double alpha, double beta, double gamma, int iter, double burnin);
double lower_bound_aspect_cuda(const arma::cube &E_Z, const arma::mat &E_log_W, const arma::mat &E_log_H, const arma::mat &E_log_1_H,
                               const arma::mat &gamma_vb, const arma::mat &alpha_vb, const arma::mat &beta_vb, double alpha, double beta,
                               double gamma, const arma::imat &V);
arma::mat compute_expectation_W_Rcpp_cuda(const arma::icube &Z_samples, double gamma);
arma::mat compute_expectation_H_Rcpp_cuda(const arma::icube &Z_samples, const arma::imat &V, double alpha, double beta);
SEXP samples_VWH_DirBer_cuda(const arma::imat &V, const arma::icube &Z_samples, double gamma, double alpha, double beta);
SEXP sample_gibbs_cpp_cuda(const arma::ivec &v_n, const arma::mat &W, arma::imat C, double alpha, int iter, double burnin);
extern unsigned long long nbmf_glue_last_seed;

static arma::imat to_imat(const int *p, int r, int c) { arma::imat m(r, c); std::copy(p, p + (size_t)r * c, m.mem); return m; }
static void copy_real(SEXP l, const char *name, double *dst)
{
    shim::Obj *o = l->get(name);
    if (o && dst) std::copy(o->real.begin(), o->real.end(), dst);
}
static arma::mat to_mat(const double *p, int r, int c) { arma::mat m(r, c); std::copy(p, p + (size_t)r * c, m.memptr()); return m; }
static arma::cube to_cube(const double *p, int r, int c, int s) { arma::cube m(r, c, s); std::copy(p, p + (size_t)r * c * s, m.memptr()); return m; }
static arma::icube to_icube(const int *p, int r, int c, int s) { arma::icube m(r, c, s); std::copy(p, p + (size_t)r * c * s, m.memptr()); return m; }
static void copy_int(SEXP l, const char *name, int *dst)
{
    shim::Obj *o = l->get(name);
    if (o && dst) std::copy(o->integer.begin(), o->integer.end(), dst);
}
static char g_err[512];

extern "C" {
const char *glue_last_error() { return g_err; }
int glue_warning_count() { return shim::warning_count(); }

int glue_vb_aspect(const int *V, const int *Z, int F, int N, int K, double alpha, double beta, double gamma, int iter, int checklb,
                   double *E_W, double *E_H, double *lowerbounds, double *likelihoods, double *E_Z)
{
    try {
        SEXP r = VB_Aspect_Rcpp_cuda(to_imat(V, F, N), to_imat(Z, F, N), K, alpha, beta, gamma, iter, checklb != 0);
        copy_real(r, "E_W", E_W); copy_real(r, "E_H", E_H); copy_real(r, "lowerbounds", lowerbounds);
        copy_real(r, "likelihoods", likelihoods); copy_real(r, "E_Z", E_Z);
        delete r;
        return 0;
    } catch (const std::exception &e) { snprintf(g_err, sizeof g_err, "%s", e.what()); return 1; }
}

int glue_vb_dirber(const int *V, const int *Z, int F, int N, int K, double alpha, double beta, double gamma, int iter, double *E_W,
                   double *E_H, double *E_Z)
{
    try {
        SEXP r = VB_DirBer_Rcpp_cuda(to_imat(V, F, N), to_imat(Z, F, N), K, alpha, beta, gamma, iter);
        copy_real(r, "E_W", E_W); copy_real(r, "E_H", E_H); copy_real(r, "E_Z", E_Z);
        delete r;
        return 0;
    } catch (const std::exception &e) { snprintf(g_err, sizeof g_err, "%s", e.what()); return 1; }
}

// returns the number of slices; *seed_used = the seed the body derived from the (shim's) R RNG
int glue_gibbs_dirber_finite(const int *V, const int *Z, int F, int N, int K, double alpha, double beta, double gamma, int iter,
                             double burnin, unsigned long long rng_seed, int *Z_samples, unsigned long long *seed_used)
{
    try {
        shim::rng().seed(rng_seed);                      // "set.seed()"
        SEXP r = Gibbs_DirBer_finite_Rcpp_cuda(to_imat(V, F, N), to_imat(Z, F, N), K, alpha, beta, gamma, iter, burnin);
        shim::Obj *zs = r->get("Z_samples");
        int ns = zs->dim.size() == 3 ? (int)zs->dim[2] : 0;
        if (Z_samples) std::copy(zs->integer.begin(), zs->integer.end(), Z_samples);
        if (seed_used) *seed_used = nbmf_glue_last_seed;
        delete r;
        return ns;
    } catch (const std::exception &e) { snprintf(g_err, sizeof g_err, "%s", e.what()); return -1; }
}

int glue_lower_bound_aspect(const double *E_Z, const double *ElW, const double *ElH, const double *El1H, const double *gvb,
                            const double *avb, const double *bvb, double alpha, double beta, double gamma, const int *V, int F, int N,
                            int K, double *lb)
{
    try {
        *lb = lower_bound_aspect_cuda(to_cube(E_Z, F, K, N), to_mat(ElW, F, K), to_mat(ElH, N, K), to_mat(El1H, N, K), to_mat(gvb, F, K),
                                      to_mat(avb, N, K), to_mat(bvb, N, K), alpha, beta, gamma, to_imat(V, F, N));
        return 0;
    } catch (const std::exception &e) { snprintf(g_err, sizeof g_err, "%s", e.what()); return 1; }
}

// E_W F x Kq, E_H N x Kq with Kq = max label + 1; returns Kq
int glue_compute_expectations(const int *Zs, const int *V, int F, int N, int J, double alpha, double beta, double gamma, double *E_W,
                              double *E_H)
{
    try {
        arma::icube Z = to_icube(Zs, F, N, J);
        arma::mat w = compute_expectation_W_Rcpp_cuda(Z, gamma);
        arma::mat h = compute_expectation_H_Rcpp_cuda(Z, to_imat(V, F, N), alpha, beta);
        std::copy(w.memptr(), w.memptr() + (size_t)w.n_rows * w.n_cols, E_W);
        std::copy(h.memptr(), h.memptr() + (size_t)h.n_rows * h.n_cols, E_H);
        return (int)w.n_cols;
    } catch (const std::exception &e) { snprintf(g_err, sizeof g_err, "%s", e.what()); return -1; }
}

int glue_samples_VWH(const int *V, const int *Zs, int F, int N, int J, int K, double gamma, double alpha, double beta,
                     unsigned long long rng_seed, double *W_samples, double *H_samples, int *V_samples)
{
    try {
        shim::rng().seed(rng_seed);
        SEXP r = samples_VWH_DirBer_cuda(to_imat(V, F, N), to_icube(Zs, F, N, J), gamma, alpha, beta);
        copy_real(r, "W_samples", W_samples); copy_real(r, "H_samples", H_samples); copy_int(r, "V_samples", V_samples);
        delete r;
        return 0;
    } catch (const std::exception &e) { snprintf(g_err, sizeof g_err, "%s", e.what()); return 1; }
}

int glue_sample_gibbs_cpp(const int *v_n, const double *W, const int *C, int F, int K, double alpha, int iter, double burnin,
                          unsigned long long rng_seed, double *C_mean)
{
    try {
        shim::rng().seed(rng_seed);
        arma::ivec v(F);
        std::copy(v_n, v_n + F, v.memptr());
        SEXP r = sample_gibbs_cpp_cuda(v, to_mat(W, F, K), to_imat(C, F, K), alpha, iter, burnin);
        if (C_mean) std::copy(r->real.begin(), r->real.end(), C_mean);
        delete r;
        return 0;
    } catch (const std::exception &e) { snprintf(g_err, sizeof g_err, "%s", e.what()); return 1; }
}
}
