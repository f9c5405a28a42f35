// glue_driver.cpp -- extern "C" access to integration/nbmf_glue.cpp compiled against oracle/ref_shim
// (test infrastructure: the shim stands in for Rcpp/Armadillo, which the image does not have).
#include <RcppArmadillo.h>

SEXP VB_Aspect_Rcpp_cuda(const arma::imat &V, arma::imat Z, int K, double alpha, double beta, double gamma, int iter, bool checklb);
SEXP VB_DirBer_Rcpp_cuda(const arma::imat &V, arma::imat Z, int K, double alpha, double beta, double gamma, int iter);
SEXP Gibbs_DirBer_finite_Rcpp_cuda(const arma::imat &V, arma::imat Z, int K, double alpha, double beta, double gamma, int iter, double burnin);
extern unsigned long long nbmf_glue_last_seed;

static arma::imat to_imat(const int *p, int r, int c) { arma::imat m(r, c); std::copy(p, p + (size_t)r * c, m.mem); return m; }
static void copy_real(SEXP l, const char *name, double *dst)
{
    shim::Obj *o = l->get(name);
    if (o && dst) std::copy(o->real.begin(), o->real.end(), dst);
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
}
