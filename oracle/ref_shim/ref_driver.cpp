// ref_driver.cpp -- extern "C" entry points into the reference's own functions (compiled
// from /root/reference/src/*.cpp against the shim).  Test infrastructure only.
#include <RcppArmadillo.h>

// the reference's exported functions (signatures as in src/RcppExports.cpp)
SEXP VB_Aspect_Rcpp(const arma::imat &V, arma::imat Z, int K, double alpha, double beta, double gamma, int iter, bool checklb);
SEXP VB_DirBer_Rcpp(const arma::imat &V, arma::imat Z, int K, double alpha, double beta, double gamma, int iter);
SEXP Gibbs_DirBer_finite_Rcpp(const arma::imat &V, arma::imat Z, int K, double alpha, double beta, double gamma, int iter, double burnin);
SEXP Gibbs_DirBer_DP_Rcpp(const arma::imat &V, arma::imat Z, double alpha, double beta, double gamma, int iter, double burnin);
SEXP Gibbs_DirDir_finite_Rcpp(const arma::imat &V, arma::imat Z, int K, double alpha, double gamma, int iter, double burnin);
double lower_bound_aspect(const arma::cube &E_Z, const arma::mat &E_log_W, const arma::mat &E_log_H, const arma::mat &E_log_1_H,
                          const arma::mat &gamma_vb, const arma::mat &alpha_vb, const arma::mat &beta_vb, double alpha, double beta,
                          double gamma, const arma::imat &V);
arma::mat compute_expectation_W_Rcpp(const arma::icube &Z_samples, double gamma);
arma::mat compute_expectation_H_Rcpp(const arma::icube &Z_samples, const arma::imat &V, double alpha, double beta);
SEXP sample_gibbs_cpp(const arma::ivec &v_n, const arma::mat &W, arma::imat C, double alpha, int iter, double burnin);
SEXP matrix_to_tensor_R(const arma::imat &Z, int Kmax);

static arma::imat to_imat(const int *p, int r, int c) { arma::imat m(r, c); std::copy(p, p + (size_t)r * c, m.mem); return m; }
static void copy_real(SEXP l, const char *name, double *dst)
{
    shim::Obj *o = l->get(name);
    if (o && dst) std::copy(o->real.begin(), o->real.end(), dst);
}

extern "C" {
void ref_set_seed(unsigned long long s) { shim::rng().seed(s); }
int ref_warning_count() { return shim::warning_count(); }

int ref_vb_aspect(const int *V, const int *Z, int F, int N, int K, double alpha, double beta, double gamma, int iter, int checklb,
                  double *E_W, double *E_H, double *lowerbounds, double *E_Z)
{
    try {
        SEXP r = VB_Aspect_Rcpp(to_imat(V, F, N), to_imat(Z, F, N), K, alpha, beta, gamma, iter, checklb != 0);
        copy_real(r, "E_W", E_W); copy_real(r, "E_H", E_H); copy_real(r, "lowerbounds", lowerbounds); copy_real(r, "E_Z", E_Z);
        delete r;
        return 0;
    } catch (const std::exception &) { return 1; }
}

int ref_vb_dirber(const int *V, const int *Z, int F, int N, int K, double alpha, double beta, double gamma, int iter, double *E_W,
                  double *E_H, double *E_Z)
{
    try {
        SEXP r = VB_DirBer_Rcpp(to_imat(V, F, N), to_imat(Z, F, N), K, alpha, beta, gamma, iter);
        copy_real(r, "E_W", E_W); copy_real(r, "E_H", E_H); copy_real(r, "E_Z", E_Z);
        delete r;
        return 0;
    } catch (const std::exception &) { return 1; }
}

// Z_samples: F x N x (iter - floor(iter*burnin)); returns the number of slices
int ref_gibbs_dirber_finite(const int *V, const int *Z, int F, int N, int K, double alpha, double beta, double gamma, int iter,
                            double burnin, unsigned long long seed, int *Z_samples, double *E_Z)
{
    try {
        shim::rng().seed(seed);
        SEXP r = Gibbs_DirBer_finite_Rcpp(to_imat(V, F, N), to_imat(Z, F, N), K, alpha, beta, gamma, iter, burnin);
        shim::Obj *zs = r->get("Z_samples");
        int ns = zs->dim.size() == 3 ? (int)zs->dim[2] : 0;
        if (Z_samples) std::copy(zs->integer.begin(), zs->integer.end(), Z_samples);
        copy_real(r, "E_Z", E_Z);
        delete r;
        return ns;
    } catch (const std::exception &) { return -1; }
}

// the sibling samplers (collapsed_gibbs_z.cpp:275-359, :441-557); sample cubes F x N x nsamples
int ref_gibbs_dirber_dp(const int *V, const int *Z, int F, int N, double alpha, double beta, double gamma, int iter, double burnin,
                        unsigned long long seed, int *Z_samples)
{
    try {
        shim::rng().seed(seed);
        SEXP r = Gibbs_DirBer_DP_Rcpp(to_imat(V, F, N), to_imat(Z, F, N), alpha, beta, gamma, iter, burnin);
        shim::Obj *zs = r->get("Z_samples");
        int ns = zs->dim.size() == 3 ? (int)zs->dim[2] : 0;
        if (Z_samples) std::copy(zs->integer.begin(), zs->integer.end(), Z_samples);
        delete r;
        return ns;
    } catch (const std::exception &) { return -1; }
}
int ref_gibbs_dirdir(const int *V, const int *Z, int F, int N, int K, double alpha, double gamma, int iter, double burnin,
                     unsigned long long seed, int *Z_samples, int *C_samples)
{
    try {
        shim::rng().seed(seed);
        SEXP r = Gibbs_DirDir_finite_Rcpp(to_imat(V, F, N), to_imat(Z, F, N), K, alpha, gamma, iter, burnin);
        shim::Obj *zs = r->get("Z_samples"), *cs = r->get("C_samples");
        int ns = zs->dim.size() == 3 ? (int)zs->dim[2] : 0;
        if (Z_samples) std::copy(zs->integer.begin(), zs->integer.end(), Z_samples);
        if (C_samples) std::copy(cs->integer.begin(), cs->integer.end(), C_samples);
        delete r;
        return ns;
    } catch (const std::exception &) { return -1; }
}

double ref_lower_bound_aspect(const double *E_Z, const double *ElogW, const double *ElogH, const double *Elog1H, const double *gvb,
                              const double *avb, const double *bvb, double alpha, double beta, double gamma, const int *V, int F,
                              int K, int N)
{
    arma::cube ez(F, K, N); std::copy(E_Z, E_Z + (size_t)F * K * N, ez.d.begin());
    auto M = [](const double *p, int r, int c) { arma::mat m(r, c); std::copy(p, p + (size_t)r * c, m.mem); return m; };
    return lower_bound_aspect(ez, M(ElogW, F, K), M(ElogH, N, K), M(Elog1H, N, K), M(gvb, F, K), M(avb, N, K), M(bvb, N, K), alpha,
                              beta, gamma, to_imat(V, F, N));
}

// E_W: F x (max label + 1), E_H: N x (max label + 1); returns max label + 1
int ref_compute_expectations(const int *Z_samples, const int *V, int F, int N, int J, double alpha, double beta, double gamma,
                             double *E_W, double *E_H)
{
    arma::icube zs(F, N, J); std::copy(Z_samples, Z_samples + (size_t)F * N * J, zs.d.begin());
    arma::mat w = compute_expectation_W_Rcpp(zs, gamma);
    arma::mat h = compute_expectation_H_Rcpp(zs, to_imat(V, F, N), alpha, beta);
    if (E_W) std::copy(w.mem, w.mem + w.n_elem, E_W);
    if (E_H) std::copy(h.mem, h.mem + h.n_elem, E_H);
    return (int)w.n_cols;
}

int ref_sample_gibbs_cpp(const int *v_n, const double *W, const int *C, int F, int K, double alpha, int iter, double burnin,
                         unsigned long long seed, double *C_mean)
{
    shim::rng().seed(seed);
    arma::ivec v(F); std::copy(v_n, v_n + F, v.d.begin());
    arma::mat w(F, K); std::copy(W, W + (size_t)F * K, w.mem);
    SEXP r = sample_gibbs_cpp(v, w, to_imat(C, F, K), alpha, iter, burnin);
    std::copy(r->real.begin(), r->real.end(), C_mean);
    delete r;
    return 0;
}
}
