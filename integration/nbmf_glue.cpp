// nbmf_glue.cpp -- bodies for the reference's exported hot-path functions that marshal to the C ABI of
// include/nbmf_cuda.h.  In the package they REPLACE the bodies of the same-named functions in
// src/collapsed_gibbs_z_VB.cpp / src/collapsed_gibbs_z.cpp; signatures, R stubs, RcppExports and NAMESPACE do
// not change.  arma::imat is column-major int32 with NA_INTEGER == INT_MIN, which is what the
// library takes, so V.memptr() / Z.memptr() are passed without a copy.
//
// This file is also compiled by tests/glue against the minimal Rcpp/Armadillo stand-in of oracle/ref_shim (R is
// not in the build image) with NBMF_GLUE_SUFFIX defined, so that the bodies are exercised on a GPU as they stand.
#include <RcppArmadillo.h>
#include "nbmf_glue.h"

#ifdef NBMF_GLUE_SUFFIX
#define NBMF_GLUE_NAME(x) x##_cuda
#else
#define NBMF_GLUE_NAME(x) x
#endif

unsigned long long nbmf_glue_last_seed = 0;   // the seed the last Gibbs call derived from R's RNG

// VB_Aspect_Rcpp (collapsed_gibbs_z_VB.cpp:373-526).  One device: a handle; several devices (NBMF_DEVICES): a
// device group, the columns of V sharded inside the library -- the R side cannot tell the difference.
SEXP NBMF_GLUE_NAME(VB_Aspect_Rcpp)(const arma::imat &V, arma::imat Z, int K, double alpha, double beta, double gamma,
                                    int iter, bool checklb)
{
    const int F = V.n_rows, N = V.n_cols;
    arma::mat E_W(F, K), E_H(N, K);
    arma::rowvec lowerbounds(iter), likelihoods(iter), kldivergences(iter);
    likelihoods.fill(arma::datum::nan); kldivergences.fill(arma::datum::nan);             // :389-391
    const std::vector<int> devs = nbmf_glue_devices();
    const bool grouped = devs.size() > 1 && N >= 64 * (int)devs.size();
    const bool want_EZ = !grouped && (double)F * K * N <= 1.3e8;   // the cube is only materialised when it is small
    arma::cube E_Z;
    if (want_EZ) E_Z.set_size(F, K, N);
    if (grouped) {
        NbmfGroup g(devs, NBMF_PREC_AUTO, iter);
        g.ck(nbmf_group_set_data_i32(g.g, V.memptr(), F, N));
        g.ck(nbmf_group_init_from_labels(g.g, Z.memptr(), K));
        g.ck(nbmf_group_vb_begin(g.g, K, alpha, beta, gamma, 0));
        for (int i = 0; i < iter; i++) {
            Rcpp::Rcout << "iter: " << i << '/' << iter << std::endl;
            g.ck(nbmf_group_vb_step(g.g, checklb));
        }
        g.ck(nbmf_group_vb_end(g.g, E_W.memptr(), E_H.memptr(), lowerbounds.memptr(), iter));
    } else {
        NbmfHandle g(NBMF_PREC_AUTO, iter);                           // AUTO precision plans for this call's iter
        g.ck(nbmf_set_data_i32(g.h, V.memptr(), F, N));
        g.ck(nbmf_init_from_labels(g.h, Z.memptr(), K));              // vb_matrix_to_tensor + compute_vbcounts (:406-412)
        g.ck(nbmf_vb_begin(g.h, K, alpha, beta, gamma, 0));
        for (int i = 0; i < iter; i++) {
            Rcpp::Rcout << "iter: " << i << '/' << iter << std::endl;                      // :416 (progress stays in R)
            g.ck(nbmf_vb_step(g.h, checklb));
        }
        g.ck(nbmf_vb_end(g.h, E_W.memptr(), E_H.memptr(), lowerbounds.memptr(), iter, want_EZ ? E_Z.memptr() : nullptr));
    }
    if (!checklb) lowerbounds.zeros();                                                     // :474-477: never written
    for (int i = 2; i < iter; i++)                                                         // :487-494
        if (checklb && lowerbounds[i] < (float)lowerbounds[i - 1]) Rcpp::warning("Decrease detected in likelihood. Something is wrong");
    return Rcpp::List::create(Rcpp::Named("E_Z") = want_EZ ? Rcpp::wrap(E_Z) : R_NilValue,
                              Rcpp::Named("E_W") = Rcpp::wrap(E_W), Rcpp::Named("E_H") = Rcpp::wrap(E_H),
                              Rcpp::Named("lowerbounds") = Rcpp::wrap(lowerbounds),
                              Rcpp::Named("likelihoods") = Rcpp::wrap(likelihoods),
                              Rcpp::Named("kldivergences") = Rcpp::wrap(kldivergences));
}

// VB_DirBer_Rcpp (collapsed_gibbs_z_VB.cpp:97-168)
SEXP NBMF_GLUE_NAME(VB_DirBer_Rcpp)(const arma::imat &V, arma::imat Z, int K, double alpha, double beta, double gamma, int iter)
{
    const int F = V.n_rows, N = V.n_cols, Kmax = K + 1;                                    // :101
    NbmfHandle g(NBMF_PREC_FP64);
    g.ck(nbmf_set_data_i32(g.h, V.memptr(), F, N));
    g.ck(nbmf_init_from_labels(g.h, Z.memptr(), Kmax));
    arma::mat E_W(F, Kmax), E_H(N, Kmax);
    const bool exact = (double)F * Kmax * N * 8 <= 6e10;          // E_Z must exist for the reference's own schedule
    arma::cube E_Z;
    if (exact) E_Z.set_size(F, Kmax, N);
    g.ck(nbmf_vb_dirber(g.h, K, alpha, beta, gamma, iter, exact ? NBMF_DIRBER_EXACT_WAVEFRONT : NBMF_DIRBER_CONTRACTION,
                        E_W.memptr(), E_H.memptr(), exact ? E_Z.memptr() : nullptr));
    return Rcpp::List::create(Rcpp::Named("E_Z") = exact ? Rcpp::wrap(E_Z) : R_NilValue,
                              Rcpp::Named("E_W") = Rcpp::wrap(E_W), Rcpp::Named("E_H") = Rcpp::wrap(E_H));
}

// Gibbs_DirBer_finite_Rcpp (collapsed_gibbs_z.cpp:363-437), the reference's own (collapsed) chain; the seed comes
// from R's RNG under the existing RNGScope (RcppExports.cpp), so set.seed() still governs the run
SEXP NBMF_GLUE_NAME(Gibbs_DirBer_finite_Rcpp)(const arma::imat &V, arma::imat Z, int K, double alpha, double beta,
                                              double gamma, int iter, double burnin)
{
    const int F = V.n_rows, N = V.n_cols;
    const int nsamples = iter - (int)std::floor(iter * burnin);                            // :378
    NbmfHandle g(NBMF_PREC_FP64);
    g.ck(nbmf_set_data_i32(g.h, V.memptr(), F, N));
    g.ck(nbmf_init_from_labels(g.h, Z.memptr(), K));                                      // compute_counts (:388-389)
    const unsigned long long seed = (unsigned long long)(Rcpp::as<double>(Rcpp::runif(1)) * 9007199254740992.0);
    nbmf_glue_last_seed = seed;
    arma::icube Z_samples(F, N, nsamples > 0 ? nsamples : 0);
    int written = 0;
    g.ck(nbmf_gibbs_dirber_collapsed(g.h, K, alpha, beta, gamma, iter, burnin, seed,
                                     nsamples > 0 ? Z_samples.memptr() : nullptr, &written, nullptr));
#ifdef NBMF_GLUE_SUFFIX
    return Rcpp::List::create(Rcpp::Named("Z_samples") = Rcpp::wrap(Z_samples));          // test build: no reference objects linked
#else
    arma::cube E_Z = compute_expectation_Z(Z_samples);                                     // :432, stays in the package
    return Rcpp::List::create(Rcpp::Named("Z_samples") = Rcpp::wrap(Z_samples), Rcpp::Named("E_Z") = Rcpp::wrap(E_Z));
#endif
}

// lower_bound_aspect (collapsed_gibbs_z_VB.cpp:339-369), the exported eleven-argument form
double NBMF_GLUE_NAME(lower_bound_aspect)(const arma::cube &E_Z, const arma::mat &E_log_W, const arma::mat &E_log_H,
                                          const arma::mat &E_log_1_H, const arma::mat &gamma_vb, const arma::mat &alpha_vb,
                                          const arma::mat &beta_vb, double alpha, double beta, double gamma, const arma::imat &V)
{
    const int F = V.n_rows, N = V.n_cols, K = E_log_W.n_cols;
    NbmfHandle g(NBMF_PREC_FP64);
    double lb = 0.0;
    g.ck(nbmf_lower_bound_aspect_full(g.h, E_Z.memptr(), E_log_W.memptr(), E_log_H.memptr(), E_log_1_H.memptr(),
                                      gamma_vb.memptr(), alpha_vb.memptr(), beta_vb.memptr(), alpha, beta, gamma,
                                      V.memptr(), F, N, K, &lb, nullptr));
    return lb;
}

// compute_expectation_W_Rcpp (collapsed_gibbs_z.cpp:178-197)
arma::mat NBMF_GLUE_NAME(compute_expectation_W_Rcpp)(const arma::icube &Z_samples, double gamma)
{
    const int F = Z_samples.n_rows, N = Z_samples.n_cols, J = Z_samples.n_slices, Kq = Z_samples.max() + 1;   // :180
    NbmfHandle g(NBMF_PREC_FP64);
    arma::mat E_W(F, Kq);
    g.ck(nbmf_compute_expectation_W(g.h, Z_samples.memptr(), F, N, J, gamma, Kq, E_W.memptr()));
    return E_W;
}

// compute_expectation_H_Rcpp (collapsed_gibbs_z.cpp:230-272)
arma::mat NBMF_GLUE_NAME(compute_expectation_H_Rcpp)(const arma::icube &Z_samples, const arma::imat &V, double alpha, double beta)
{
    const int F = Z_samples.n_rows, N = Z_samples.n_cols, J = Z_samples.n_slices, Kq = Z_samples.max() + 1;   // :236
    NbmfHandle g(NBMF_PREC_FP64);
    arma::mat E_H(N, Kq);
    g.ck(nbmf_compute_expectation_H(g.h, Z_samples.memptr(), V.memptr(), F, N, J, alpha, beta, Kq, E_H.memptr()));
    return E_H;
}

// samples_VWH_DirBer (collapsed_gibbs_z.cpp:565-625); the seed comes from R's RNG, as for the Gibbs entry
SEXP NBMF_GLUE_NAME(samples_VWH_DirBer)(const arma::imat &V, const arma::icube &Z_samples, double gamma, double alpha, double beta)
{
    const int F = Z_samples.n_rows, N = Z_samples.n_cols, J = Z_samples.n_slices, K = Z_samples.max() + 1;    // :573
    NbmfHandle g(NBMF_PREC_FP64);
    const unsigned long long seed = (unsigned long long)(Rcpp::as<double>(Rcpp::runif(1)) * 9007199254740992.0);
    nbmf_glue_last_seed = seed;
    arma::cube W_samples(F, K, J), H_samples(N, K, J);
    arma::icube V_samples(F, N, J);
    g.ck(nbmf_samples_VWH_DirBer(g.h, V.memptr(), Z_samples.memptr(), F, N, J, K, gamma, alpha, beta, seed,
                                 W_samples.memptr(), H_samples.memptr(), V_samples.memptr()));
    return Rcpp::List::create(Rcpp::Named("W_samples") = Rcpp::wrap(W_samples), Rcpp::Named("H_samples") = Rcpp::wrap(H_samples),
                              Rcpp::Named("V_samples") = Rcpp::wrap(V_samples));
}

// sample_gibbs_cpp (sample_gibbs.cpp:22-80), as the reference is written (its draw is never stored, :60-69)
SEXP NBMF_GLUE_NAME(sample_gibbs_cpp)(const arma::ivec &v_n, const arma::mat &W, arma::imat C, double alpha, int iter, double burnin)
{
    const int F = W.n_rows, K = W.n_cols;
    NbmfHandle g(NBMF_PREC_FP64);
    const unsigned long long seed = (unsigned long long)(Rcpp::as<double>(Rcpp::runif(1)) * 9007199254740992.0);
    nbmf_glue_last_seed = seed;
    arma::mat C_samples_mean(F, K);
    g.ck(nbmf_sample_gibbs_cpp(g.h, v_n.memptr(), W.memptr(), C.memptr(), F, K, alpha, iter, burnin, seed, 1, C_samples_mean.memptr()));
    return Rcpp::wrap(C_samples_mean);
}
