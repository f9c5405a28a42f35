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

// VB_Aspect_Rcpp (collapsed_gibbs_z_VB.cpp:373-526)
SEXP NBMF_GLUE_NAME(VB_Aspect_Rcpp)(const arma::imat &V, arma::imat Z, int K, double alpha, double beta, double gamma,
                                    int iter, bool checklb)
{
    const int F = V.n_rows, N = V.n_cols;
    NbmfHandle g;
    g.ck(nbmf_set_data_i32(g.h, V.memptr(), F, N));
    g.ck(nbmf_init_from_labels(g.h, Z.memptr(), K));              // vb_matrix_to_tensor + compute_vbcounts (:406-412)
    arma::mat E_W(F, K), E_H(N, K);
    arma::rowvec lowerbounds(iter), likelihoods(iter), kldivergences(iter);
    likelihoods.fill(arma::datum::nan); kldivergences.fill(arma::datum::nan);             // :389-391
    const bool want_EZ = (double)F * K * N <= 1.3e8;               // the cube is only materialised when it is small
    arma::cube E_Z;
    if (want_EZ) E_Z.set_size(F, K, N);
    g.ck(nbmf_vb_begin(g.h, K, alpha, beta, gamma, 0));
    for (int i = 0; i < iter; i++) {
        Rcpp::Rcout << "iter: " << i << '/' << iter << std::endl;                          // :416 (progress stays in R)
        g.ck(nbmf_vb_step(g.h, checklb));
    }
    g.ck(nbmf_vb_end(g.h, E_W.memptr(), E_H.memptr(), lowerbounds.memptr(), iter, want_EZ ? E_Z.memptr() : nullptr));
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
