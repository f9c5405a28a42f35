// nbmf_glue.h -- what the reference's src/*.cpp need to route the hot path through libnbmf_cuda
// (INTEGRATION.md).  Header-only; include after <RcppArmadillo.h>.
#pragma once
#include "nbmf_cuda.h"

struct NbmfHandle {                       // RAII so that Rcpp::stop() cannot leak the handle
    nbmf_handle *h = nullptr;
    explicit NbmfHandle(int precision = NBMF_PREC_AUTO)
    {
        nbmf_opts o = {};
        o.device = 0; o.precision = precision;
        if (nbmf_create(&h, &o) != NBMF_OK) Rcpp::stop("libnbmf_cuda: no CUDA device (there is no CPU fallback)");
    }
    ~NbmfHandle() { nbmf_destroy(h); }
    NbmfHandle(const NbmfHandle &) = delete;
    NbmfHandle &operator=(const NbmfHandle &) = delete;
    void ck(int rc) { if (rc != NBMF_OK) Rcpp::stop(nbmf_last_error(h)); }   // status code -> R error (BEGIN_RCPP / END_RCPP)
};
