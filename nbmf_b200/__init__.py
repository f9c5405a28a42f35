"""nbmf_b200 -- B200-native engine for the VB / Gibbs hot path of alumbreras/NBMF.

The package is a thin host-side mirror of the reference's operator interface
(nbmf_b200.api) over the C ABI of libnbmf_cuda (include/nbmf_cuda.h, built from
nbmf_b200/csrc).  All arithmetic of the path runs in hand-written sm_100a CUDA
kernels; there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .api import (  # noqa: F401
    BernoulliNMF, Engine, Group, Gibbs_DirBer, Gibbs_DirBer_DP, Gibbs_DirBer_DP_Rcpp, Gibbs_DirBer_finite_Rcpp, Gibbs_DirDir,
    Gibbs_DirDir_finite_Rcpp,
    NA_INTEGER, NbmfError, VB_Aspect_Rcpp,
    VB_DirBer, VB_DirBer_Rcpp, VB_aspectRcpp, as_imat, expectation, loglikelihood, lower_bound_aspect,
)

__all__ = [
    "BernoulliNMF", "Engine", "Group", "Gibbs_DirBer", "Gibbs_DirBer_DP", "Gibbs_DirBer_DP_Rcpp", "Gibbs_DirBer_finite_Rcpp",
    "Gibbs_DirDir", "Gibbs_DirDir_finite_Rcpp", "NA_INTEGER", "NbmfError",
    "VB_Aspect_Rcpp", "VB_DirBer", "VB_DirBer_Rcpp", "VB_aspectRcpp", "as_imat", "expectation",
    "loglikelihood", "lower_bound_aspect",
]
