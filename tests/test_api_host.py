"""The posterior-mean helpers the R wrappers apply to the sample cubes (compute_expectation_*_Rcpp), as the
package exposes them under the reference's names -- device reducers since round 2 -- checked against the
oracle's restatements."""
import numpy as np
import pytest

from nbmf_b200.api import (compute_expectation_H_dirdir_Rcpp, compute_expectation_H_Rcpp, compute_expectation_W_Rcpp)
from oracle import oracle

NA = oracle.NA
pytestmark = pytest.mark.gpu


def _samples(F=14, N=19, K=4, J=7, na=0.2, seed=3):
    rng = np.random.default_rng(seed)
    V = rng.integers(0, 2, size=(F, N)).astype(np.int32)
    V[rng.random((F, N)) < na] = NA
    Zs = rng.integers(0, K, size=(F, N, J)).astype(np.int32)
    Zs[V == NA] = NA
    Zs[:, :, J - 1] = 0            # an unwritten slice (burnin = 0): all zeros, also where V is NA (Appendix B.10)
    return np.asfortranarray(V), np.asfortranarray(Zs)


def test_expectation_W_and_H_match_the_oracle_restatements():
    V, Zs = _samples()
    F, N, J = Zs.shape
    Kq = int(Zs.max()) + 1
    L = oracle.lib()
    EW = np.zeros((F, Kq), order="F"); EH = np.zeros((N, Kq), order="F")
    L.oracle_compute_expectation_W(oracle._p(Zs, oracle._ip), F, N, J, 0.7, Kq, oracle._p(EW))
    L.oracle_compute_expectation_H(oracle._p(Zs, oracle._ip), oracle._p(V, oracle._ip), F, N, J, 1.3, 0.6, Kq, oracle._p(EH))
    W = compute_expectation_W_Rcpp(Zs, 0.7)
    H = compute_expectation_H_Rcpp(Zs, V, 1.3, 0.6)
    assert W.shape == (F, Kq) and np.allclose(W, EW, rtol=1e-13, atol=0) and np.allclose(W.sum(1), 1.0)
    assert H.shape == (N, Kq) and np.allclose(H, EH, rtol=1e-13, atol=0)
    assert np.all(H[:, Kq - 1] == 0)          # the reference's k < max(label) loop (:236-242)


def test_expectation_H_dirdir_is_the_normalised_column_count():
    _, Cs = _samples(na=0.0)
    F, N, J = Cs.shape
    K = int(Cs.max()) + 1
    ref = np.zeros((N, K))
    for n in range(N):
        for k in range(K):
            ref[n, k] = 0.9 + (Cs[:, n, :] == k).sum() / J
        ref[n] /= ref[n].sum()
    assert np.allclose(compute_expectation_H_dirdir_Rcpp(Cs, 0.9), ref, rtol=1e-13)
