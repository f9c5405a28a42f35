"""integration/nbmf_glue.cpp -- the bodies a maintainer drops into the reference's Rcpp functions -- compiled
as they stand against the Rcpp/Armadillo stand-in of oracle/ref_shim (tests/glue) and run on the GPU: the
reference-side binding of INTEGRATION.md is exercised, not just described."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "glue", "libnbmf_glue_test.so")
_ip = C.POINTER(C.c_int32); _dp = C.POINTER(C.c_double)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def _lib():
    if not os.path.exists(SO):
        pytest.skip("tests/glue/libnbmf_glue_test.so not built (make -C tests/glue)")
    L = C.CDLL(SO)
    L.glue_last_error.restype = C.c_char_p
    L.glue_vb_aspect.argtypes = [_ip, _ip, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                 _dp, _dp, _dp, _dp, _dp]
    L.glue_vb_dirber.argtypes = [_ip, _ip, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, _dp, _dp, _dp]
    L.glue_gibbs_dirber_finite.argtypes = [_ip, _ip, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int,
                                           C.c_double, C.c_uint64, _ip, C.POINTER(C.c_uint64)]
    return L


def _problem(F, N, K, na, seed):
    V, _, _ = oracle.generate(F, N, 3, 0.4, 0.4, seed)
    V = np.asfortranarray(V)
    if na > 0:
        V[np.random.default_rng(seed).random(V.shape) < na] = oracle.NA
    return V, oracle.random_labels(V, K, seed + 1)


def test_glue_library_exports_the_driver_symbols():
    L = _lib()
    for s in ("glue_vb_aspect", "glue_vb_dirber", "glue_gibbs_dirber_finite", "glue_last_error", "glue_warning_count"):
        assert hasattr(L, s)


@pytest.mark.gpu
def test_glue_vb_aspect_body_matches_oracle():
    L = _lib()
    F, N, K, it = 60, 90, 5, 25
    V, Z = _problem(F, N, K, 0.15, 3)
    E_W = np.zeros((F, K), order="F"); E_H = np.zeros((N, K), order="F"); lb = np.zeros(it); lik = np.zeros(it)
    E_Z = np.zeros((F, K, N), order="F")
    rc = L.glue_vb_aspect(_p(V, _ip), _p(Z, _ip), F, N, K, 0.9, 1.2, 0.4, it, 1, _p(E_W, _dp), _p(E_H, _dp), _p(lb, _dp),
                          _p(lik, _dp), _p(E_Z, _dp))
    assert rc == 0, L.glue_last_error()
    r = oracle.vb_aspect(V, Z, K, 0.9, 1.2, 0.4, iter=it, checklb=True, return_E_Z=True)
    assert np.allclose(E_W, r["E_W"], rtol=1e-10) and np.allclose(E_H, r["E_H"], rtol=1e-10)
    assert np.allclose(lb, r["lowerbounds"], rtol=1e-11) and np.all(np.isnan(lik))          # :389-391
    assert np.max(np.abs(E_Z - r["E_Z"])) < 1e-10
    assert L.glue_warning_count() == 0                                                        # ELBO never decreased


@pytest.mark.gpu
def test_glue_vb_dirber_body_is_bit_identical():
    L = _lib()
    F, N, K, it = 40, 70, 4, 9
    V, _ = _problem(F, N, K, 0.1, 5)
    Z = oracle.random_labels(V, K + 1, 6)
    E_W = np.zeros((F, K + 1), order="F"); E_H = np.zeros((N, K + 1), order="F"); E_Z = np.zeros((F, K + 1, N), order="F")
    rc = L.glue_vb_dirber(_p(V, _ip), _p(Z, _ip), F, N, K, 1.0, 1.0, 0.7, it, _p(E_W, _dp), _p(E_H, _dp), _p(E_Z, _dp))
    assert rc == 0, L.glue_last_error()
    r = oracle.vb_dirber(V, Z, K, 1.0, 1.0, 0.7, iter=it, return_E_Z=True)
    assert np.array_equal(E_W, r["E_W"], equal_nan=True) and np.array_equal(E_H, r["E_H"], equal_nan=True)
    assert np.array_equal(E_Z, r["E_Z"])


@pytest.mark.gpu
def test_glue_gibbs_body_returns_the_reference_cube_and_obeys_set_seed():
    L = _lib()
    F, N, K, it, burnin = 50, 80, 4, 11, 0.5
    V, Z = _problem(F, N, K, 0.1, 7)
    ns = it - int(np.floor(it * burnin))
    Zs = np.zeros((F, N, ns), dtype=np.int32, order="F"); used = C.c_uint64(0)
    n = L.glue_gibbs_dirber_finite(_p(V, _ip), _p(Z, _ip), F, N, K, 1.0, 1.0, 0.5, it, burnin, 1234, _p(Zs, _ip), C.byref(used))
    assert n == ns, L.glue_last_error()
    ref = oracle.gibbs_dirber_finite_keyed(V, Z, K, 1.0, 1.0, 0.5, iter=it, burnin=burnin, seed=used.value)
    obs = V != oracle.NA
    assert np.array_equal(Zs[obs], ref["Z_samples"][obs])
    # the same "set.seed()" gives the same run, another one a different chain
    Zs2 = np.zeros_like(Zs, order="F"); used2 = C.c_uint64(0)
    L.glue_gibbs_dirber_finite(_p(V, _ip), _p(Z, _ip), F, N, K, 1.0, 1.0, 0.5, it, burnin, 1234, _p(Zs2, _ip), C.byref(used2))
    assert used2.value == used.value and np.array_equal(Zs2, Zs)
    L.glue_gibbs_dirber_finite(_p(V, _ip), _p(Z, _ip), F, N, K, 1.0, 1.0, 0.5, it, burnin, 99, _p(Zs2, _ip), C.byref(used2))
    assert used2.value != used.value and (Zs2[obs] != Zs[obs]).mean() > 0.05
