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
    L.glue_lower_bound_aspect.argtypes = [_dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, _ip, C.c_int,
                                          C.c_int, C.c_int, _dp]
    L.glue_compute_expectations.argtypes = [_ip, _ip, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, _dp, _dp]
    L.glue_samples_VWH.argtypes = [_ip, _ip, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_uint64,
                                   _dp, _dp, _ip]
    L.glue_sample_gibbs_cpp.argtypes = [_ip, _dp, _ip, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double, C.c_uint64, _dp]
    return L


def _problem(F, N, K, na, seed):
    V, _, _ = oracle.generate(F, N, 3, 0.4, 0.4, seed)
    V = np.asfortranarray(V)
    if na > 0:
        V[np.random.default_rng(seed).random(V.shape) < na] = oracle.NA
    return V, oracle.random_labels(V, K, seed + 1)


def test_glue_library_exports_the_driver_symbols():
    L = _lib()
    for s in ("glue_vb_aspect", "glue_vb_dirber", "glue_gibbs_dirber_finite", "glue_last_error", "glue_warning_count",
              "glue_lower_bound_aspect", "glue_compute_expectations", "glue_samples_VWH", "glue_sample_gibbs_cpp"):
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


@pytest.mark.gpu
def test_glue_remaining_registered_symbols():
    """lower_bound_aspect, compute_expectation_{W,H}_Rcpp, samples_VWH_DirBer and sample_gibbs_cpp
    (src/RcppExports.cpp:316-339) through the bodies of integration/nbmf_glue.cpp."""
    import scipy.special as sp
    from oracle import ref
    L = _lib()
    F, N, K = 30, 44, 4
    V, Z = _problem(F, N, K, 0.1, 11)
    alpha, beta, gamma = 0.9, 1.2, 0.6
    # -- lower_bound_aspect with the parameters implied by an E_Z of the oracle
    r = oracle.vb_aspect(V, Z, K, alpha, beta, gamma, iter=4, checklb=True, return_E_Z=True)
    E_Z = np.asfortranarray(r["E_Z"]); obs = V != oracle.NA
    nt = np.einsum("fkn,fn->fk", E_Z, obs.astype(float))
    fo = np.einsum("fkn,fn->nk", E_Z, (obs & (V == 1)).astype(float)); fz = np.einsum("fkn,fn->nk", E_Z, (obs & (V == 0)).astype(float))
    g, a, b = (np.asfortranarray(x) for x in (gamma + nt, alpha + fo, beta + fz))
    ElW = np.asfortranarray(sp.digamma(g) - sp.digamma(g.sum(1, keepdims=True)))
    ElH = np.asfortranarray(sp.digamma(a) - sp.digamma(a + b)); El1H = np.asfortranarray(sp.digamma(b) - sp.digamma(a + b))
    lb = C.c_double(0.0)
    rc = L.glue_lower_bound_aspect(_p(E_Z, _dp), _p(ElW, _dp), _p(ElH, _dp), _p(El1H, _dp), _p(g, _dp), _p(a, _dp), _p(b, _dp),
                                   alpha, beta, gamma, _p(V, _ip), F, N, K, C.byref(lb))
    assert rc == 0, L.glue_last_error()
    olb, _ = oracle.lower_bound_aspect(E_Z, ElW, ElH, El1H, g, a, b, alpha, beta, gamma, V)
    assert abs(lb.value - olb) <= 1e-10 * abs(olb)
    # -- compute_expectation_{W,H}_Rcpp on a sample cube
    J = 5
    Zs = np.asfortranarray(np.random.default_rng(2).integers(0, K, size=(F, N, J)).astype(np.int32))
    Zs[np.repeat((V == oracle.NA)[:, :, None], J, axis=2)] = oracle.NA
    EW = np.zeros((F, K), order="F"); EH = np.zeros((N, K), order="F")
    kq = L.glue_compute_expectations(_p(Zs, _ip), _p(V, _ip), F, N, J, alpha, beta, gamma, _p(EW, _dp), _p(EH, _dp))
    assert kq == K, L.glue_last_error()
    assert np.allclose(EW, oracle.compute_expectation_W(Zs, gamma), rtol=1e-14)
    assert np.array_equal(EH, oracle.compute_expectation_H(Zs, V, alpha, beta))
    if ref.available():
        rW, rH = ref.compute_expectations(Zs, V, alpha, beta, gamma)
        assert np.allclose(EW, rW, rtol=1e-14) and np.allclose(EH, rH, rtol=1e-14)
    # -- samples_VWH_DirBer: shapes, simplex rows, H in (0,1), binary V, same "set.seed()" same draws
    Ws = np.zeros((F, K, J), order="F"); Hs = np.zeros((N, K, J), order="F"); Vs = np.zeros((F, N, J), dtype=np.int32, order="F")
    assert L.glue_samples_VWH(_p(V, _ip), _p(Zs, _ip), F, N, J, K, gamma, alpha, beta, 77, _p(Ws, _dp), _p(Hs, _dp), _p(Vs, _ip)) == 0
    assert np.allclose(Ws.sum(1), 1.0, atol=1e-12) and np.all((Hs > 0) & (Hs < 1)) and set(np.unique(Vs)) <= {0, 1}
    Ws2 = np.zeros_like(Ws, order="F")
    L.glue_samples_VWH(_p(V, _ip), _p(Zs, _ip), F, N, J, K, gamma, alpha, beta, 77, _p(Ws2, _dp), None, None)
    assert np.array_equal(Ws2, Ws)
    # -- sample_gibbs_cpp as the reference is written
    rng = np.random.default_rng(4)
    W = np.asfortranarray(rng.dirichlet(np.ones(K), size=F)); v = rng.integers(0, 2, size=F).astype(np.int32)
    C0 = np.zeros((F, K), dtype=np.int32, order="F"); C0[np.arange(F), rng.integers(0, K, size=F)] = 1
    Cm = np.zeros((F, K), order="F")
    assert L.glue_sample_gibbs_cpp(_p(v, _ip), _p(W, _dp), _p(C0, _ip), F, K, 1.0, 30, 0.5, 5, _p(Cm, _dp)) == 0
    kept = 1 + sum(1 for i in range(1, 30) if i >= 15)
    assert np.allclose(Cm, C0 / kept, rtol=1e-14)
    if ref.available():
        assert np.allclose(Cm, ref.sample_gibbs_cpp(v, W, C0, 1.0, 30, 0.5, seed=5), rtol=1e-14, atol=1e-15)


@pytest.mark.gpu
def test_glue_vb_aspect_body_over_two_devices(monkeypatch):
    """NBMF_DEVICES=0,1: the same Rcpp body shards the columns of V over two GPUs inside the library (device group,
    NCCL communicator owned by libnbmf_cuda) and returns what one GPU returns."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    L = _lib()
    F, N, K, it = 64, 256, 5, 8
    V, Z = _problem(F, N, K, 0.1, 13)
    out = []
    for devs in ("0", "0,1"):
        monkeypatch.setenv("NBMF_DEVICES", devs)
        E_W = np.zeros((F, K), order="F"); E_H = np.zeros((N, K), order="F"); lb = np.zeros(it); lik = np.zeros(it)
        rc = L.glue_vb_aspect(_p(V, _ip), _p(Z, _ip), F, N, K, 1.0, 1.0, 0.8, it, 1, _p(E_W, _dp), _p(E_H, _dp), _p(lb, _dp),
                              _p(lik, _dp), None)
        assert rc == 0, L.glue_last_error()
        out.append((E_W, E_H, lb))
    for a, b in zip(out[0], out[1]):
        assert np.allclose(a, b, rtol=1e-11, atol=1e-15)
    r = oracle.vb_aspect(V, Z, K, 1.0, 1.0, 0.8, iter=it, checklb=True)
    assert np.allclose(out[1][0], r["E_W"], rtol=1e-9) and np.allclose(out[1][2], r["lowerbounds"], rtol=1e-10)
