"""ctypes wrapper around oracle/_ref/libnbmf_ref.so: the reference's OWN source files
(/root/reference/src/*.cpp) compiled against the minimal Rcpp/Armadillo/Boost shim in
oracle/ref_shim.  TEST INFRASTRUCTURE ONLY (tests, golden-vector generation, the CPU
baseline of bench.py)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .oracle import NA, _i32, _p, _dp, _ip, HERE

SO = os.path.join(HERE, "_ref", "libnbmf_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(SO)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise FileNotFoundError(SO + " not built (make -C oracle ref; needs /root/reference)")
        L = C.CDLL(SO)
        L.ref_set_seed.argtypes = [C.c_uint64]
        L.ref_vb_aspect.restype = C.c_int
        L.ref_vb_aspect.argtypes = [_ip, _ip, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int,
                                    C.c_int, _dp, _dp, _dp, _dp]
        L.ref_vb_dirber.restype = C.c_int
        L.ref_vb_dirber.argtypes = [_ip, _ip, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int,
                                    _dp, _dp, _dp]
        L.ref_gibbs_dirber_finite.restype = C.c_int
        L.ref_gibbs_dirber_finite.argtypes = [_ip, _ip, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                              C.c_int, C.c_double, C.c_uint64, _ip, _dp]
        L.ref_gibbs_dirber_dp.restype = C.c_int
        L.ref_gibbs_dirber_dp.argtypes = [_ip, _ip, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int,
                                          C.c_double, C.c_uint64, _ip]
        L.ref_gibbs_dirdir.restype = C.c_int
        L.ref_gibbs_dirdir.argtypes = [_ip, _ip, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double,
                                       C.c_uint64, _ip, _ip]
        L.ref_lower_bound_aspect.restype = C.c_double
        L.ref_lower_bound_aspect.argtypes = [_dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, _ip,
                                             C.c_int, C.c_int, C.c_int]
        L.ref_compute_expectations.restype = C.c_int
        L.ref_compute_expectations.argtypes = [_ip, _ip, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                               _dp, _dp]
        L.ref_sample_gibbs_cpp.restype = C.c_int
        L.ref_sample_gibbs_cpp.argtypes = [_ip, _dp, _ip, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double, C.c_uint64, _dp]
        _lib = L
    return _lib


def vb_aspect(V, Z, K, alpha=1.0, beta=1.0, gamma=1.0, iter=100, checklb=False, return_E_Z=False):
    """VB_Aspect_Rcpp (src/collapsed_gibbs_z_VB.cpp:373-526), the reference's own code."""
    V = _i32(V); Z = _i32(Z)
    F, N = V.shape
    E_W = np.zeros((F, K), order="F"); E_H = np.zeros((N, K), order="F"); lbs = np.zeros(max(iter, 1))
    E_Z = np.zeros((F, K, N), order="F") if return_E_Z else None
    rc = lib().ref_vb_aspect(_p(V, _ip), _p(Z, _ip), F, N, K, alpha, beta, gamma, iter, int(checklb), _p(E_W), _p(E_H),
                             _p(lbs), _p(E_Z))
    return {"E_W": E_W, "E_H": E_H, "lowerbounds": lbs[:iter], "E_Z": E_Z, "stopped": rc != 0}


def vb_dirber(V, Z, K, alpha=1.0, beta=1.0, gamma=1.0, iter=100, return_E_Z=False):
    """VB_DirBer_Rcpp (:97-168)."""
    V = _i32(V); Z = _i32(Z)
    F, N = V.shape
    E_W = np.zeros((F, K + 1), order="F"); E_H = np.zeros((N, K + 1), order="F")
    E_Z = np.zeros((F, K + 1, N), order="F") if return_E_Z else None
    rc = lib().ref_vb_dirber(_p(V, _ip), _p(Z, _ip), F, N, K, alpha, beta, gamma, iter, _p(E_W), _p(E_H), _p(E_Z))
    return {"E_W": E_W, "E_H": E_H, "E_Z": E_Z, "stopped": rc != 0}


def gibbs_dirber_finite(V, Z, K, alpha=1.0, beta=1.0, gamma=1.0, iter=100, burnin=0.5, seed=1):
    """Gibbs_DirBer_finite_Rcpp (src/collapsed_gibbs_z.cpp:363-437) on the shim RNG."""
    V = _i32(V); Z = _i32(Z)
    F, N = V.shape
    ns = iter - int(np.floor(iter * burnin))
    Zs = np.zeros((F, N, max(ns, 1)), dtype=np.int32, order="F")
    n = lib().ref_gibbs_dirber_finite(_p(V, _ip), _p(Z, _ip), F, N, K, alpha, beta, gamma, iter, burnin, seed,
                                      _p(Zs, _ip), None)
    assert n == ns, (n, ns)
    return {"Z_samples": Zs[:, :, :ns]}


def gibbs_dirber_dp(V, Z, alpha=1.0, beta=1.0, gamma=1.0, iter=100, burnin=0.5, seed=1):
    """Gibbs_DirBer_DP_Rcpp (src/collapsed_gibbs_z.cpp:275-359) on the shim RNG."""
    V = _i32(V); Z = _i32(Z)
    F, N = V.shape
    ns = iter - int(np.floor(iter * burnin))
    Zs = np.zeros((F, N, max(ns, 1)), dtype=np.int32, order="F")
    n = lib().ref_gibbs_dirber_dp(_p(V, _ip), _p(Z, _ip), F, N, alpha, beta, gamma, iter, burnin, seed, _p(Zs, _ip))
    assert n == ns, (n, ns)
    return {"Z_samples": Zs[:, :, :ns]}


def gibbs_dirdir(V, Z, K, alpha=1.0, gamma=1.0, iter=100, burnin=0.5, seed=1):
    """Gibbs_DirDir_finite_Rcpp (src/collapsed_gibbs_z.cpp:441-557) on the shim RNG."""
    V = _i32(V); Z = _i32(Z)
    F, N = V.shape
    ns = iter - int(np.floor(iter * burnin))
    Zs = np.zeros((F, N, max(ns, 1)), dtype=np.int32, order="F"); Cs = np.zeros_like(Zs, order="F")
    n = lib().ref_gibbs_dirdir(_p(V, _ip), _p(Z, _ip), F, N, K, alpha, gamma, iter, burnin, seed, _p(Zs, _ip), _p(Cs, _ip))
    assert n == ns, (n, ns)
    return {"Z_samples": Zs[:, :, :ns], "C_samples": Cs[:, :, :ns]}


def compute_expectations(Z_samples, V, alpha, beta, gamma):
    """compute_expectation_W_Rcpp / compute_expectation_H_Rcpp (:178-197, :230-272)."""
    Zs = np.asfortranarray(Z_samples, dtype=np.int32); V = _i32(V)
    F, N, J = Zs.shape
    Kq = int(Zs.max()) + 1
    E_W = np.zeros((F, Kq), order="F"); E_H = np.zeros((N, Kq), order="F")
    k = lib().ref_compute_expectations(_p(Zs, _ip), _p(V, _ip), F, N, J, alpha, beta, gamma, _p(E_W), _p(E_H))
    assert k == Kq
    return E_W, E_H


def lower_bound_aspect(E_Z, ElogW, ElogH, Elog1H, gvb, avb, bvb, alpha, beta, gamma, V):
    V = _i32(V)
    F, N = V.shape
    K = ElogW.shape[1]
    f = lambda a: np.asfortranarray(a, dtype=np.float64)
    a = [f(x) for x in (E_Z, ElogW, ElogH, Elog1H, gvb, avb, bvb)]
    return lib().ref_lower_bound_aspect(*[_p(x) for x in a], alpha, beta, gamma, _p(V, _ip), F, K, N)


def sample_gibbs_cpp(v_n, W, Cinit, alpha=1.0, iter=100, burnin=0.5, seed=1):
    v = np.ascontiguousarray(v_n, dtype=np.int32); W = np.asfortranarray(W, dtype=np.float64)
    Ci = np.asfortranarray(Cinit, dtype=np.int32)
    F, K = W.shape
    out = np.zeros((F, K), order="F")
    lib().ref_sample_gibbs_cpp(_p(v, _ip), _p(W), _p(Ci, _ip), F, K, alpha, iter, burnin, seed, _p(out))
    return out
