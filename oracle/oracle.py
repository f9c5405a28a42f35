"""ctypes wrapper around the CPU oracle (oracle/nbmf_oracle.c) and, when it has
been built, around the reference's own sources compiled against the shim
(oracle/_ref/libnbmf_ref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Nothing under nbmf_b200/
imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NA = np.iinfo(np.int32).min

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def build(force: bool = False) -> str:
    so = os.path.join(HERE, "libnbmf_oracle.so")
    src = os.path.join(HERE, "nbmf_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "libnbmf_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def build_ref() -> str | None:
    """Compile the reference's own .cpp files against oracle/ref_shim (only
    possible where /root/reference exists).  Returns the .so path or None."""
    so = os.path.join(HERE, "_ref", "libnbmf_ref.so")
    if os.path.isdir("/root/reference/src") and os.path.exists(os.path.join(HERE, "ref_shim", "ref_driver.cpp")):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)
    return so if os.path.exists(so) else None


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        L.oracle_digamma.restype = C.c_double
        L.oracle_digamma.argtypes = [C.c_double]
        L.oracle_lgamma.restype = C.c_double
        L.oracle_lgamma.argtypes = [C.c_double]
        L.oracle_vb_aspect.restype = C.c_int
        L.oracle_vb_aspect.argtypes = [_ip, _ip, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                       C.c_double, C.c_int, C.c_int, _dp, _dp, _dp, _dp]
        L.oracle_vb_aspect_batch.restype = C.c_int
        L.oracle_vb_aspect_batch.argtypes = [_ip, _ip, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                             C.c_double, C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_int]
        L.oracle_vb_dirber.restype = C.c_int
        L.oracle_vb_dirber.argtypes = [_ip, _ip, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                       C.c_double, C.c_int, _dp, _dp, _dp]
        L.oracle_vb_dirber_wavefront.restype = C.c_int
        L.oracle_vb_dirber_wavefront.argtypes = [_ip, _ip, C.c_int, C.c_int, C.c_int, C.c_double,
                                                 C.c_double, C.c_double, C.c_int, _dp, _dp]
        L.oracle_lower_bound_aspect.restype = C.c_double
        L.oracle_lower_bound_aspect.argtypes = [_dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_double, C.c_double,
                                                C.c_double, _ip, C.c_int, C.c_int, C.c_int, _dp]
        L.oracle_gibbs_dirber_finite.restype = C.c_int
        L.oracle_gibbs_dirber_finite.argtypes = [_ip, _ip, C.c_int, C.c_int, C.c_int, C.c_double,
                                                 C.c_double, C.c_double, C.c_int, C.c_double,
                                                 C.c_uint64, _ip]
        L.oracle_gibbs_dirber_finite_keyed.restype = C.c_int
        L.oracle_gibbs_dirber_finite_keyed.argtypes = L.oracle_gibbs_dirber_finite.argtypes
        L.oracle_gibbs_dirber_finite_keyed_wavefront.restype = C.c_int
        L.oracle_gibbs_dirber_finite_keyed_wavefront.argtypes = L.oracle_gibbs_dirber_finite.argtypes
        L.oracle_gibbs_dirber_dp_keyed.restype = C.c_int
        L.oracle_gibbs_dirber_dp_keyed.argtypes = [_ip, _ip, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double,
                                                   C.c_uint64, _ip, C.c_int]
        L.oracle_gibbs_dirdir_keyed.restype = C.c_int
        L.oracle_gibbs_dirdir_keyed.argtypes = [_ip, _ip, _ip, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                                C.c_int, C.c_double, C.c_uint64, _ip, _ip, C.c_int]
        L.oracle_gibbs_expectation.restype = C.c_int
        L.oracle_gibbs_expectation.argtypes = [_ip, _ip, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                               C.c_double, C.c_double, _dp, _dp, _dp]
        L.oracle_compute_expectation_W.restype = C.c_int
        L.oracle_compute_expectation_W.argtypes = [_ip, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, _dp]
        L.oracle_compute_expectation_H.restype = C.c_int
        L.oracle_compute_expectation_H.argtypes = [_ip, _ip, C.c_int, C.c_int, C.c_int, C.c_double,
                                                   C.c_double, C.c_int, _dp]
        L.oracle_gibbs_uncollapsed.restype = C.c_int
        L.oracle_gibbs_uncollapsed.argtypes = [_ip, _ip, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                               C.c_double, C.c_int, C.c_double, C.c_uint64, _dp, _dp, _dp]
        L.oracle_loglikelihood.restype = C.c_double
        L.oracle_loglikelihood.argtypes = [_dp, _dp, _ip, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64)]
        L.oracle_loglikelihood_EV.restype = C.c_double
        L.oracle_loglikelihood_EV.argtypes = [_dp, _ip, C.c_int, C.c_int]
        L.oracle_generate.restype = C.c_int
        L.oracle_generate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_uint64, _ip, _dp, _dp]
    return _lib


def _i32(a):
    a = np.asarray(a)
    if a.dtype.kind == "f":
        out = np.full(a.shape, NA, dtype=np.int32)
        m = ~np.isnan(a)
        out[m] = a[m].astype(np.int32)
        a = out
    return np.asfortranarray(a, dtype=np.int32)


def _p(a, t=_dp):
    return a.ctypes.data_as(t) if a is not None else None


def digamma(x: float) -> float:
    return lib().oracle_digamma(float(x))


def generate(F, N, Kt, a, b, seed):
    """R/utils.R:7-16 + xp_paper_basic_example.R:22-30 generator (oracle RNG)."""
    V = np.zeros((F, N), dtype=np.int32, order="F")
    W = np.zeros((F, Kt), order="F")
    H = np.zeros((N, Kt), order="F")
    rc = lib().oracle_generate(F, N, Kt, a, b, seed, _p(V, _ip), _p(W), _p(H))
    assert rc == 0
    return V, W, H


def random_labels(V, K, seed):
    """BernoulliNMF.R:23-33: uniform labels on observed entries, 0-based, NA elsewhere."""
    V = _i32(V)
    rng = np.random.default_rng(seed)
    Z = rng.integers(0, K, size=V.shape, dtype=np.int32)
    Z = np.asfortranarray(Z)
    Z[V == NA] = NA
    return Z


def vb_aspect(V, Z, K, alpha=1.0, beta=1.0, gamma=1.0, iter=100, checklb=False, return_E_Z=False):
    V = _i32(V); Z = _i32(Z)
    F, N = V.shape
    E_W = np.zeros((F, K), order="F"); E_H = np.zeros((N, K), order="F")
    lbs = np.zeros(max(iter, 1))
    E_Z = np.zeros((F, K, N), order="F") if return_E_Z else None
    rc = lib().oracle_vb_aspect(_p(V, _ip), _p(Z, _ip), F, N, K, alpha, beta, gamma, iter, int(checklb),
                                _p(E_W), _p(E_H), _p(lbs), _p(E_Z))
    if rc < 0:
        raise MemoryError("oracle_vb_aspect")
    return {"E_W": E_W, "E_H": E_H, "lowerbounds": lbs[:iter], "E_Z": E_Z, "negative": rc == 1}


def vb_aspect_batch(V, Z, K, alpha=1.0, beta=1.0, gamma=1.0, iter=100, checklb=False, stats=None):
    V = _i32(V)
    F, N = V.shape
    Zc = _i32(Z) if Z is not None else np.zeros((F, N), dtype=np.int32, order="F")
    E_W = np.zeros((F, K), order="F"); E_H = np.zeros((N, K), order="F")
    lbs = np.zeros(max(iter, 1))
    st = np.zeros(F * K + 2 * N * K)
    use = 0
    if stats is not None:
        nt, fo, fz = stats
        st[:F * K] = np.asfortranarray(nt).ravel(order="F")
        st[F * K:F * K + N * K] = np.asfortranarray(fo).ravel(order="F")
        st[F * K + N * K:] = np.asfortranarray(fz).ravel(order="F")
        use = 1
    rc = lib().oracle_vb_aspect_batch(_p(V, _ip), _p(Zc, _ip), F, N, K, alpha, beta, gamma, iter,
                                      int(checklb), _p(E_W), _p(E_H), _p(lbs), _p(st), use)
    assert rc == 0
    out_stats = (st[:F * K].reshape((F, K), order="F"), st[F * K:F * K + N * K].reshape((N, K), order="F"),
                 st[F * K + N * K:].reshape((N, K), order="F"))
    return {"E_W": E_W, "E_H": E_H, "lowerbounds": lbs[:iter], "stats": out_stats}


def vb_dirber(V, Z, K, alpha=1.0, beta=1.0, gamma=1.0, iter=100, return_E_Z=False, wavefront=False):
    V = _i32(V); Z = _i32(Z)
    F, N = V.shape
    Kc = K + 1
    E_W = np.zeros((F, Kc), order="F"); E_H = np.zeros((N, Kc), order="F")
    E_Z = np.zeros((F, Kc, N), order="F") if return_E_Z else None
    if wavefront:
        rc = lib().oracle_vb_dirber_wavefront(_p(V, _ip), _p(Z, _ip), F, N, K, alpha, beta, gamma, iter,
                                              _p(E_W), _p(E_H))
    else:
        rc = lib().oracle_vb_dirber(_p(V, _ip), _p(Z, _ip), F, N, K, alpha, beta, gamma, iter, _p(E_W),
                                    _p(E_H), _p(E_Z))
    if rc < 0:
        raise MemoryError("oracle_vb_dirber")
    return {"E_W": E_W, "E_H": E_H, "E_Z": E_Z}


def gibbs_dirber_finite(V, Z, K, alpha=1.0, beta=1.0, gamma=1.0, iter=100, burnin=0.5, seed=1):
    V = _i32(V); Z = _i32(Z).copy(order="F")
    F, N = V.shape
    ns = iter - int(np.floor(iter * burnin))
    Zs = np.zeros((F, N, max(ns, 1)), dtype=np.int32, order="F")
    j = lib().oracle_gibbs_dirber_finite(_p(V, _ip), _p(Z, _ip), F, N, K, alpha, beta, gamma, iter, burnin,
                                         seed, _p(Zs, _ip))
    assert j >= 0
    return {"Z_samples": Zs[:, :, :ns], "written": j, "Z": Z}


def gibbs_dirber_finite_keyed(V, Z, K, alpha=1.0, beta=1.0, gamma=1.0, iter=100, burnin=0.5, seed=1, wavefront=False):
    """The reference's collapsed sweep with the per-entry Philox key (seed; f, n, sweep) as its uniform;
    `wavefront=True` visits the entries by anti-diagonals instead of column-major."""
    V = _i32(V); Z = _i32(Z).copy(order="F")
    F, N = V.shape
    ns = iter - int(np.floor(iter * burnin))
    Zs = np.zeros((F, N, max(ns, 1)), dtype=np.int32, order="F")
    fn = lib().oracle_gibbs_dirber_finite_keyed_wavefront if wavefront else lib().oracle_gibbs_dirber_finite_keyed
    j = fn(_p(V, _ip), _p(Z, _ip), F, N, K, alpha, beta, gamma, iter, burnin,
                                               seed, _p(Zs, _ip))
    assert j >= 0
    return {"Z_samples": Zs[:, :, :ns], "written": j, "Z": Z}


def gibbs_dirber_dp_keyed(V, Z, gamma=1.0, iter=100, burnin=0.5, seed=1, wavefront=False):
    """Gibbs_DirBer_DP_Rcpp (collapsed_gibbs_z.cpp:275-359) with the keyed uniform; Kmax = 200, V NA-free.
    `wavefront=2`: column-major on the sequential stream instead (the form pinned against the reference)."""
    V = _i32(V); Z = _i32(Z).copy(order="F")
    F, N = V.shape
    ns = iter - int(np.floor(iter * burnin))
    Zs = np.zeros((F, N, max(ns, 1)), dtype=np.int32, order="F")
    j = lib().oracle_gibbs_dirber_dp_keyed(_p(V, _ip), _p(Z, _ip), F, N, gamma, iter, burnin, seed, _p(Zs, _ip),
                                           int(wavefront))
    if j == -2:
        raise ValueError("Gibbs_DirBer_DP_Rcpp is undefined for V with NA (no NA test at :308-310)")
    assert j >= 0
    return {"Z_samples": Zs[:, :, :ns], "written": j, "Z": Z}


def gibbs_dirdir_keyed(V, Z, K, alpha=1.0, gamma=1.0, iter=100, burnin=0.5, seed=1, wavefront=False):
    """Gibbs_DirDir_finite_Rcpp (collapsed_gibbs_z.cpp:441-557) with the keyed uniforms (`wavefront=2`: the
    sequential stream, column-major)."""
    V = _i32(V); Z = _i32(Z).copy(order="F")
    Cm = np.zeros_like(Z, order="F")
    F, N = V.shape
    ns = iter - int(np.floor(iter * burnin))
    Zs = np.zeros((F, N, max(ns, 1)), dtype=np.int32, order="F")
    Cs = np.zeros((F, N, max(ns, 1)), dtype=np.int32, order="F")
    j = lib().oracle_gibbs_dirdir_keyed(_p(V, _ip), _p(Z, _ip), _p(Cm, _ip), F, N, K, alpha, gamma, iter, burnin, seed,
                                        _p(Zs, _ip), _p(Cs, _ip), int(wavefront))
    assert j >= 0
    return {"Z_samples": Zs[:, :, :ns], "C_samples": Cs[:, :, :ns], "written": j, "Z": Z, "C": Cm}


def gibbs_expectation(V, Z_samples, K, alpha, beta, gamma, want_EV=True):
    V = _i32(V)
    Zs = np.asfortranarray(Z_samples, dtype=np.int32)
    F, N, J = Zs.shape
    E_W = np.zeros((F, K), order="F"); E_H = np.zeros((N, K), order="F")
    E_V = np.zeros((F, N), order="F") if want_EV else None
    rc = lib().oracle_gibbs_expectation(_p(V, _ip), _p(Zs, _ip), F, N, K, J, alpha, beta, gamma, _p(E_W),
                                        _p(E_H), _p(E_V))
    assert rc == 0
    return E_W, E_H, E_V


def gibbs_uncollapsed(V, Z, K, alpha=1.0, beta=1.0, gamma=1.0, iter=100, burnin=0.5, seed=1, want_EV=True):
    V = _i32(V); Z = _i32(Z).copy(order="F")
    F, N = V.shape
    E_W = np.zeros((F, K), order="F"); E_H = np.zeros((N, K), order="F")
    E_V = np.zeros((F, N), order="F") if want_EV else None
    kept = lib().oracle_gibbs_uncollapsed(_p(V, _ip), _p(Z, _ip), F, N, K, alpha, beta, gamma, iter, burnin,
                                          seed, _p(E_W), _p(E_H), _p(E_V))
    assert kept >= 0
    return {"E_W": E_W, "E_H": E_H, "E_V": E_V, "kept": kept, "Z": Z}


def loglikelihood(E_W, E_H, V_test):
    Vt = _i32(V_test)
    F, N = Vt.shape
    E_W = np.asfortranarray(E_W, dtype=np.float64); E_H = np.asfortranarray(E_H, dtype=np.float64)
    n = C.c_int64(0)
    ll = lib().oracle_loglikelihood(_p(E_W), _p(E_H), _p(Vt, _ip), F, N, E_W.shape[1], C.byref(n))
    return ll, n.value


def loglikelihood_EV(E_V, V_test):
    Vt = _i32(V_test)
    E_V = np.asfortranarray(E_V, dtype=np.float64)
    return lib().oracle_loglikelihood_EV(_p(E_V), _p(Vt, _ip), Vt.shape[0], Vt.shape[1])


def compute_expectation_W(Z_samples, gamma):
    """compute_expectation_W_Rcpp (collapsed_gibbs_z.cpp:178-197)."""
    Zs = np.asfortranarray(Z_samples, dtype=np.int32)
    F, N, J = Zs.shape
    Kq = int(Zs.max()) + 1
    E = np.zeros((F, Kq), order="F")
    assert lib().oracle_compute_expectation_W(_p(Zs, _ip), F, N, J, gamma, Kq, _p(E)) == 0
    return E


def compute_expectation_H(Z_samples, V, alpha, beta):
    """compute_expectation_H_Rcpp (collapsed_gibbs_z.cpp:230-272), last column left at 0 (:236-242)."""
    Zs = np.asfortranarray(Z_samples, dtype=np.int32); V = _i32(V)
    F, N, J = Zs.shape
    Kq = int(Zs.max()) + 1
    E = np.zeros((N, Kq), order="F")
    assert lib().oracle_compute_expectation_H(_p(Zs, _ip), _p(V, _ip), F, N, J, alpha, beta, Kq, _p(E)) == 0
    return E


def lower_bound_aspect(E_Z, E_log_W, E_log_H, E_log_1_H, gamma_vb, alpha_vb, beta_vb, alpha, beta, gamma, V):
    """lower_bound_aspect (collapsed_gibbs_z_VB.cpp:339-369): (lb, terms7 = pW,pZ,pH,pV,qW,qZ,qH)."""
    V = _i32(V)
    F, N = V.shape
    f = lambda a: np.asfortranarray(a, dtype=np.float64)
    a = [f(x) for x in (E_Z, E_log_W, E_log_H, E_log_1_H, gamma_vb, alpha_vb, beta_vb)]
    K = a[1].shape[1]
    t7 = np.zeros(7)
    lb = lib().oracle_lower_bound_aspect(*[_p(x) for x in a], alpha, beta, gamma, _p(V, _ip), F, K, N, _p(t7))
    return lb, t7


def sample_gibbs_ZH_column(v_n, W, labels0, alpha=1.0, iter=100, burnin=0.5, seed=1):
    """One column of the E-step with h collapsed (R/MCEM - sample_gibbs.R:1-37 sample_gibbs_Z, the chain
    sample_gibbs_cpp, sample_gibbs.cpp:22-80, would run if it stored its draw): NumPy restatement with its
    own RNG -- compared with the device chain within Monte-Carlo error.  Returns the mean one-hot matrix
    over the initial state and the kept sweeps (i >= floor(iter*burnin), :73-78)."""
    rng = np.random.default_rng(seed)
    v = np.asarray(v_n, dtype=np.int64); W = np.asarray(W, dtype=np.float64)
    F, K = W.shape
    lab = np.asarray(labels0, dtype=np.int64).copy()
    L = np.bincount(lab, minlength=K).astype(np.float64)
    lik = np.where(v[:, None] == 1, W, 1.0 - W)
    acc = np.zeros((F, K)); acc[np.arange(F), lab] += 1.0
    kept = 1
    nb = int(np.floor(iter * burnin))
    for i in range(1, iter):
        u = rng.random(F)
        for f in range(F):
            L[lab[f]] -= 1.0
            p = (alpha + L) * lik[f]
            c = np.cumsum(p)
            k = int(np.searchsorted(c, u[f] * c[-1], side="right"))
            k = min(k, K - 1)
            lab[f] = k; L[k] += 1.0
        if i >= nb:
            acc[np.arange(F), lab] += 1.0
            kept += 1
    return acc / kept
