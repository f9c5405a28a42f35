"""A NumPy model of the tensor path's arithmetic (DESIGN 5.1 "precision design"), runnable without a GPU:
power-of-two operand scaling into [2^14, 2^15), f16 hi (+ lo) operands, fp32 accumulation, ONE reciprocal per
two entries (1/a = b/(ab), an unselected entry standing in as 2^64), R rounded to f16 with saturation, exact-count
row renormalisation.  It documents why one f16 pass is enough for stage 1, why the W_lo term matters in stage 2,
and that the shared reciprocal costs nothing -- and it fails if someone breaks those assumptions in the design."""
import numpy as np
import scipy.special as sp


def _operands(F, N, K, seed):
    rng = np.random.default_rng(seed)
    W = rng.dirichlet(np.full(4, 0.8), size=F); H = rng.beta(0.8, 0.8, size=(N, 4))
    V = (rng.random((F, N)) < W @ H.T).astype(np.int8)
    Z = rng.integers(0, K, size=(F, N))
    nt = np.zeros((F, K)); fo = np.zeros((N, K)); fz = np.zeros((N, K))
    for k in range(K):
        m = Z == k
        nt[:, k] = m.sum(1); fo[:, k] = (m & (V == 1)).sum(0); fz[:, k] = (m & (V == 0)).sum(0)
    g = 1.0 / K + nt; a = 1.0 + fo; b = 1.0 + fz
    A = np.exp(sp.digamma(g) - sp.digamma(g.sum(1, keepdims=True)))
    B1 = np.exp(sp.digamma(a) - sp.digamma(a + b)); B0 = np.exp(sp.digamma(b) - sp.digamma(a + b))
    return V, A, B1, B0


def _split(X):
    """scale so that max lies in [2^14, 2^15); f16 hi and lo parts (as float64 values) and the exponent"""
    e = 14 - int(np.floor(np.log2(X.max())))
    Xs = np.ldexp(X, e)
    hi = Xs.astype(np.float16).astype(np.float64)
    lo = (Xs - hi).astype(np.float16).astype(np.float64)
    return hi, lo, e


def _rows_statistic(V, A, B1, B0, use_lo, pair_rcp):
    """n_total = A (.) (R1 B1 + R0 B0) the way the rows pass computes it"""
    F, N = V.shape
    Ah, _, eA = _split(A)
    BB = np.empty((2 * N, A.shape[1])); BB[1::2] = B1; BB[0::2] = B0
    Bh, Bl, eB = _split(BB)
    S = (Ah.astype(np.float32) @ Bh.astype(np.float32).T).astype(np.float32)       # stage 1: hi parts only
    sel = np.zeros((F, 2 * N), dtype=bool)
    sel[:, 1::2] = V == 1; sel[:, 0::2] = V == 0
    c0 = np.float32(2.0 ** (eA + eB - 14))
    if pair_rcp:   # the selected entries of columns (n, n+1) share a reciprocal
        s = np.where(V == 1, S[:, 1::2], S[:, 0::2]).astype(np.float32)                   # F x N selected values
        p = (s[:, 0::2] * s[:, 1::2] + np.float32(2.0 ** -60)).astype(np.float32)
        t = (c0 * (np.float32(1) / p)).astype(np.float32)
        r = np.empty_like(s)
        r[:, 0::2] = t * s[:, 1::2]; r[:, 1::2] = t * s[:, 0::2]
    else:
        r = (c0 / np.where(V == 1, S[:, 1::2], S[:, 0::2])).astype(np.float32)
    r16 = np.minimum(r, 65504).astype(np.float16).astype(np.float32)
    R = np.zeros((F, 2 * N), dtype=np.float32)
    R[:, 1::2] = np.where(V == 1, r16, 0); R[:, 0::2] = np.where(V == 0, r16, 0)
    Wst = (Bh + (Bl if use_lo else 0)).astype(np.float32)
    G = (R @ Wst).astype(np.float64) * 2.0 ** (14 - eB)                                   # stage 2, un-scaled
    stat = A * G
    return stat * (N / stat.sum(1, keepdims=True))                                        # exact-count renormalisation


def _exact(V, A, B1, B0):
    D = np.where(V == 1, A @ B1.T, A @ B0.T)
    return A * ((np.where(V == 1, 1 / D, 0)) @ B1 + (np.where(V == 0, 1 / D, 0)) @ B0)


def test_precision_design_of_the_tensor_path():
    V, A, B1, B0 = _operands(256, 512, 64, 5)
    ref = _exact(V, A, B1, B0)
    rel = lambda x: np.linalg.norm(x - ref) / np.linalg.norm(ref)
    e_full = rel(_rows_statistic(V, A, B1, B0, use_lo=True, pair_rcp=True))
    e_hi = rel(_rows_statistic(V, A, B1, B0, use_lo=False, pair_rcp=True))
    e_single = rel(_rows_statistic(V, A, B1, B0, use_lo=True, pair_rcp=False))
    assert np.allclose(ref.sum(1), V.shape[1])            # every observed entry's E_Z sums to one
    assert e_full < 1e-5                                  # the gate of the north star, with margin at this small size
    assert e_hi < 6e-5 and e_hi > 1.5 * e_full            # dropping W_lo costs accuracy at short reductions ...
    assert abs(e_single - e_full) < 0.2 * e_full          # ... sharing a reciprocal between two entries does not


def test_hi_only_error_falls_with_the_reduction_length():
    """the W_lo term buys less as the reduction gets longer (why AUTO drops it at 131072)"""
    errs = []
    for N in (128, 1024):
        V, A, B1, B0 = _operands(128, N, 64, 7)
        ref = _exact(V, A, B1, B0)
        errs.append(np.linalg.norm(_rows_statistic(V, A, B1, B0, False, True) - ref) / np.linalg.norm(ref))
    assert errs[1] < 0.7 * errs[0]
