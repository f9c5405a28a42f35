"""GPU parity tests of the tcgen05 tensor path (32 < K <= 128), through the C ABI.

Gate (BASELINE.json north_star): W, H and the ELBO within 1e-5 relative of the
reference's fp64 VB.  W and H are compared norm-wise;
the element-wise error with a floor is reported by tools/tp_accuracy.py.
The FP64 FMA path (itself held to 1e-9 against the oracle in test_gpu_parity.py)
is the stand-in for the oracle at sizes the CPU oracle cannot finish in seconds.
"""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu
NA = oracle.NA
TOL = 1e-5


def relF(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def test_tensor_path_matches_oracle_medium():
    from nbmf_b200 import Engine, _lib
    F, N, K, iters = 512, 640, 64, 6
    V, _, _ = oracle.generate(F, N, 4, 0.8, 0.8, 5)
    V = V.copy(order="F")
    V[np.random.default_rng(2).random(V.shape) < 0.15] = NA
    Z = oracle.random_labels(V, K, 3)
    ref = oracle.vb_aspect_batch(V, Z, K, 1.0, 1.0, 0.5, iter=iters, checklb=True)
    e = Engine(precision=_lib.PREC_TENSOR)
    e.set_data(V); e.init_from_labels(Z, K)
    W, H, lb, _ = e.vb_aspect(K, 1.0, 1.0, 0.5, iter=iters, checklb=True)
    nt, fo, fz = e.get_stats(K)
    e.close()
    assert relF(W, ref["E_W"]) < TOL and relF(H, ref["E_H"]) < TOL
    assert np.max(np.abs(lb - ref["lowerbounds"]) / np.abs(ref["lowerbounds"])) < TOL
    # exact-count property of the statistics (sum_k E_Z = 1 per observed entry)
    obs = V != NA
    assert np.allclose(nt.sum(1), obs.sum(1), rtol=1e-12)
    assert np.allclose(fo.sum(1), (obs & (V == 1)).sum(0), rtol=1e-12)
    assert np.allclose(fz.sum(1), (obs & (V == 0)).sum(0), rtol=1e-12)


@pytest.mark.parametrize("F,N,K,na,iters,gamma", [
    (1024, 1536, 64, 0.1, 10, 1.0),
    (2048, 2048, 128, 0.0, 10, 1.0 / 128),
    (1500, 1100, 100, 0.2, 8, 0.5),      # K padded to 128, ragged tiles in both directions
    (4096, 4096, 128, 0.0, 20, 1.0 / 128),
    (300, 5000, 40, 0.05, 6, 1.0),       # K padded to 64, F smaller than two owner tiles
])
def test_tensor_path_matches_fp64_path(F, N, K, na, iters, gamma):
    from nbmf_b200 import Engine, _lib
    ef = Engine(precision=_lib.PREC_FP64)
    ef.generate_synthetic(F, N, 8, 1.0, 1.0, 11)
    V = ef.get_data()
    if na > 0:
        V = V.copy(order="F")
        V[np.random.default_rng(4).random(V.shape) < na] = NA
        ef.set_data(V)
    ef.init_random_labels(K, 12)
    st = ef.get_stats(K)
    Wf, Hf, lf, _ = ef.vb_aspect(K, 1.0, 1.0, gamma, iter=iters, checklb=True)
    ef.close()
    et = Engine(precision=_lib.PREC_TENSOR)
    et.set_data(V); et.set_stats(*st)
    Wt, Ht, lt, _ = et.vb_aspect(K, 1.0, 1.0, gamma, iter=iters, checklb=True)
    et.close()
    assert relF(Wt, Wf) < TOL, relF(Wt, Wf)
    assert relF(Ht, Hf) < TOL, relF(Ht, Hf)
    assert np.max(np.abs(lt - lf) / np.abs(lf)) < TOL
    assert np.all(np.diff(lt) >= -2e-6 * np.abs(lt[:-1]))     # ELBO never decreases (:487-494)
    assert np.allclose(Wt.sum(1), 1.0, atol=1e-12)


def test_tensor_single_update_and_fast_mode():
    """One update from identical statistics: the injected error per update is ~1e-6 or below; the
    opt-in hi-only mode (NBMF_PREC_TENSOR_FAST) stays within 1e-5 at this size."""
    from nbmf_b200 import Engine, _lib
    F = N = 4096; K = 128
    res = {}
    for prec in (_lib.PREC_FP64, _lib.PREC_TENSOR, _lib.PREC_TENSOR_FAST):
        e = Engine(precision=prec)
        e.generate_synthetic(F, N, 8, 1.0, 1.0, 5)
        e.init_random_labels(K, 6)
        e.vb_aspect(K, 1.0, 1.0, 1.0 / K, iter=3, checklb=True)
        res[prec] = e.get_stats(K)
        e.close()
    for prec, tol in ((_lib.PREC_TENSOR, 2e-6), (_lib.PREC_TENSOR_FAST, 1e-5)):
        for a, b in zip(res[prec], res[_lib.PREC_FP64]):
            assert relF(a, b) < tol, (prec, relF(a, b))


def test_auto_precision_selects_tensor_path_for_large_k():
    from nbmf_b200 import Engine, _lib
    e = Engine(precision=_lib.PREC_AUTO)
    e.generate_synthetic(1024, 1024, 4, 1.0, 1.0, 3)
    e.init_random_labels(64, 1)
    st = e.get_stats(64)
    W, H, _, _ = e.vb_aspect(64, iter=3)
    path = e.timing()["path"]
    # 3 updates at n = 1024: within the horizon of the e5m2-lo variant, beyond the hi-only one's
    assert path == _lib.PATH_TENSOR_LO8
    e2 = Engine(precision=_lib.PREC_TENSOR_LO8)
    e2.set_data(e.get_data()); e2.set_stats(*st)
    W2, H2, _, _ = e2.vb_aspect(64, iter=3)
    assert np.array_equal(W, W2) and np.array_equal(H, H2)   # same kernels, deterministic
    # the reference's default, 100 updates, is beyond every f16 variant's horizon: the FP64 tensor pipe takes it
    e.set_stats(*st)
    e.vb_aspect(64, iter=100)
    assert e.timing()["path"] == _lib.PATH_FP64
    # small K stays on the FP64 path even when asked for AUTO
    e.init_random_labels(8, 1)
    st8 = e.get_stats(8)
    W8, _, _, _ = e.vb_aspect(8, iter=2)
    e3 = Engine(precision=_lib.PREC_FP64)
    e3.set_data(e.get_data()); e3.set_stats(*st8)
    W8f, _, _, _ = e3.vb_aspect(8, iter=2)
    assert np.allclose(W8, W8f, rtol=1e-12, atol=0)   # FP64 path: atomics on split reductions reorder sums
    e.close(); e2.close(); e3.close()


def test_tensor_path_contraction_mode_dirber():
    """VB_DirBer contraction form (mean parameters) on the tensor path vs the FP64 path."""
    from nbmf_b200 import Engine, _lib
    K = 63  # Kc = K+1 = 64
    outs = []
    for prec in (_lib.PREC_FP64, _lib.PREC_TENSOR):
        e = Engine(precision=prec)
        e.generate_synthetic(1024, 1280, 4, 1.0, 1.0, 9)
        e.init_random_labels(K + 1, 2)
        outs.append(e.vb_dirber(K, 1.0, 1.0, 1.0, iter=6, mode=_lib.DIRBER_CONTRACTION))
        e.close()
    assert relF(outs[1][0], outs[0][0]) < TOL and relF(outs[1][1], outs[0][1]) < TOL


@pytest.mark.parametrize("F,N,K", [(257, 300, 33), (1000, 513, 128), (640, 1999, 65)])
def test_tensor_path_ragged_shapes_and_empty_lines(F, N, K):
    """Dimensions that are no multiple of any tile size, K at the edges of the padded ranges, plus a
    row and a column with no observed entry at all (their statistics stay 0, E_H is the prior mean)."""
    from nbmf_b200 import Engine, _lib
    V, _, _ = oracle.generate(F, N, 4, 0.8, 0.8, 5)
    V = V.copy(order="F")
    V[np.random.default_rng(2).random(V.shape) < 0.1] = NA
    V[3, :] = NA
    V[:, 7] = NA
    Z = oracle.random_labels(V, K, 3)
    out = {}
    for prec in (_lib.PREC_FP64, _lib.PREC_TENSOR):
        e = Engine(precision=prec)
        e.set_data(V); e.init_from_labels(Z, K)
        out[prec] = e.vb_aspect(K, 1.0, 2.0, 0.7, iter=5, checklb=True) + (e.get_stats(K),)
        e.close()
    Wf, Hf, lf, _, sf = out[_lib.PREC_FP64]
    Wt, Ht, lt, _, st = out[_lib.PREC_TENSOR]
    assert relF(Wt, Wf) < TOL and relF(Ht, Hf) < TOL and np.max(np.abs(lt - lf) / np.abs(lf)) < TOL
    assert np.all(st[0][3] == 0) and np.all(st[1][7] == 0) and np.all(st[2][7] == 0)
    assert np.allclose(Ht[7], 1.0 / 3.0) and np.allclose(Wt[3], 1.0 / K)
    if F * N * K < 3e7:   # and against the oracle itself where it finishes in seconds
        ref = oracle.vb_aspect_batch(V, Z, K, 1.0, 2.0, 0.7, iter=5, checklb=True)
        assert relF(Wf, ref["E_W"]) < 1e-9 and relF(Wt, ref["E_W"]) < TOL
