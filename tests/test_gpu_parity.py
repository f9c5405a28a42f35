"""GPU parity tests: the CUDA path, called through the C ABI (nbmf_b200.api ->
libnbmf_cuda), against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): VB results within 1e-5 relative on W, H and
the ELBO -- the FP64 FMA path is held to 1e-9; the exact CVB0 wavefront must be
bit-identical; Gibbs must agree within Monte-Carlo error over seeds.
"""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu
NA = oracle.NA


def make_problem(F, N, Kt, seed, na_frac=0.0, a=0.3, b=0.3):
    V, _, _ = oracle.generate(F, N, Kt, a, b, seed)
    if na_frac > 0:
        rng = np.random.default_rng(seed + 100)
        V = V.copy(order="F")
        V[rng.random(V.shape) < na_frac] = NA
    return V


def relF(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


# ---------------------------------------------------------------------------
# data path: packing round trip, statistics
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("F,N,na", [(1, 1, 0.0), (31, 33, 0.3), (64, 257, 0.0), (200, 300, 0.1), (513, 70, 0.5)])
def test_pack_roundtrip_bit_exact(F, N, na):
    from nbmf_b200 import Engine
    V = make_problem(F, N, 2, 3, na)
    e = Engine()
    e.set_data(V)
    assert np.array_equal(e.get_data(), V)
    assert e.dims() == (F, N, int((V != NA).sum()))
    e.close()


def test_non_binary_rejected():
    from nbmf_b200 import Engine, NbmfError
    V = np.asfortranarray(np.array([[0, 1], [2, 1]], dtype=np.int32))
    e = Engine()
    with pytest.raises(NbmfError):
        e.set_data(V)
    e.close()


@pytest.mark.parametrize("F,N,K,na", [(37, 53, 4, 0.25), (200, 300, 3, 0.0)])
def test_initial_statistics_are_label_histograms(F, N, K, na):
    from nbmf_b200 import Engine
    V = make_problem(F, N, 3, 5, na)
    Z = oracle.random_labels(V, K, 7)
    e = Engine(); e.set_data(V); e.init_from_labels(Z, K)
    nt, fo, fz = e.get_stats(K)
    obs = V != NA
    for k in range(K):
        assert np.array_equal(nt[:, k], ((Z == k) & obs).sum(1))
        assert np.array_equal(fo[:, k], ((Z == k) & obs & (V == 1)).sum(0))
        assert np.array_equal(fz[:, k], ((Z == k) & obs & (V == 0)).sum(0))
    # set/get round trip (the checkpoint path)
    e.set_stats(nt * 0.5, fo + 0.25, fz)
    a, b, c = e.get_stats(K)
    assert np.array_equal(a, nt * 0.5) and np.array_equal(b, fo + 0.25) and np.array_equal(c, fz)
    e.close()


def test_bad_labels_rejected():
    from nbmf_b200 import Engine, NbmfError
    V = make_problem(8, 8, 2, 1)
    Z = oracle.random_labels(V, 5, 1)
    e = Engine(); e.set_data(V)
    with pytest.raises(NbmfError):
        e.init_from_labels(Z, 3)  # labels up to 4 with Kc=3: the reference writes out of bounds
    e.close()


# ---------------------------------------------------------------------------
# VB_Aspect_Rcpp
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("F,N,K,na,gam,iters", [
    (200, 300, 3, 0.0, 1.0, 200),      # BASELINE config 1 shape
    (100, 543, 10, 0.2, 1.0, 60),      # config 2 shape (unvotes100 / 10, 20% masked)
    (23, 31, 4, 0.2, 0.7, 25),
    (65, 129, 16, 0.1, 1.0 / 16, 30),
    (50, 40, 33, 0.0, 1.0, 10),        # K > 32: three k-chunks per lane group
    (1, 7, 2, 0.0, 1.0, 5),            # ragged: a single row
])
def test_vb_aspect_matches_oracle(F, N, K, na, gam, iters):
    from nbmf_b200 import VB_Aspect_Rcpp
    from nbmf_b200 import _lib
    V = make_problem(F, N, 3, 42, na, a=0.1, b=0.1)
    Z = oracle.random_labels(V, K, 1)
    ref = oracle.vb_aspect(V, Z, K, 1.0, 1.0, gam, iter=iters, checklb=True, return_E_Z=(F * N * K < 2e5))
    got = VB_Aspect_Rcpp(V, Z, K, 1.0, 1.0, gam, iters, True, precision=_lib.PREC_FP64)
    assert relF(got["E_W"], ref["E_W"]) < 1e-9
    assert relF(got["E_H"], ref["E_H"]) < 1e-9
    assert np.max(np.abs(got["E_W"] - ref["E_W"]) / np.maximum(np.abs(ref["E_W"]), 1e-12)) < 1e-5
    assert np.max(np.abs(got["lowerbounds"] - ref["lowerbounds"]) / np.abs(ref["lowerbounds"])) < 1e-10
    assert np.all(np.diff(got["lowerbounds"]) >= -1e-8 * np.abs(got["lowerbounds"][:-1]))
    assert np.all(np.isnan(got["likelihoods"])) and np.all(np.isnan(got["kldivergences"]))
    if ref["E_Z"] is not None:
        assert np.max(np.abs(got["E_Z"] - ref["E_Z"])) < 1e-9
        assert np.all(got["E_Z"][:, :, :][np.broadcast_to((V == NA)[:, None, :], got["E_Z"].shape)] == 0)


def test_vb_aspect_lowerbounds_zero_without_checklb():
    from nbmf_b200 import VB_Aspect_Rcpp
    V = make_problem(20, 30, 2, 4)
    Z = oracle.random_labels(V, 3, 1)
    got = VB_Aspect_Rcpp(V, Z, 3, iter=5, checklb=False)
    assert np.all(got["lowerbounds"] == 0.0)


def test_vb_aspect_wrapper_zero_bases_labels():
    from nbmf_b200 import VB_aspectRcpp
    V = make_problem(20, 30, 2, 4, 0.1)
    Z0 = oracle.random_labels(V, 3, 1)
    Z1 = Z0.copy(); Z1[Z1 != NA] += 1  # R's 1-based labels
    a = VB_aspectRcpp(V, Z1, 3, iter=8)
    b = oracle.vb_aspect(V, Z0, 3, iter=8)
    assert relF(a["E_W"], b["E_W"]) < 1e-9 and a["class"][0] == "VBaspect"


def test_lower_bound_entry_point():
    from nbmf_b200 import lower_bound_aspect
    V = make_problem(25, 35, 3, 8, 0.2)
    K = 4
    Z = oracle.random_labels(V, K, 2)
    prev = oracle.vb_aspect_batch(V, Z, K, 1.1, 0.9, 0.6, iter=6, checklb=False)
    ref = oracle.vb_aspect_batch(V, Z, K, 1.1, 0.9, 0.6, iter=7, checklb=True)["lowerbounds"][-1]
    nt, fo, fz = prev["stats"]
    got = lower_bound_aspect(V, nt, fo, fz, 1.1, 0.9, 0.6)
    assert abs(got - ref) < 1e-10 * abs(ref)


def test_vb_resume_from_statistics_checkpoint():
    """state = (n_total, f_ones, f_zeros): 10 iterations == 4, checkpoint, 6 more."""
    from nbmf_b200 import Engine
    V = make_problem(40, 60, 3, 9, 0.1)
    K = 5
    Z = oracle.random_labels(V, K, 3)
    e = Engine(); e.set_data(V); e.init_from_labels(Z, K)
    W10, H10, _, _ = e.vb_aspect(K, iter=10)
    e.init_from_labels(Z, K)
    e.vb_aspect(K, iter=4)
    st = e.get_stats(K)
    e2 = Engine(); e2.set_data(V); e2.set_stats(*st)
    W6, H6, _, _ = e2.vb_aspect(K, iter=6)
    assert np.array_equal(W10, W6) and np.array_equal(H10, H6)
    e.close(); e2.close()


# ---------------------------------------------------------------------------
# VB_DirBer_Rcpp
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("F,N,K,na,iters", [(200, 300, 3, 0.0, 30), (23, 31, 4, 0.2, 12), (100, 543, 10, 0.2, 6),
                                            (9, 40, 3, 0.0, 12), (33, 8, 2, 0.4, 12),
                                            (48, 70, 39, 0.1, 4),     # Kc = 40 > 32: lanes of a group take several k
                                            (300, 40, 16, 0.0, 4)])   # Kc = 17, diagonal x lanes > one block pass
def test_vb_dirber_wavefront_bit_identical(F, N, K, na, iters):
    from nbmf_b200 import VB_DirBer_Rcpp
    V = make_problem(F, N, 3, 17, na)
    Z = oracle.random_labels(V, K + 1, 4)
    ref = oracle.vb_dirber(V, Z, K, 1.0, 1.0, 0.7, iter=iters, return_E_Z=True)
    got = VB_DirBer_Rcpp(V, Z, K, 1.0, 1.0, 0.7, iters)
    assert got["E_W"].shape == (F, K + 1)
    assert np.array_equal(got["E_W"], ref["E_W"], equal_nan=True)
    assert np.array_equal(got["E_H"], ref["E_H"], equal_nan=True)
    assert np.array_equal(got["E_Z"], ref["E_Z"])


def test_vb_dirber_multiblock_wavefront():
    """min(F,N) > 1024 takes the cooperative multi-block schedule; still bit-identical."""
    from nbmf_b200 import VB_DirBer_Rcpp
    F, N, K = 1100, 1200, 2
    V = make_problem(F, N, 2, 3, 0.05)
    Z = oracle.random_labels(V, K + 1, 4)
    ref = oracle.vb_dirber(V, Z, K, iter=3)
    got = VB_DirBer_Rcpp(V, Z, K, iter=3)
    assert np.array_equal(got["E_W"], ref["E_W"], equal_nan=True)
    assert np.array_equal(got["E_H"], ref["E_H"], equal_nan=True)


def test_vb_dirber_contraction_mode_is_cvb0_without_self_term():
    from nbmf_b200 import VB_DirBer_Rcpp, _lib
    F, N, K = 60, 80, 3
    V = make_problem(F, N, 3, 6, 0.1)
    Z = oracle.random_labels(V, K + 1, 4)
    got = VB_DirBer_Rcpp(V, Z, K, 1.0, 1.0, 1.0, 8, mode=_lib.DIRBER_CONTRACTION)
    # numpy restatement of the documented deviation
    obs = V != NA
    Kc = K + 1
    nt = np.stack([((Z == k) & obs).sum(1) for k in range(Kc)], 1).astype(float)
    fo = np.stack([((Z == k) & obs & (V == 1)).sum(0) for k in range(Kc)], 1).astype(float)
    fz = np.stack([((Z == k) & obs & (V == 0)).sum(0) for k in range(Kc)], 1).astype(float)
    for _ in range(7):
        g, a, b = 1.0 + nt, 1.0 + fo, 1.0 + fz
        A, B1, B0 = g, a / (a + b), b / (a + b)
        D = np.where(V == 1, A @ B1.T, A @ B0.T)
        R1 = np.where(obs & (V == 1), 1 / D, 0.0); R0 = np.where(obs & (V == 0), 1 / D, 0.0)
        nt = A * (R1 @ B1 + R0 @ B0); fo = B1 * (R1.T @ A); fz = B0 * (R0.T @ A)
    assert relF(got["E_W"], g / g.sum(1, keepdims=True)) < 1e-10
    assert relF(got["E_H"], a / (a + b)) < 1e-10


# ---------------------------------------------------------------------------
# predictive log-likelihood
# ---------------------------------------------------------------------------
def test_predictive_ll_matches_oracle():
    from nbmf_b200 import loglikelihood
    V = make_problem(70, 90, 3, 6, 0.6)
    rng = np.random.default_rng(1)
    W = rng.dirichlet(np.ones(6), size=70); H = rng.random((90, 6))
    ref, n = oracle.loglikelihood(W, H, V)
    got = loglikelihood({"E_W": W, "E_H": H}, V)
    assert got["n"] == n and abs(got["loglikelihood"] - ref) < 1e-10 * abs(ref)


# ---------------------------------------------------------------------------
# Gibbs: statistical parity with the reference's collapsed sampler
# ---------------------------------------------------------------------------
def test_gibbs_counts_are_consistent():
    from nbmf_b200 import Gibbs_DirBer_finite_Rcpp
    V = make_problem(40, 50, 3, 2, 0.1)
    K = 4
    Z = oracle.random_labels(V, K, 1)
    r = Gibbs_DirBer_finite_Rcpp(V, Z, K, iter=30, burnin=0.5, seed=3)
    obs = V != NA
    assert np.all((r["Z_last"][obs] >= 0) & (r["Z_last"][obs] < K)) and np.all(r["Z_last"][~obs] == NA)
    assert np.allclose(r["E_W"].sum(1), 1.0, atol=1e-9)
    assert np.all((r["E_H"] > 0) & (r["E_H"] < 1))
    # determinism: same seed, same chain (counter-based Philox)
    r2 = Gibbs_DirBer_finite_Rcpp(V, Z, K, iter=30, burnin=0.5, seed=3)
    assert np.array_equal(r["Z_last"], r2["Z_last"]) and np.array_equal(r["E_W"], r2["E_W"])
    r3 = Gibbs_DirBer_finite_Rcpp(V, Z, K, iter=30, burnin=0.5, seed=4)
    assert not np.array_equal(r["Z_last"], r3["Z_last"])


def test_gibbs_matches_reference_sampler_within_mc_error():
    from nbmf_b200 import Gibbs_DirBer_finite_Rcpp
    F, N, K = 30, 40, 3
    V = make_problem(F, N, 3, 2, 0.0, a=0.5, b=0.5)
    rng = np.random.default_rng(0)
    test = rng.random(V.shape) < 0.2
    Vtr = V.copy(order="F"); Vtr[test] = NA
    fidx, nidx = np.nonzero(test)
    idx = (fidx + F * nidx).astype(np.int64)
    vt = V[test]
    ll_ref, ll_gpu, ev_ref, ev_gpu = [], [], [], []
    for s in (1, 2, 3, 4, 5):
        Z = oracle.random_labels(Vtr, K, s)
        c = oracle.gibbs_dirber_finite(Vtr, Z, K, iter=400, burnin=0.5, seed=s)
        _, _, EV = oracle.gibbs_expectation(Vtr, c["Z_samples"][:, :, :c["written"]], K, 1.0, 1.0, 1.0)
        p = EV[test]
        ll_ref.append(np.mean(np.log(np.where(vt == 1, p, 1 - p)))); ev_ref.append(p)
        g = Gibbs_DirBer_finite_Rcpp(Vtr, Z, K, 1.0, 1.0, 1.0, 2000, 0.5, seed=s, test_idx=idx)
        q = g["E_V_test"]
        ll_gpu.append(np.mean(np.log(np.where(vt == 1, q, 1 - q)))); ev_gpu.append(q)
    spread = max(np.ptp(ll_ref), np.ptp(ll_gpu), 0.004)
    assert abs(np.mean(ll_ref) - np.mean(ll_gpu)) < 2 * spread
    between = np.sqrt(np.mean((np.mean(ev_ref, 0) - np.mean(ev_gpu, 0)) ** 2))
    within = np.sqrt(np.mean((ev_ref[0] - ev_ref[1]) ** 2))
    assert between < 2 * max(within, 0.01)


def test_gibbs_sweep_given_W_posterior_means():
    """nbmf_gibbs_sweep_given_W (sample_gibbs.cpp:22-80 / sample_gibbs_ZH): rows whose
    likelihood strongly favours one component end up there."""
    from nbmf_b200 import Engine
    F, K = 64, 3
    W = np.full((F, K), 0.05); W[:, 0] = 0.95
    v = np.ones(F, dtype=np.int32)
    e = Engine()
    Cm = e.gibbs_sweep_given_W(v, W, alpha=1.0, iter=200, burnin=0.5, seed=1)
    e.close()
    assert np.allclose(Cm.sum(1), 1.0, atol=1e-9)
    assert Cm[:, 0].mean() > 0.85


# ---------------------------------------------------------------------------
# synthetic generator + size-independent properties at a larger size
# ---------------------------------------------------------------------------
def test_synthetic_generator_and_properties():
    from nbmf_b200 import Engine
    F, N, K = 1024, 1536, 8
    e = Engine()
    e.generate_synthetic(F, N, 4, 1.0, 1.0, 4)
    V = e.get_data()
    assert V.shape == (F, N) and set(np.unique(V)) <= {0, 1}
    assert 0.3 < V.mean() < 0.7
    # sharding invariance of the generator: columns [512, 1024) of the same seed
    e2 = Engine(n_global=N, n_offset=512)
    e2.generate_synthetic(F, 512, 4, 1.0, 1.0, 4)
    assert np.array_equal(e2.get_data(), V[:, 512:1024])
    e2.close()
    e.init_random_labels(K, 5)
    nt, fo, fz = e.get_stats(K)
    assert nt.sum() == F * N and np.array_equal(nt.sum(1), np.full(F, N))
    assert np.array_equal((fo + fz).sum(1), np.full(N, F)) and fo.sum() == V.sum()
    # one VB iteration conserves the observation count: sum_k E_Z = 1 per entry
    e.vb_aspect(K, iter=3)
    nt, fo, fz = e.get_stats(K)
    assert np.allclose(nt.sum(1), N, rtol=1e-12) and np.allclose((fo + fz).sum(1), F, rtol=1e-12)
    assert np.allclose(fo.sum(1), V.sum(0), rtol=1e-12)
    # and agrees with the oracle started from the same statistics
    e.init_random_labels(K, 5)
    st = e.get_stats(K)
    W, H, lbs, _ = e.vb_aspect(K, iter=3, checklb=True)
    ref = oracle.vb_aspect_batch(V, None, K, iter=3, checklb=True, stats=st)
    assert relF(W, ref["E_W"]) < 1e-10 and relF(H, ref["E_H"]) < 1e-10
    assert np.max(np.abs(lbs - ref["lowerbounds"]) / np.abs(ref["lowerbounds"])) < 1e-11
    e.close()


# ---------------------------------------------------------------------------
# against the golden vectors produced by the reference's own sources
# (tests/golden/make_golden.py, oracle/ref_shim)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("c", ["a", "b", "c"])
def test_gpu_matches_reference_golden_vectors(c):
    import os
    from nbmf_b200 import VB_Aspect_Rcpp, VB_DirBer_Rcpp
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors.npz"))
    K, al, be, ga, it = G[f"aspect_{c}_par"]
    V = G[f"aspect_{c}_V"]
    got = VB_Aspect_Rcpp(V, G[f"aspect_{c}_Z"], int(K), al, be, ga, int(it), True)
    assert relF(got["E_W"], G[f"aspect_{c}_E_W"]) < 1e-10 and relF(got["E_H"], G[f"aspect_{c}_E_H"]) < 1e-10
    assert np.max(np.abs(got["lowerbounds"] - G[f"aspect_{c}_lowerbounds"]) / np.abs(G[f"aspect_{c}_lowerbounds"])) < 1e-11
    assert np.max(np.abs(got["E_Z"] - G[f"aspect_{c}_E_Z"])) < 1e-10
    d = VB_DirBer_Rcpp(V, G[f"dirber_{c}_Z"], int(K), al, be, ga, max(int(it) // 2, 2))
    for k in ("E_W", "E_H", "E_Z"):
        assert np.array_equal(d[k], G[f"dirber_{c}_{k}"], equal_nan=True), k


def test_argmax_labels_hand_off_matches_which_max():
    """BernoulliNMF.R:45-52: Z <- which.max(E_Z[f,,n]) after the VB warm start, on the device."""
    from nbmf_b200 import Engine
    V = make_problem(45, 60, 3, 12, 0.15)
    K = 6
    Z = oracle.random_labels(V, K, 2)
    ref = oracle.vb_aspect(V, Z, K, 1.0, 1.0, 0.5, iter=9, return_E_Z=True)
    e = Engine(); e.set_data(V); e.init_from_labels(Z, K)
    e.vb_aspect(K, 1.0, 1.0, 0.5, iter=9)
    Zg = e.argmax_labels(want_host=True)
    e.close()
    obs = V != NA
    want = np.argmax(ref["E_Z"], axis=1)
    # ties / near-ties aside (none at this tolerance), identical
    top2 = np.sort(ref["E_Z"], axis=1)[:, -2:, :]
    clear = obs & ((top2[:, 1, :] - top2[:, 0, :]) > 1e-9)
    assert np.array_equal(Zg[clear], want[clear]) and np.all(Zg[~obs] == NA)


def test_allreduce_hook_plumbing_single_rank():
    """The exchange hook is called once per iteration with the F x Kp device buffer (plus once per
    label initialisation); a sum over a world of one leaves the results unchanged, and the ELBO
    terms split into local (sum log D, column prior) and replicated (row prior) parts."""
    from nbmf_b200 import Engine
    from nbmf_b200.dist import combine_lowerbounds
    V = make_problem(70, 90, 3, 21, 0.1)
    K = 5
    Z = oracle.random_labels(V, K, 2)
    calls = []
    e = Engine(n_global=90, n_offset=0); e.set_data(V)
    e.set_allreduce_hook(lambda ptr, count, stream: calls.append((ptr != 0, count, stream != 0)))
    e.init_from_labels(Z, K)
    W, H, lb, _ = e.vb_aspect(K, 1.0, 1.0, 0.5, iter=7, checklb=True)
    terms = e.vb_lb_terms(7)
    e.close()
    assert len(calls) == 1 + 7 and all(c[0] and c[2] for c in calls) and all(c[1] >= 70 * K for c in calls)
    ref = oracle.vb_aspect(V, Z, K, 1.0, 1.0, 0.5, iter=7, checklb=True)
    assert relF(W, ref["E_W"]) < 1e-9 and relF(H, ref["E_H"]) < 1e-9
    assert np.allclose(combine_lowerbounds(terms, lambda x: x), lb, rtol=1e-14)
    assert np.max(np.abs(lb - ref["lowerbounds"]) / np.abs(ref["lowerbounds"])) < 1e-10


def test_degenerate_inputs():
    """All entries NA; a single observed entry; K=1."""
    from nbmf_b200 import Engine, VB_Aspect_Rcpp, VB_DirBer_Rcpp
    V = np.full((5, 6), NA, dtype=np.int32, order="F")
    Z = np.full((5, 6), NA, dtype=np.int32, order="F")
    r = VB_Aspect_Rcpp(V, Z, 3, iter=3, checklb=True)
    o = oracle.vb_aspect(V, Z, 3, iter=3, checklb=True)
    assert np.allclose(r["E_W"], 1 / 3) and np.allclose(r["E_H"], 0.5)
    assert np.allclose(r["lowerbounds"], o["lowerbounds"], rtol=1e-12, atol=1e-12)
    V[2, 4] = 1; Z[2, 4] = 0
    r = VB_Aspect_Rcpp(V, Z, 3, iter=4, checklb=True)
    o = oracle.vb_aspect(V, Z, 3, iter=4, checklb=True)
    assert relF(r["E_W"], o["E_W"]) < 1e-12 and relF(r["E_H"], o["E_H"]) < 1e-12
    d = VB_DirBer_Rcpp(V, Z, 2, iter=3)
    od = oracle.vb_dirber(V, Z, 2, iter=3)
    assert np.array_equal(d["E_W"], od["E_W"], equal_nan=True) and np.array_equal(d["E_H"], od["E_H"], equal_nan=True)
    V2 = make_problem(9, 11, 2, 3)
    Z2 = oracle.random_labels(V2, 1, 1)
    r = VB_Aspect_Rcpp(V2, Z2, 1, iter=3, checklb=True)
    o = oracle.vb_aspect(V2, Z2, 1, iter=3, checklb=True)
    assert np.allclose(r["E_W"], 1.0) and relF(r["E_H"], o["E_H"]) < 1e-12
    assert np.allclose(r["lowerbounds"], o["lowerbounds"], rtol=1e-12)
