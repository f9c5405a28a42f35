"""CPU tests of the oracle itself (oracle/nbmf_oracle.c): the anchors the
reference offers for this path turned into permanent tests.

 * digamma / lgamma against scipy.special (the reference uses boost::math)
 * oracle_batch == oracle_seq          (Appendix E probe 1)
 * wavefront == column-major sweep     (Appendix E probe 2, bit-exact)
 * pZ + pV - qZ == sum log D           (identity behind the fused ELBO)
 * ELBO never decreases                (collapsed_gibbs_z_VB.cpp:487-494)
 * E_W rows sum to 1, E_H in (0,1), non-negativity asserts (:501-518)
 * collapsed vs uncollapsed Gibbs agree within Monte-Carlo error
"""
import numpy as np
import pytest
import scipy.special as sp

from oracle import oracle

NA = oracle.NA


def make_problem(F, N, Kt, seed, na_frac=0.0, a=0.3, b=0.3):
    V, W, H = oracle.generate(F, N, Kt, a, b, seed)
    if na_frac > 0:
        rng = np.random.default_rng(seed + 100)
        V = V.copy(order="F")
        V[rng.random(V.shape) < na_frac] = NA
    return V


def test_digamma_lgamma_match_scipy():
    xs = np.concatenate([np.logspace(-6, 6, 400), np.linspace(0.01, 30, 500)])
    d = np.array([oracle.digamma(x) for x in xs])
    ref = sp.digamma(xs)
    err = np.abs(d - ref) / np.maximum(1.0, np.abs(ref))
    assert err.max() < 5e-15
    L = oracle.lib()
    lg = np.array([L.oracle_lgamma(x) for x in xs])
    assert np.max(np.abs(lg - sp.gammaln(xs)) / np.maximum(1.0, np.abs(sp.gammaln(xs)))) < 1e-14


@pytest.mark.parametrize("F,N,K,na,gam", [(23, 31, 4, 0.2, 0.7), (40, 17, 3, 0.0, 1.0), (16, 50, 7, 0.35, 1.0 / 7)])
def test_batch_equals_sequential(F, N, K, na, gam):
    V = make_problem(F, N, 3, 11, na)
    Z = oracle.random_labels(V, K, 5)
    s = oracle.vb_aspect(V, Z, K, 1.0, 0.8, gam, iter=25, checklb=True)
    b = oracle.vb_aspect_batch(V, Z, K, 1.0, 0.8, gam, iter=25, checklb=True)
    assert np.max(np.abs(s["E_W"] - b["E_W"]) / np.abs(s["E_W"])) < 1e-11
    assert np.max(np.abs(s["E_H"] - b["E_H"]) / np.abs(s["E_H"])) < 1e-11
    assert np.max(np.abs(s["lowerbounds"] - b["lowerbounds"]) / np.abs(s["lowerbounds"])) < 1e-12


def test_elbo_monotone_and_invariants():
    V = make_problem(30, 45, 3, 3, 0.1)
    K = 5
    Z = oracle.random_labels(V, K, 9)
    r = oracle.vb_aspect(V, Z, K, 1.0, 1.0, 1.0, iter=60, checklb=True)
    lb = r["lowerbounds"]
    assert np.all(np.diff(lb) >= -1e-9 * np.abs(lb[:-1]))
    assert np.allclose(r["E_W"].sum(axis=1), 1.0, atol=1e-12)
    assert np.all(r["E_W"] >= 0) and np.all((r["E_H"] > 0) & (r["E_H"] < 1))
    assert not r["negative"]


def test_lowerbounds_zero_without_checklb():
    V = make_problem(10, 12, 2, 1)
    Z = oracle.random_labels(V, 3, 1)
    r = oracle.vb_aspect(V, Z, 3, iter=4, checklb=False)
    assert np.all(r["lowerbounds"] == 0.0)  # collapsed_gibbs_z_VB.cpp:474-477


def test_elbo_identity_sum_log_D():
    """E_log_p_Z + E_log_p_V - E_log_q_Z == sum_obs log D_fn at the q(Z) optimum."""
    V = make_problem(12, 15, 3, 21, 0.15)
    K = 4
    Z = oracle.random_labels(V, K, 2)
    it = 7
    full = oracle.vb_aspect(V, Z, K, 1.2, 0.9, 0.5, iter=it, checklb=True, return_E_Z=True)
    # statistics before the last update = state of a run with it-1 iterations
    prev = oracle.vb_aspect_batch(V, Z, K, 1.2, 0.9, 0.5, iter=it - 1, checklb=False)
    nt, fo, fz = prev["stats"]
    g = 0.5 + nt; a = 1.2 + fo; b = 0.9 + fz
    elw = sp.digamma(g) - sp.digamma(g.sum(1, keepdims=True))
    elh = sp.digamma(a) - sp.digamma(a + b); el1h = sp.digamma(b) - sp.digamma(a + b)
    A, B1, B0 = np.exp(elw), np.exp(elh), np.exp(el1h)
    obs = V != NA
    D = np.where(V == 1, A @ B1.T, A @ B0.T)
    sumlogD = np.log(D[obs]).sum()
    E_Z = full["E_Z"]
    F, N = V.shape
    pZ = pV = qZ = 0.0
    for n in range(N):
        for f in range(F):
            if obs[f, n]:
                e = E_Z[f, :, n]
                pZ += e @ elw[f]
                pV += e @ (elh[n] if V[f, n] == 1 else el1h[n])
                qZ += e @ np.log(e)
    assert abs((pZ + pV - qZ) - sumlogD) < 1e-9 * abs(sumlogD)


@pytest.mark.parametrize("F,N,K,na", [(23, 31, 4, 0.2), (9, 40, 3, 0.0), (33, 8, 2, 0.4)])
def test_wavefront_is_bit_identical(F, N, K, na):
    V = make_problem(F, N, 3, 17, na)
    Z = oracle.random_labels(V, K + 1, 4)
    s = oracle.vb_dirber(V, Z, K, 1.0, 1.0, 0.7, iter=12)
    w = oracle.vb_dirber(V, Z, K, 1.0, 1.0, 0.7, iter=12, wavefront=True)
    assert np.array_equal(s["E_W"], w["E_W"], equal_nan=True)
    assert np.array_equal(s["E_H"], w["E_H"], equal_nan=True)


def test_vb_dirber_quirks():
    """K+1 columns (:101), iter-1 sweeps (:119): iter=1 does no sweep at all."""
    V = make_problem(8, 9, 2, 5)
    Z = oracle.random_labels(V, 3, 4)
    r = oracle.vb_dirber(V, Z, 2, iter=5)
    assert r["E_W"].shape == (8, 3) and r["E_H"].shape == (9, 3)
    assert np.allclose(r["E_W"].sum(1), 1.0)
    r1 = oracle.vb_dirber(V, Z, 2, iter=1)
    assert np.all(np.isnan(r1["E_W"]))  # gamma_vb never written (uninitialised in the reference)


def test_gibbs_bookkeeping_and_quirks():
    V = make_problem(10, 14, 2, 8, 0.1)
    K = 3
    Z = oracle.random_labels(V, K, 3)
    g = oracle.gibbs_dirber_finite(V, Z, K, iter=11, burnin=0.5, seed=4)
    # nsamples = iter - floor(iter*burnin) = 6 slices, i in 5..10 written => 6
    assert g["Z_samples"].shape[2] == 6 and g["written"] == 6
    g0 = oracle.gibbs_dirber_finite(V, Z, K, iter=6, burnin=0.0, seed=4)
    # burnin=0: 6 slices allocated, only i=1..5 written -> last slice all zero (Appendix B.10)
    assert g0["written"] == 5 and np.all(g0["Z_samples"][:, :, 5] == 0)
    obs = V != NA
    assert np.all((g["Z_samples"][obs] >= 0) & (g["Z_samples"][obs] < K))
    # compute_expectation_H_Rcpp leaves the last column at zero (collapsed_gibbs_z.cpp:236-242)
    L = oracle.lib()
    Zs = np.asfortranarray(g["Z_samples"])
    Kq = int(Zs.max()) + 1
    EH = np.zeros((14, Kq), order="F")
    L.oracle_compute_expectation_H(oracle._p(Zs, oracle._ip), oracle._p(oracle._i32(V), oracle._ip), 10, 14, 6,
                                   1.0, 1.0, Kq, oracle._p(EH))
    assert np.all(EH[:, Kq - 1] == 0) and np.all(EH[:, :Kq - 1] > 0)
    EW = np.zeros((10, Kq), order="F")
    L.oracle_compute_expectation_W(oracle._p(Zs, oracle._ip), 10, 14, 6, 1.0, Kq, oracle._p(EW))
    assert np.allclose(EW.sum(1), 1.0)


@pytest.mark.parametrize("F,N,K,na", [(23, 31, 4, 0.2), (9, 40, 3, 0.0), (33, 8, 5, 0.4)])
def test_keyed_collapsed_gibbs_is_order_independent(F, N, K, na):
    """With the uniform of entry (f, n) in sweep i keyed by (seed; f, n, i), the reference's column-major
    sweep and the anti-diagonal order visit every conflicting pair in the same order: same labels, same
    Z_samples (this is what lets the GPU kernel be bit-exact)."""
    V = make_problem(F, N, 3, 19, na)
    Z = oracle.random_labels(V, K, 6)
    a = oracle.gibbs_dirber_finite_keyed(V, Z, K, 1.0, 1.0, 0.5, iter=9, burnin=0.4, seed=12)
    b = oracle.gibbs_dirber_finite_keyed(V, Z, K, 1.0, 1.0, 0.5, iter=9, burnin=0.4, seed=12, wavefront=True)
    assert a["written"] == b["written"] == 6
    assert np.array_equal(a["Z"], b["Z"]) and np.array_equal(a["Z_samples"], b["Z_samples"])
    obs = V != NA
    assert np.all((a["Z"][obs] >= 0) & (a["Z"][obs] < K))


@pytest.mark.parametrize("F,N", [(17, 29), (40, 12)])
def test_keyed_sibling_samplers_are_order_independent(F, N):
    """DP (Kmax = 200, first empty component gets gamma/2) and Dir-Dir (two assignments per entry): the
    column-major sweep and the anti-diagonal order give the same labels with keyed uniforms."""
    V = make_problem(F, N, 3, 29)
    Z = oracle.random_labels(V, 4, 3)
    a = oracle.gibbs_dirber_dp_keyed(V, Z, 0.8, iter=8, burnin=0.5, seed=5)
    b = oracle.gibbs_dirber_dp_keyed(V, Z, 0.8, iter=8, burnin=0.5, seed=5, wavefront=True)
    assert a["written"] == b["written"] == 4 and np.array_equal(a["Z_samples"], b["Z_samples"])
    assert a["Z"].min() >= 0 and a["Z"].max() < 200 and a["Z"].max() >= 4      # new components were opened
    Vn = make_problem(F, N, 3, 31, 0.15)
    Zn = oracle.random_labels(Vn, 5, 3)
    c = oracle.gibbs_dirdir_keyed(Vn, Zn, 5, 1.0, 0.7, iter=8, burnin=0.25, seed=6)
    d = oracle.gibbs_dirdir_keyed(Vn, Zn, 5, 1.0, 0.7, iter=8, burnin=0.25, seed=6, wavefront=True)
    assert c["written"] == d["written"] == 6
    assert np.array_equal(c["Z_samples"], d["Z_samples"]) and np.array_equal(c["C_samples"], d["C_samples"])
    obs = Vn != NA
    ones = obs & (Vn == 1); zeros = obs & (Vn == 0)
    assert np.array_equal(c["Z"][ones], c["C"][ones])          # V = 1: drawn together (:481-497)
    assert np.all(c["Z"][zeros] != c["C"][zeros])              # V = 0: the two components differ (:505,522)
    with pytest.raises(ValueError):
        oracle.gibbs_dirber_dp_keyed(Vn, Zn, 1.0, iter=3)


def test_keyed_and_stream_collapsed_gibbs_sample_the_same_posterior():
    """The keyed uniforms replace the sequential stream, nothing else: the posterior mean of V (label
    invariant) of the two chains agrees within the seed-to-seed spread of either."""
    F, N, K = 20, 30, 3
    V = make_problem(F, N, 3, 23)
    ev_s, ev_k = [], []
    for seed in (1, 2, 3):
        Z = oracle.random_labels(V, K, seed)
        a = oracle.gibbs_dirber_finite(V, Z, K, iter=300, burnin=0.5, seed=seed)
        ev_s.append(oracle.gibbs_expectation(V, a["Z_samples"][:, :, :a["written"]], K, 1.0, 1.0, 1.0)[2])
        b = oracle.gibbs_dirber_finite_keyed(V, Z, K, iter=300, burnin=0.5, seed=seed)
        ev_k.append(oracle.gibbs_expectation(V, b["Z_samples"][:, :, :b["written"]], K, 1.0, 1.0, 1.0)[2])
    rms_between = np.sqrt(np.mean((np.mean(ev_s, 0) - np.mean(ev_k, 0)) ** 2))
    rms_within = np.sqrt(np.mean((ev_s[0] - ev_s[1]) ** 2))
    assert rms_between < 2.0 * max(rms_within, 0.01)


def test_collapsed_vs_uncollapsed_gibbs_agree():
    """Different chains, same posterior: predictive LL per held-out entry and E_V
    agree within the seed-to-seed spread (SURVEY.md Appendix E probe 4)."""
    F, N, K = 30, 40, 3
    V = make_problem(F, N, 3, 2, 0.0, a=0.5, b=0.5)
    rng = np.random.default_rng(0)
    test = rng.random(V.shape) < 0.2
    Vtr = V.copy(order="F"); Vtr[test] = NA
    Vte = np.full_like(V, NA); Vte[test] = V[test]
    nte = test.sum()
    ll_c, ll_u, ev_c, ev_u = [], [], [], []
    for s in (1, 2, 3):
        Z = oracle.random_labels(Vtr, K, s)
        c = oracle.gibbs_dirber_finite(Vtr, Z, K, iter=400, burnin=0.5, seed=s)
        _, _, EV = oracle.gibbs_expectation(Vtr, c["Z_samples"][:, :, :c["written"]], K, 1.0, 1.0, 1.0)
        ll_c.append(oracle.loglikelihood_EV(EV, Vte) / nte); ev_c.append(EV)
        u = oracle.gibbs_uncollapsed(Vtr, Z, K, iter=1500, burnin=0.5, seed=s)
        ll_u.append(oracle.loglikelihood_EV(u["E_V"], Vte) / nte); ev_u.append(u["E_V"])
    spread = max(np.ptp(ll_c), np.ptp(ll_u), 0.004)
    assert abs(np.mean(ll_c) - np.mean(ll_u)) < 3 * spread
    rms_between = np.sqrt(np.mean((np.mean(ev_c, 0) - np.mean(ev_u, 0)) ** 2))
    rms_within = np.sqrt(np.mean((ev_c[0] - ev_c[1]) ** 2))
    assert rms_between < 2.5 * max(rms_within, 0.01)


def test_loglikelihood_matches_numpy():
    V = make_problem(15, 11, 2, 6, 0.3)
    rng = np.random.default_rng(1)
    W = rng.dirichlet(np.ones(4), size=15); H = rng.random((11, 4))
    ll, n = oracle.loglikelihood(W, H, V)
    P = W @ H.T
    obs = V != NA
    ref = np.log(np.where(V == 1, P, 1 - P)[obs]).sum()
    assert n == obs.sum() and abs(ll - ref) < 1e-10
