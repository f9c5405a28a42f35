"""Pins the oracle to the reference: (1) against tests/golden/ref_vectors.npz, produced by the
reference's OWN source files compiled against oracle/ref_shim (tests/golden/make_golden.py);
(2) live against oracle/_ref/libnbmf_ref.so when it has been built (build container only --
/root/reference does not exist on the GPU box).  Equality is bit-exact: the oracle is a
loop-for-loop restatement, and the shim's RNG is the oracle's."""
import os

import numpy as np
import pytest

from oracle import oracle, ref

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors.npz"))
CASES = ["a", "b", "c"]


@pytest.mark.parametrize("c", CASES)
def test_vb_aspect_matches_reference_golden(c):
    K, al, be, ga, it = G[f"aspect_{c}_par"]
    r = oracle.vb_aspect(G[f"aspect_{c}_V"], G[f"aspect_{c}_Z"], int(K), al, be, ga, iter=int(it), checklb=True,
                         return_E_Z=True)
    for k in ("E_W", "E_H", "lowerbounds", "E_Z"):
        assert np.array_equal(r[k], G[f"aspect_{c}_{k}"]), k
    # and the E_Z-free batch form the engine computes agrees with the reference to rounding
    b = oracle.vb_aspect_batch(G[f"aspect_{c}_V"], G[f"aspect_{c}_Z"], int(K), al, be, ga, iter=int(it), checklb=True)
    assert np.max(np.abs(b["E_W"] - G[f"aspect_{c}_E_W"]) / G[f"aspect_{c}_E_W"]) < 1e-11
    assert np.max(np.abs(b["lowerbounds"] - G[f"aspect_{c}_lowerbounds"]) / np.abs(G[f"aspect_{c}_lowerbounds"])) < 1e-12


@pytest.mark.parametrize("c", CASES)
def test_vb_dirber_matches_reference_golden(c):
    K, al, be, ga, it = G[f"aspect_{c}_par"]
    r = oracle.vb_dirber(G[f"aspect_{c}_V"], G[f"dirber_{c}_Z"], int(K), al, be, ga, iter=max(int(it) // 2, 2),
                         return_E_Z=True)
    for k in ("E_W", "E_H", "E_Z"):
        assert np.array_equal(r[k], G[f"dirber_{c}_{k}"], equal_nan=True), k


@pytest.mark.parametrize("c", CASES)
def test_gibbs_matches_reference_golden_draw_for_draw(c):
    K, al, be, ga, _ = G[f"aspect_{c}_par"]
    V = G[f"aspect_{c}_V"]
    g = oracle.gibbs_dirber_finite(V, G[f"aspect_{c}_Z"], int(K), al, be, ga, iter=21, burnin=0.5, seed=7)
    Zs = G[f"gibbs_{c}_Z_samples"]
    assert np.array_equal(g["Z_samples"], Zs)
    # compute_expectation_W_Rcpp / _H_Rcpp incl. the zero last column (collapsed_gibbs_z.cpp:236-242)
    L = oracle.lib()
    F, N, J = Zs.shape
    Kq = int(Zs.max()) + 1
    EW = np.zeros((F, Kq), order="F"); EH = np.zeros((N, Kq), order="F")
    Zf = np.asfortranarray(Zs)
    L.oracle_compute_expectation_W(oracle._p(Zf, oracle._ip), F, N, J, ga, Kq, oracle._p(EW))
    L.oracle_compute_expectation_H(oracle._p(Zf, oracle._ip), oracle._p(oracle._i32(V), oracle._ip), F, N, J, al, be, Kq,
                                   oracle._p(EH))
    assert np.allclose(EW, G[f"gibbs_{c}_E_W"], rtol=1e-13, atol=0)
    assert np.allclose(EH, G[f"gibbs_{c}_E_H"], rtol=1e-13, atol=0)


@pytest.mark.parametrize("c", CASES)
def test_sibling_samplers_match_reference_golden_draw_for_draw(c):
    """Gibbs_DirDir_finite_Rcpp (:441-557) and Gibbs_DirBer_DP_Rcpp (:275-359) on the sequential stream: the
    oracle's restatements reproduce the sample cubes of the reference's own code.  The keyed forms the engine
    is tested against share every line with these except the source of the uniform."""
    K, al, be, ga, _ = G[f"aspect_{c}_par"]
    V = G[f"aspect_{c}_V"]; Z = G[f"aspect_{c}_Z"]
    obs = V != oracle.NA
    d = oracle.gibbs_dirdir_keyed(V, Z, int(K), al, ga, iter=13, burnin=0.5, seed=9, wavefront=2)
    assert np.array_equal(d["Z_samples"][obs], G[f"dirdir_{c}_Z_samples"][obs])
    assert np.array_equal(d["C_samples"][obs], G[f"dirdir_{c}_C_samples"][obs])
    if f"dp_{c}_Z_samples" in G:
        p = oracle.gibbs_dirber_dp_keyed(V, Z, ga, iter=13, burnin=0.5, seed=9, wavefront=2)
        assert np.array_equal(p["Z_samples"], G[f"dp_{c}_Z_samples"])


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_live_reference_build_equals_oracle():
    V, _, _ = oracle.generate(37, 29, 3, 0.4, 0.4, 99)
    V = V.copy(order="F"); V[np.random.default_rng(3).random(V.shape) < 0.25] = oracle.NA
    K = 5
    Z = oracle.random_labels(V, K, 2)
    a = oracle.vb_aspect(V, Z, K, 0.9, 1.3, 0.2, iter=30, checklb=True, return_E_Z=True)
    b = ref.vb_aspect(V, Z, K, 0.9, 1.3, 0.2, iter=30, checklb=True, return_E_Z=True)
    for k in ("E_W", "E_H", "lowerbounds", "E_Z"):
        assert np.array_equal(a[k], b[k]), k
    Z1 = oracle.random_labels(V, K + 1, 2)
    a = oracle.vb_dirber(V, Z1, K, 0.9, 1.3, 0.2, iter=9, return_E_Z=True)
    b = ref.vb_dirber(V, Z1, K, 0.9, 1.3, 0.2, iter=9, return_E_Z=True)
    for k in ("E_W", "E_H", "E_Z"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    g = oracle.gibbs_dirber_finite(V, Z, K, iter=25, burnin=0.4, seed=11)
    h = ref.gibbs_dirber_finite(V, Z, K, iter=25, burnin=0.4, seed=11)
    assert np.array_equal(g["Z_samples"], h["Z_samples"])
    # lower_bound_aspect of the reference on the reference's own E_Z == the trace value
    assert not b["stopped"]
    # sibling samplers on the sequential stream
    obs = V != oracle.NA
    d = oracle.gibbs_dirdir_keyed(V, Z, K, 1.2, 0.4, iter=15, burnin=0.3, seed=5, wavefront=2)
    r = ref.gibbs_dirdir(V, Z, K, 1.2, 0.4, iter=15, burnin=0.3, seed=5)
    assert np.array_equal(d["Z_samples"][obs], r["Z_samples"][obs]) and np.array_equal(d["C_samples"][obs], r["C_samples"][obs])
    Vf, _, _ = oracle.generate(31, 23, 3, 0.4, 0.4, 5)
    Zf = oracle.random_labels(Vf, 6, 8)
    p = oracle.gibbs_dirber_dp_keyed(Vf, Zf, 1.7, iter=15, burnin=0.3, seed=5, wavefront=2)
    q = ref.gibbs_dirber_dp(Vf, Zf, 1.0, 1.0, 1.7, iter=15, burnin=0.3, seed=5)
    assert np.array_equal(p["Z_samples"], q["Z_samples"]) and p["Z_samples"].max() >= 6


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_reference_sample_gibbs_cpp_is_dead_code():
    """sample_gibbs.cpp never stores its draw (:60-69): after the first sweep C is all zero, so
    the returned mean is C_init / (number of kept sweeps + 1)."""
    F, K = 12, 3
    rng = np.random.default_rng(0)
    W = rng.random((F, K)); v = rng.integers(0, 2, F)
    C0 = np.zeros((F, K), dtype=np.int32); C0[np.arange(F), rng.integers(0, K, F)] = 1
    out = ref.sample_gibbs_cpp(v, W, C0, alpha=1.0, iter=11, burnin=0.5, seed=3)
    assert np.allclose(out, C0 / 7.0)
