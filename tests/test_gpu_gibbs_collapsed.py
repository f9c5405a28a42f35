"""Collapsed Gibbs with the reference's own schedule (nbmf_gibbs_dirber_collapsed): integer work,
so the bar is bit-exact -- labels and Z_samples must equal the oracle's restatement of
Gibbs_DirBer_finite_Rcpp (collapsed_gibbs_z.cpp:363-437) run with the same per-entry key."""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu
NA = oracle.NA


def _case(F, N, Kt, K, na, seed):
    V, _, _ = oracle.generate(F, N, Kt, 0.5, 0.5, seed)
    V = np.asfortranarray(V)
    if na > 0:
        V[np.random.default_rng(seed).random(V.shape) < na] = NA
    Z = oracle.random_labels(V, K, seed + 1)
    return V, Z


@pytest.mark.parametrize("F,N,Kt,K,na,iter,burnin", [
    (200, 300, 3, 3, 0.0, 12, 0.5),      # C1 shape, one block
    (100, 400, 4, 10, 0.1, 9, 0.3),      # NA entries (unvotes-like), K = 10
    (37, 53, 2, 5, 0.2, 8, 0.0),         # ragged, burnin 0: every sweep kept, first slice stays zero (i starts at 1)
    (1100, 1200, 4, 4, 0.05, 3, 0.5),    # diagonals longer than one block: cooperative grid barrier
    (48, 70, 3, 40, 0.1, 5, 0.4),        # K = 40 > 32: lanes of a group take several k
    (300, 40, 3, 17, 0.0, 4, 0.5),       # 32 lanes per entry, several passes of one block per diagonal
])
def test_collapsed_sweep_is_bit_exact(F, N, Kt, K, na, iter, burnin):
    from nbmf_b200 import Engine, _lib
    V, Z = _case(F, N, Kt, K, na, 21)
    ref = oracle.gibbs_dirber_finite_keyed(V, Z, K, 1.0, 1.0, 0.5, iter=iter, burnin=burnin, seed=99)
    e = Engine(precision=_lib.PREC_FP64)
    e.set_data(V); e.init_from_labels(Z, K)
    Zs, nw, Zl = e.gibbs_dirber_collapsed(K, 1.0, 1.0, 0.5, iter, burnin, 99)
    obs = V != NA
    assert nw == ref["written"]
    assert np.array_equal(Zl[obs], ref["Z"][obs])
    assert Zs.shape == ref["Z_samples"].shape
    for j in range(Zs.shape[2]):
        assert np.array_equal(Zs[:, :, j][obs], ref["Z_samples"][:, :, j][obs]), j
    # a different seed is a different chain
    e2 = Engine(precision=_lib.PREC_FP64)
    e2.set_data(V); e2.init_from_labels(Z, K)
    _, _, Zl2 = e2.gibbs_dirber_collapsed(K, 1.0, 1.0, 0.5, iter, burnin, 100)
    assert (Zl2[obs] != Zl[obs]).mean() > 0.05


def test_reference_signature_collapsed_mode():
    from nbmf_b200 import Gibbs_DirBer_finite_Rcpp
    V, Z = _case(60, 90, 3, 4, 0.1, 5)
    out = Gibbs_DirBer_finite_Rcpp(V, Z, 4, 1.0, 1.0, 1.0, 10, 0.5, seed=3, collapsed=True)
    ref = oracle.gibbs_dirber_finite_keyed(V, Z, 4, 1.0, 1.0, 1.0, iter=10, burnin=0.5, seed=3)
    obs = V != NA
    assert out["written"] == ref["written"] == 5
    assert np.array_equal(out["Z_samples"][obs], ref["Z_samples"][obs])
    # labels stay in range and NA entries keep no label influence: counts of the last slice add up
    assert out["Z_last"][obs].min() >= 0 and out["Z_last"][obs].max() < 4


def test_collapsed_needs_labels_and_rejects_bad_labels():
    from nbmf_b200 import Engine, NbmfError, _lib
    V, Z = _case(40, 50, 2, 3, 0.0, 7)
    e = Engine(precision=_lib.PREC_FP64)
    e.set_data(V)
    with pytest.raises(NbmfError):
        e.gibbs_dirber_collapsed(3, 1.0, 1.0, 1.0, 5, 0.5, 1)
    e.init_from_labels(Z, 3)
    with pytest.raises(NbmfError):
        e.gibbs_dirber_collapsed(2, 1.0, 1.0, 1.0, 5, 0.5, 1)   # labels up to 2 with K = 2


# ---------------------------------------------------------------------------
# sibling samplers: DP and Dir-Dir, same bit-exact bar
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("F,N,iter,burnin", [(60, 90, 8, 0.5), (200, 300, 5, 0.4), (1100, 1150, 3, 0.5)])
def test_dp_sampler_is_bit_exact(F, N, iter, burnin):
    from nbmf_b200 import Gibbs_DirBer_DP_Rcpp
    V, Z = _case(F, N, 3, 4, 0.0, 33)
    ref = oracle.gibbs_dirber_dp_keyed(V, Z, 0.8, iter=iter, burnin=burnin, seed=17)
    out = Gibbs_DirBer_DP_Rcpp(V, Z, 1.0, 1.0, 0.8, iter, burnin, seed=17)
    assert out["written"] == ref["written"]
    assert np.array_equal(out["Z_last"], ref["Z"]) and np.array_equal(out["Z_samples"], ref["Z_samples"])
    assert out["Z_last"].max() >= 4        # new components were opened (the CRP term, :325-329)


def test_dp_sampler_rejects_na():
    from nbmf_b200 import Gibbs_DirBer_DP_Rcpp, NbmfError
    V, Z = _case(30, 40, 2, 3, 0.2, 5)
    with pytest.raises(NbmfError):
        Gibbs_DirBer_DP_Rcpp(V, Z, 1.0, 1.0, 1.0, 4, 0.5)


@pytest.mark.parametrize("F,N,K,na,iter,burnin", [(60, 90, 5, 0.1, 8, 0.25), (200, 300, 3, 0.0, 5, 0.5),
                                                  (48, 70, 40, 0.1, 4, 0.5), (1100, 1150, 4, 0.05, 3, 0.5)])
def test_dirdir_sampler_is_bit_exact(F, N, K, na, iter, burnin):
    from nbmf_b200 import Gibbs_DirDir_finite_Rcpp
    V, Z = _case(F, N, 3, K, na, 41)
    ref = oracle.gibbs_dirdir_keyed(V, Z, K, 1.0, 0.7, iter=iter, burnin=burnin, seed=23)
    out = Gibbs_DirDir_finite_Rcpp(V, Z, K, 1.0, 0.7, iter, burnin, seed=23)
    obs = V != NA
    assert out["written"] == ref["written"]
    assert np.array_equal(out["Z_last"][obs], ref["Z"][obs]) and np.array_equal(out["C_last"][obs], ref["C"][obs])
    assert np.array_equal(out["Z_samples"][obs], ref["Z_samples"][obs])
    assert np.array_equal(out["C_samples"][obs], ref["C_samples"][obs])
    ones = obs & (V == 1); zeros = obs & (V == 0)
    assert np.array_equal(out["Z_last"][ones], out["C_last"][ones]) and np.all(out["Z_last"][zeros] != out["C_last"][zeros])


def test_bernoulli_nmf_model_matrix():
    """BernoulliNMF(method="Gibbs", model=...) (BernoulliNMF.R:41-71): VB warm start at K_init = 10, which.max
    hand-off, then the sampler of the model; the R wrappers' posterior means on the sample cubes."""
    from nbmf_b200 import BernoulliNMF
    V, _, _ = oracle.generate(40, 60, 3, 0.4, 0.4, 9)
    V = np.asfortranarray(V)
    r = BernoulliNMF(V, 12, model="DirDir", method="Gibbs", iter=12, burnin=0.5, seed=4)
    assert r["class"] == ("GibbsDirDir", "Bernoulli") and r["E_W"].shape[0] == 40 and r["E_H"].shape[0] == 60
    assert np.allclose(r["E_W"].sum(1), 1.0) and np.allclose(r["E_H"].sum(1), 1.0) and r["Z_samples"].shape == (40, 60, 6)
    p = BernoulliNMF(V, 12, model="DirBerDP", method="Gibbs", iter=12, burnin=0.5, gamma=0.5, seed=4)
    assert p["class"] == ("GibbsDirBer", "Bernoulli") and p["Z_samples"].shape == (40, 60, 6)
    assert np.allclose(p["E_W"].sum(1), 1.0) and np.all((p["E_H"] >= 0) & (p["E_H"] <= 1))
    # the same call twice is the same chain
    p2 = BernoulliNMF(V, 12, model="DirBerDP", method="Gibbs", iter=12, burnin=0.5, gamma=0.5, seed=4)
    assert np.array_equal(p2["Z_samples"], p["Z_samples"])
