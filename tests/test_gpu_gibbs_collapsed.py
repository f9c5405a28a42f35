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
