"""nbmf_vb_aspect_bits: the single-call host-memory form of VB_Aspect_Rcpp whose upload is
pipelined with the first iteration.  It must agree with the staged calls (set_data_bits +
set_stats + vb_aspect) and with the oracle."""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu


def _pack_cols(V):
    """F x N 0/1 matrix -> [N][ceil(F/32)] uint32, bit f%32 of word f//32 (nbmf_set_data_bits layout)."""
    F, N = V.shape
    fw = (F + 31) // 32
    pad = np.zeros((fw * 32, N), dtype=np.uint8)
    pad[:F] = V
    return np.packbits(pad.T.reshape(N, fw, 32), axis=2, bitorder="little").view(np.uint32).reshape(N, fw)


@pytest.mark.parametrize("F,N,K,blk,prec", [(512, 1024, 64, 256, 2), (640, 768, 40, 128, 2), (512, 2048, 128, 512, 4),
                                             (384, 1000, 64, 256, 4)])
def test_pipelined_call_matches_staged_calls_and_oracle(F, N, K, blk, prec):
    from nbmf_b200 import Engine
    V, _, _ = oracle.generate(F, N, 4, 0.8, 0.8, 11)
    V = np.asfortranarray(V)
    Z = oracle.random_labels(V, K, 12)
    bits = _pack_cols(V)
    iters = 2
    # staged reference on the same engine settings
    e0 = Engine(precision=prec)
    e0.set_data(V); e0.init_from_labels(Z, K)
    stats = e0.get_stats(K)
    W0, H0, lb0, _ = e0.vb_aspect(K, 1.0, 1.0, 0.5, iter=iters, checklb=True)
    e1 = Engine(precision=prec, pipe_block_cols=blk)
    W1, H1, lb1 = e1.vb_aspect_bits(bits, F, N, K, stats, 1.0, 1.0, 0.5, iter=iters, checklb=True)
    # the pipelined call cuts the reductions into other chunks than the staged one: same arithmetic, other fp32 chains
    relF = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    assert relF(W1, W0) < 2e-6 and relF(H1, H0) < 2e-6
    assert np.allclose(lb1, lb0, rtol=1e-6)
    # a second call on the same handle (buffers reused, same shape) gives the same answer
    W2, H2, lb2 = e1.vb_aspect_bits(bits, F, N, K, stats, 1.0, 1.0, 0.5, iter=iters, checklb=True)
    assert np.array_equal(W2, W1) and np.array_equal(H2, H1) and np.allclose(lb2, lb1, rtol=1e-12)   # the ELBO sum uses fp64 atomics
    # and the oracle
    o = oracle.vb_aspect(V, Z, K, 1.0, 1.0, 0.5, iter=iters, checklb=True)
    Wo, Ho, lbo = o["E_W"], o["E_H"], o["lowerbounds"]
    tol = 1e-5      # the north star's gate, norm-wise (hi+lo variants; the hi-only one is for 10^4 rows and more)
    assert relF(W1, Wo) < tol and relF(H1, Ho) < tol, (relF(W1, Wo), relF(H1, Ho))
    assert abs(lb1[-1] - lbo[-1]) / abs(lbo[-1]) < 1e-5


def test_pipelined_call_falls_back_on_the_fp64_path():
    from nbmf_b200 import Engine, _lib
    F, N, K = 96, 200, 5
    V, _, _ = oracle.generate(F, N, 3, 0.8, 0.8, 5)
    V = np.asfortranarray(V)
    Z = oracle.random_labels(V, K, 6)
    e0 = Engine(precision=_lib.PREC_FP64)
    e0.set_data(V); e0.init_from_labels(Z, K)
    stats = e0.get_stats(K)
    e1 = Engine(precision=_lib.PREC_FP64, pipe_block_cols=64)
    W1, H1, lb1 = e1.vb_aspect_bits(_pack_cols(V), F, N, K, stats, 1.0, 1.0, 1.0, iter=4, checklb=True)
    o = oracle.vb_aspect(V, Z, K, 1.0, 1.0, 1.0, iter=4, checklb=True)
    Wo, Ho, lbo = o["E_W"], o["E_H"], o["lowerbounds"]
    assert np.allclose(W1, Wo, rtol=1e-9) and np.allclose(H1, Ho, rtol=1e-9) and np.allclose(lb1, lbo, rtol=1e-10)
