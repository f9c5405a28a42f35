"""GPU probe for the tcgen05 contraction pass: checks the raw stage-1 accumulators of the
first tile against NumPy, then whole VB iterations against the FP64 path.  Debug aid."""
import ctypes as C
import os
import sys

import numpy as np
import scipy.special as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nbmf_b200 import Engine, _lib  # noqa: E402
from oracle import oracle  # noqa: E402

NA = oracle.NA


def probe(F, N, K, na=0.1, gamma=0.5, iters=3, seed=3):
    print(f"=== probe F={F} N={N} K={K} na={na}", flush=True)
    V, _, _ = oracle.generate(F, N, 4, 0.8, 0.8, seed)
    V = V.copy(order="F")
    if na > 0:
        V[np.random.default_rng(seed).random(V.shape) < na] = NA
    Z = oracle.random_labels(V, K, seed + 1)
    et = Engine(precision=_lib.PREC_TENSOR)
    et.set_data(V); et.init_from_labels(Z, K)
    nt, fo, fz = et.get_stats(K)
    g, a, b = gamma + nt, 1.0 + fo, 1.0 + fz
    A = np.exp(sp.digamma(g) - sp.digamma(g.sum(1, keepdims=True)))
    B1 = np.exp(sp.digamma(a) - sp.digamma(a + b)); B0 = np.exp(sp.digamma(b) - sp.digamma(a + b))
    BB = np.empty((2 * N, K)); BB[1::2] = B1; BB[0::2] = B0
    lib = _lib.load()
    lib.nbmf_debug_tp_dump.restype = C.c_int
    lib.nbmf_debug_tp_dump.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    et.vb_begin(K, 1.0, 1.0, gamma)
    # params + conversion happen inside vb_step; run one step, then re-run the passes for the dump
    et.init_from_labels(Z, K)
    et.vb_begin(K, 1.0, 1.0, gamma)
    et.vb_step(False)
    et.synchronize()
    # after one step the operands A/BB on the device are those of the INITIAL statistics
    for rows in (1, 0):
        S = np.zeros((128, 256), dtype=np.float32); ex = np.zeros(2, dtype=np.int32)
        rc = lib.nbmf_debug_tp_dump(et.h, rows, S.ctypes.data, ex.ctypes.data)
        if rc != 0:
            print("  dump rc", rc, lib.nbmf_last_error(et.h).decode()); continue
        sc = 2.0 ** (int(ex[0]) + int(ex[1]))
        if rows:
            U = A[:128]; W = BB[:256]
        else:
            U = BB[:128]; W = A[:256]
        U = np.pad(U, ((0, max(0, 128 - U.shape[0])), (0, 0))); W = np.pad(W, ((0, max(0, 256 - W.shape[0])), (0, 0)))
        ref = sc * (U @ W.T)
        err = np.abs(S - ref) / np.maximum(np.abs(ref), 1e-30)
        print(f"  stage-1 S ({'rows' if rows else 'cols'} pass): exps={ex.tolist()} max rel err {err.max():.3e} "
              f"median {np.median(err):.3e}  S[0,:4]={S[0,:4]} ref={ref[0,:4]}", flush=True)
        if err.max() > 1e-3:
            bad = np.argwhere(err > 1e-3)
            print("   first bad (lane, col):", bad[:8].tolist(), " n_bad", len(bad))
    try:
        et.vb_end(0, want_outputs=False)
    except Exception as e:
        print("  vb_end:", e)
    # whole iterations vs the FP64 path
    ef = Engine(precision=_lib.PREC_FP64)
    ef.set_data(V); ef.init_from_labels(Z, K)
    Wf, Hf, lf, _ = ef.vb_aspect(K, 1.0, 1.0, gamma, iter=iters, checklb=True)
    sf = ef.get_stats(K)
    et.init_from_labels(Z, K)
    try:
        Wt, Ht, lt, _ = et.vb_aspect(K, 1.0, 1.0, gamma, iter=iters, checklb=True)
    except Exception as e:
        print("  tensor vb_aspect failed:", e); return
    st = et.get_stats(K)
    rel = lambda x, y: np.linalg.norm(x - y) / np.linalg.norm(y)
    print(f"  {iters} iters: relF E_W {rel(Wt, Wf):.3e} E_H {rel(Ht, Hf):.3e} ELBO rel {np.max(np.abs(lt - lf) / np.abs(lf)):.3e}")
    print(f"  stats relF: n_total {rel(st[0], sf[0]):.3e} f_ones {rel(st[1], sf[1]):.3e} f_zeros {rel(st[2], sf[2]):.3e}"
          f"  sum n_total {st[0].sum():.6f} vs {sf[0].sum():.6f}", flush=True)
    print("  timing", et.timing())
    et.close(); ef.close()


if __name__ == "__main__":
    probe(256, 256, 64, na=0.0)
    probe(256, 256, 128, na=0.0)
    probe(300, 520, 64, na=0.15)
    probe(1000, 777, 128, na=0.1)
    probe(4096, 4096, 128, na=0.0, iters=5)
