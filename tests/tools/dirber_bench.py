"""Time per sweep of the exact CVB0 wavefront (VB_DirBer_Rcpp, nbmf_vb_dirber EXACT_WAVEFRONT) at the C1
and C2 shapes, beside the oracle's single-threaded restatement of the reference loop."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nbmf_b200 import Engine, _lib  # noqa: E402
from oracle import oracle  # noqa: E402

for (F, N, K, it) in ((200, 300, 3, 200), (100, 5429, 10, 30), (784, 2000, 16, 8)):
    V, _, _ = oracle.generate(F, N, 3, 0.5, 0.5, 1)
    V = np.asfortranarray(V)
    Z = oracle.random_labels(V, K + 1, 2)
    e = Engine(precision=_lib.PREC_FP64)
    e.set_data(V); e.init_from_labels(Z, K + 1)
    W, H, _ = e.vb_dirber(K, 1.0, 1.0, 1.0, it)
    tm = e.timing()
    t1 = time.perf_counter()
    r = oracle.vb_dirber(V, Z, K, 1.0, 1.0, 1.0, iter=it)
    dc = time.perf_counter() - t1
    g = tm["total_ms"] / (it - 1)
    same = np.array_equal(W, r["E_W"], equal_nan=True) and np.array_equal(H, r["E_H"], equal_nan=True)
    print(f"CVB0 wavefront {F}x{N} Kc={K + 1}: {g:.3f} ms/sweep on the GPU ({F * N * (K + 1) / g * 1e3:.3e} entries*K/s), "
          f"oracle {1e3 * dc / (it - 1):.3f} ms/sweep (1 thread); bit-identical: {same}", flush=True)
