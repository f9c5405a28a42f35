"""Time per sweep of the collapsed wavefront sampler (nbmf_gibbs_dirber_collapsed) at the C1 and
C2 shapes, beside the oracle's single-threaded restatement of the reference loop."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nbmf_b200 import Engine, _lib  # noqa: E402
from oracle import oracle  # noqa: E402

for (F, N, K, it) in ((200, 300, 3, 200), (100, 5429, 10, 50), (784, 2000, 16, 20)):
    V, _, _ = oracle.generate(F, N, 3, 0.5, 0.5, 1)
    V = np.asfortranarray(V)
    Z = oracle.random_labels(V, K, 2)
    e = Engine(precision=_lib.PREC_FP64)
    e.set_data(V); e.init_from_labels(Z, K)
    e.gibbs_dirber_collapsed(K, 1.0, 1.0, 1.0, it, 0.5, 1, want_samples=False)
    tm = e.timing()
    t1 = time.perf_counter()
    oracle.gibbs_dirber_finite_keyed(V, Z, K, 1.0, 1.0, 1.0, iter=it, burnin=0.5, seed=1)
    dc = time.perf_counter() - t1
    g = tm["total_ms"] / (it - 1)
    print(f"collapsed wavefront {F}x{N} K={K}: {g:.3f} ms/sweep on the GPU ({F * N * K / g * 1e3:.3e} entries*K/s), "
          f"oracle {1e3 * dc / (it - 1):.3f} ms/sweep (1 thread, {F * N * K * (it - 1) / dc:.3e} entries*K/s)", flush=True)
