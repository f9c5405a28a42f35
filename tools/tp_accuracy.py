"""Accuracy of the tensor path vs the FP64 path as a function of size and chunk length."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbmf_b200 import Engine, _lib  # noqa: E402


def run(F, N, K, iters, chunk, gamma, prec=_lib.PREC_TENSOR, seed=3):
    os.environ["NBMF_TP_CHUNK_STEPS"] = str(chunk)
    e = Engine(precision=prec)
    e.generate_synthetic(F, N, 8, 1.0, 1.0, seed)
    e.init_random_labels(K, seed + 1)
    t0 = time.time()
    W, H, lb, _ = e.vb_aspect(K, 1.0, 1.0, gamma, iter=iters, checklb=True)
    dt = time.time() - t0
    tim = e.timing()
    e.close()
    return W, H, lb, tim, dt


rel = lambda x, y: np.linalg.norm(x - y) / np.linalg.norm(y)
cases = [(512, 512, 64, 10), (1024, 1024, 128, 10), (2048, 2048, 128, 10), (4096, 4096, 128, 10),
         (8192, 8192, 64, 10), (8192, 8192, 128, 20)]
if len(sys.argv) > 1:
    cases = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
for F, N, K, iters in cases:
    gamma = 1.0 / K
    Wf, Hf, lf, tf, _ = run(F, N, K, iters, 32, gamma, prec=_lib.PREC_FP64)
    for chunk in (4, 16, 32):
        Wt, Ht, lt, tt, _ = run(F, N, K, iters, chunk, gamma)
        mx = np.max(np.abs(Wt - Wf) / np.maximum(Wf, 1e-6))
        print(f"F={F} N={N} K={K} it={iters} chunk={chunk:3d}: relF E_W {rel(Wt, Wf):.2e} E_H {rel(Ht, Hf):.2e} "
              f"ELBO {np.max(np.abs(lt - lf) / np.abs(lf)):.2e} maxrel(E_W,floor1e-6) {mx:.2e} | "
              f"rows {tt['pass_rows_ms']:.3f} ms cols {tt['pass_cols_ms']:.3f} ms params {tt['params_ms']:.3f} "
              f"(fp64 rows {tf['pass_rows_ms']:.1f} cols {tf['pass_cols_ms']:.1f})", flush=True)
