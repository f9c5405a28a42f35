"""Accuracy / speed of NBMF_PREC_TENSOR_FAST (hi-only stage 2) vs NBMF_PREC_TENSOR vs the FP64 path."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbmf_b200 import Engine, _lib
F = N = int(sys.argv[1]); K = int(sys.argv[2]); iters = int(sys.argv[3])
rel = lambda x, y: np.linalg.norm(x - y) / np.linalg.norm(y)
res = {}
for name, prec in (("fp64", _lib.PREC_FP64), ("tensor", _lib.PREC_TENSOR), ("fast", _lib.PREC_TENSOR_FAST)):
    e = Engine(precision=prec)
    e.generate_synthetic(F, N, 8, 1.0, 1.0, 3); e.init_random_labels(K, 4)
    W, H, lb, _ = e.vb_aspect(K, 1.0, 1.0, 1.0 / K, iter=iters, checklb=True)
    t = e.timing(); e.close()
    res[name] = (W, H, lb)
    print(f"{name}: rows {t['pass_rows_ms']:.2f} ms cols {t['pass_cols_ms']:.2f} ms total/iter {t['total_ms']/iters:.2f}", flush=True)
for name in ("tensor", "fast"):
    W, H, lb = res[name]; W0, H0, lb0 = res["fp64"]
    print(f"F=N={F} K={K} it={iters} {name} vs fp64: relF E_W {rel(W, W0):.2e} E_H {rel(H, H0):.2e} ELBO {np.max(np.abs(lb - lb0) / np.abs(lb0)):.2e} "
          f"max-rel(E_W, floor 1e-6) {np.max(np.abs(W - W0) / np.maximum(W0, 1e-6)):.2e}", flush=True)
