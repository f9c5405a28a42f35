"""Error of the tensor modes vs the FP64 path as a function of the iteration count (checkpoints
of the statistics inside one VB session)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbmf_b200 import Engine, _lib
F = N = int(sys.argv[1]); K = int(sys.argv[2]); cps = [int(x) for x in sys.argv[3].split(",")]
gamma = 1.0 / K
rel = lambda x, y: np.linalg.norm(x - y) / np.linalg.norm(y)
def run(prec):
    e = Engine(precision=prec)
    e.generate_synthetic(F, N, 8, 1.0, 1.0, 3); e.init_random_labels(K, 4)
    e.vb_begin(K, 1.0, 1.0, gamma)
    out = {}
    for i in range(1, max(cps) + 1):
        e.vb_step(True)
        if i in cps:
            nt, fo, fz = e.get_stats(K)
            g = gamma + nt
            out[i] = (g / g.sum(1, keepdims=True), (1 + fo) / (2 + fo + fz))
    _, _, lb = e.vb_end(max(cps), want_outputs=False)
    e.close()
    return out, lb
ref, lb0 = run(_lib.PREC_FP64)
for name, prec in (("tensor", _lib.PREC_TENSOR), ("fast", _lib.PREC_TENSOR_FAST)):
    got, lb = run(prec)
    for i in cps:
        print(f"F=N={F} K={K} {name:6s} after {i:3d} updates: relF E_W {rel(got[i][0], ref[i][0]):.2e} E_H {rel(got[i][1], ref[i][1]):.2e} "
              f"ELBO(max over first {i}) {np.max(np.abs(lb[:i] - lb0[:i]) / np.abs(lb0[:i])):.2e}", flush=True)
