"""Does the chunk length (number of MMA accumulations into one TMEM accumulator) change the result?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbmf_b200 import Engine, _lib
F = N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
K = 128
res = {}
for chunk in (32, 2048):
    e = Engine(precision=_lib.PREC_TENSOR, flush_interval=chunk)
    e.generate_synthetic(F, N, 8, 1.0, 1.0, 3); e.init_random_labels(K, 4)
    W, H, lb, _ = e.vb_aspect(K, 1.0, 1.0, 1.0 / K, iter=6, checklb=True)
    res[chunk] = (W, H, lb); print(chunk, e.timing()["pass_rows_ms"], flush=True); e.close()
rel = lambda x, y: np.linalg.norm(x - y) / np.linalg.norm(y)
print("chunk 32 vs 2048: relF E_W %.2e E_H %.2e ELBO %.2e" % (rel(res[32][0], res[2048][0]), rel(res[32][1], res[2048][1]),
      np.max(np.abs(res[32][2] - res[2048][2]) / np.abs(res[2048][2]))))
if F <= 16384:
    e = Engine(precision=_lib.PREC_FP64)
    e.generate_synthetic(F, N, 8, 1.0, 1.0, 3); e.init_random_labels(K, 4)
    W, H, lb, _ = e.vb_aspect(K, 1.0, 1.0, 1.0 / K, iter=6, checklb=True); e.close()
    for c in res: print("chunk", c, "vs fp64: E_W %.2e E_H %.2e ELBO %.2e" % (rel(res[c][0], W), rel(res[c][1], H), np.max(np.abs(res[c][2]-lb)/np.abs(lb))))
