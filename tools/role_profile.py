"""Cycle accounting of the tensor pass per role (CTA 0 only).  Needs a library built with
`make -C nbmf_b200/csrc clean all EXTRA=-DTP_PROFILE`.  Debug aid, not part of the product."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbmf_b200 import Engine, _lib  # noqa: E402

F = N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
K = int(sys.argv[3]) if len(sys.argv) > 3 else 128
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lib = _lib.load()
lib.nbmf_debug_tp_prof.argtypes = [C.c_void_p, C.c_int]
flush = int(sys.argv[4]) if len(sys.argv) > 4 else 0
e = Engine(precision=prec, flush_interval=flush)
e.generate_synthetic(F, N, 16, 0.8, 0.8, 1)
e.init_random_labels(K, 2)
e.vb_begin(K, 1.0, 1.0, 1.0 / K)
for _ in range(2):
    e.vb_step(False)
e.synchronize()
lib.nbmf_debug_tp_prof(None, 1)
iters = 4
for _ in range(iters):
    e.vb_step(False)
e.synchronize()
out = np.zeros(64, dtype=np.uint64)
lib.nbmf_debug_tp_prof(out.ctypes.data, 0)
tm = e.timing()
# steps handled by CTA 0 per pass: items/148 * chunk steps  (approx: total steps / 148)
steps = iters * (((2 * N + 63) // 64) * ((F + 127) // 128) + ((F + 63) // 64) * ((2 * N + 127) // 128)) / 148.0
names = {0: ["empty_wait", "issue"], 1: ["uready", "sfree_wait", "full_wait", "mma_issue", "commit", "tail"],
         2: ["uready", "rready_wait", "mma_issue", "commits"],
         3: ["item_head", "bits", "sfull_wait", "tmem_ld", "math", "rfree_wait", "st+arrive", "fold"]}
print(f"F=N={F} K={K} prec={prec} flush={flush}: ms/iter {tm['total_ms'] / tm['iters']:.3f}; cycles per step (CTA 0), steps/CTA ~{steps:.0f}")
for role, label in ((0, "producer"), (1, "mma1"), (2, "mma2"), (3, "epi g0"), (4, "epi g1"), (5, "epi g2"), (6, "epi g3")):
    v = out[role * 8:(role + 1) * 8].astype(np.float64)
    div = steps if role < 3 else steps / 4
    nm = names[min(role, 3)]
    print(f"  {label:9s} total {v.sum() / div:8.1f} | " + "  ".join(f"{n} {x / div:7.1f}" for n, x in zip(nm, v)))
