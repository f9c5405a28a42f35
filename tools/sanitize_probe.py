"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_probe.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbmf_b200 import Engine, _lib  # noqa: E402

rng = np.random.default_rng(0)
F, N, K = 384, 640, 48
V = (rng.random((F, N)) < 0.4).astype(np.int32)
V[rng.random((F, N)) < 0.1] = np.iinfo(np.int32).min
V = np.asfortranarray(V)
Z = np.asfortranarray(rng.integers(0, K, size=(F, N), dtype=np.int32)); Z[V < 0] = np.iinfo(np.int32).min
for prec in (_lib.PREC_TENSOR, _lib.PREC_TENSOR_FAST, _lib.PREC_FP64):
    e = Engine(precision=prec)
    e.set_data(V); e.init_from_labels(Z, K)
    W, H, lb, _ = e.vb_aspect(K, 1.0, 1.0, 0.5, iter=2, checklb=True)
    print("vb_aspect prec", prec, "ok", float(lb[-1]), flush=True)
    e.close()
# pipelined single call (no NA)
Vb = (rng.random((F, N)) < 0.4).astype(np.uint8)
fw = (F + 31) // 32
pad = np.zeros((fw * 32, N), dtype=np.uint8); pad[:F] = Vb
bits = np.packbits(pad.T.reshape(N, fw, 32), axis=2, bitorder="little").view(np.uint32).reshape(N, fw)
e = Engine(precision=_lib.PREC_TENSOR, pipe_block_cols=128)
e.set_data(np.asfortranarray(Vb.astype(np.int32))); e.init_random_labels(K, 3)
st = e.get_stats(K)
W, H, lb = e.vb_aspect_bits(bits, F, N, K, st, 1.0, 1.0, 0.5, iter=2, checklb=True)
print("vb_aspect_bits ok", float(lb[-1]), flush=True)
e.close()
# exact wavefronts
Fs, Ns, Ks = 40, 70, 5
Vs = np.asfortranarray(V[:Fs, :Ns]); Zs = np.asfortranarray(Z[:Fs, :Ns] % (Ks + 1)); Zs[Vs < 0] = np.iinfo(np.int32).min
e = Engine(precision=_lib.PREC_FP64)
e.set_data(Vs); e.init_from_labels(Zs, Ks + 1)
e.vb_dirber(Ks, 1.0, 1.0, 1.0, 4)
Zg = np.asfortranarray(Zs % Ks); Zg[Vs < 0] = np.iinfo(np.int32).min
e.init_from_labels(Zg, Ks)
e.gibbs_dirber_collapsed(Ks, 1.0, 1.0, 1.0, 5, 0.5, 1)
e.gibbs_dirber(Ks, 1.0, 1.0, 1.0, 6, 0.5, 1)
print("wavefronts + gibbs ok", flush=True)
e.close()
