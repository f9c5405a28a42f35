"""Where the end-to-end step (host buffers -> results) spends its time at C5.  Debug aid."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbmf_b200 import Engine  # noqa: E402

F = N = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
K = 128
e = Engine(precision=0)
e.generate_synthetic(F, N, 16, 0.8, 0.8, 1)
v, _ = e.get_data_bits(want_mask=False)
pv = torch.empty(v.shape, dtype=torch.int32, pin_memory=True)
pv.numpy().view(np.uint32)[...] = v
vb = pv.numpy().view(np.uint32)
del v
e.init_random_labels(K, 2)


def pinned(shape):
    t = torch.empty(int(np.prod(shape)), dtype=torch.float64, pin_memory=True)
    return t.numpy().reshape(shape, order="F")


st = tuple(pinned(x.shape) for x in e.get_stats(K))
for d, s in zip(st, e.get_stats(K)):
    d[...] = s
oW, oH = pinned((F, K)), pinned((N, K))
dv = torch.empty(pv.shape, dtype=torch.int32, device="cuda")
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); dv.copy_(pv, non_blocking=True); torch.cuda.synchronize()
    print(f"raw pinned H2D of the bit plane: {1e3 * (time.perf_counter() - t0):.1f} ms ({vb.nbytes / (time.perf_counter() - t0) / 1e9:.1f} GB/s)")
del dv
for rep in range(3):
    t0 = time.perf_counter(); e.set_data_bits(vb, None, F, N); e.synchronize()
    t1 = time.perf_counter(); e.set_stats(*st); e.synchronize()
    t2 = time.perf_counter(); e.vb_aspect(K, 1.0, 1.0, 1.0 / K, iter=1, checklb=True, out=(oW, oH)); e.synchronize()
    t3 = time.perf_counter()
    tm = e.timing()
    print(f"rep {rep}: set_data_bits {1e3 * (t1 - t0):.1f} ms ({vb.nbytes / (t1 - t0) / 1e9:.1f} GB/s)  set_stats {1e3 * (t2 - t1):.1f} ms  "
          f"vb_aspect {1e3 * (t3 - t2):.1f} ms (device total {tm['total_ms']:.1f}, rows {tm['pass_rows_ms']:.1f}, cols {tm['pass_cols_ms']:.1f}, params {tm['params_ms']:.1f})",
          flush=True)
