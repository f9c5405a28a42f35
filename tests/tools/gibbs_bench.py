"""Throughput of the uncollapsed Gibbs sweep (BASELINE config 3 shape: 784 x 60000, K=16) and a
larger synthetic, with the oracle's collapsed sampler timed beside it on a bounded sample."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nbmf_b200 import Engine  # noqa: E402
from oracle import oracle  # noqa: E402

out = []
for (F, N, K, iters) in [(784, 60000, 16, 201), (8192, 65536, 16, 21), (8192, 65536, 64, 11)]:
    e = Engine()
    e.generate_synthetic(F, N, 8, 0.3, 1.5, 3)
    e.init_random_labels(K, 4)
    e.gibbs_dirber(K, 1.0, 1.0, 1.0 / K, iter=4, burnin=0.5, seed=1)            # warm-up
    e.gibbs_dirber(K, 1.0, 1.0, 1.0 / K, iter=iters, burnin=0.5, seed=1)
    t = e.timing()
    e.close()
    sw = t["iters"]
    rec = dict(F=F, N=N, K=K, sweeps=sw, ms_per_sweep=t["total_ms"] / sw, sweeps_per_s=1e3 * sw / t["total_ms"],
               entries_K_per_s=F * N * K * sw / (t["total_ms"] * 1e-3), launches=t["launches"])
    out.append(rec)
    print(json.dumps(rec), flush=True)
# CPU reference (oracle collapsed sampler == the reference's own code draw for draw), bounded sample
if os.environ.get("NBMF_LIB_VARIANT"):
    sys.exit(0)   # variant runs: no CPU reference
F, N, K = 784, 200, 16
V, _, _ = oracle.generate(F, N, 8, 0.3, 1.5, 3)
Z = oracle.random_labels(V, K, 4)
t0 = time.perf_counter(); oracle.gibbs_dirber_finite(V, Z, K, 1.0, 1.0, 1.0 / K, iter=3, burnin=0.5, seed=1); t1 = time.perf_counter()
oracle.gibbs_dirber_finite(V, Z, K, 1.0, 1.0, 1.0 / K, iter=13, burnin=0.5, seed=1); t2 = time.perf_counter()
dt = (t2 - t1) - (t1 - t0)
print(json.dumps(dict(cpu_reference=dict(F=F, N=N, K=K, sweeps=10, ms_per_sweep=1e3 * dt / 10,
                                         entries_K_per_s=F * N * K * 10 / dt, cores=1))), flush=True)
