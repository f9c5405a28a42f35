"""Generates tests/golden/ref_vectors.npz from the reference's OWN sources
(oracle/_ref/libnbmf_ref.so = /root/reference/src/*.cpp compiled against oracle/ref_shim).
Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
The vectors pin the oracle (tests/test_oracle_pin.py) on machines without the reference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle, ref  # noqa: E402

assert ref.available(), "build oracle/_ref first: make -C oracle ref"
out = {}
NA = oracle.NA


def problem(F, N, Kt, seed, na):
    V, _, _ = oracle.generate(F, N, Kt, 0.3, 0.3, seed)
    V = V.copy(order="F")
    if na > 0:
        V[np.random.default_rng(seed + 1).random(V.shape) < na] = NA
    return V


cases = {"a": (23, 31, 4, 0.2, 1.0, 0.8, 0.7, 25), "b": (40, 17, 3, 0.0, 1.0, 1.0, 1.0, 40),
         "c": (16, 50, 7, 0.35, 0.5, 2.0, 1.0 / 7, 15)}
for name, (F, N, K, na, al, be, ga, it) in cases.items():
    V = problem(F, N, 3, 11 + len(name), na)
    Z = oracle.random_labels(V, K, 5)
    r = ref.vb_aspect(V, Z, K, al, be, ga, iter=it, checklb=True, return_E_Z=True)
    out[f"aspect_{name}_V"] = V; out[f"aspect_{name}_Z"] = Z
    out[f"aspect_{name}_par"] = np.array([K, al, be, ga, it])
    for k in ("E_W", "E_H", "lowerbounds", "E_Z"):
        out[f"aspect_{name}_{k}"] = r[k]
    Z1 = oracle.random_labels(V, K + 1, 6)
    d = ref.vb_dirber(V, Z1, K, al, be, ga, iter=max(it // 2, 2), return_E_Z=True)
    out[f"dirber_{name}_Z"] = Z1
    for k in ("E_W", "E_H", "E_Z"):
        out[f"dirber_{name}_{k}"] = d[k]
    g = ref.gibbs_dirber_finite(V, Z, K, al, be, ga, iter=21, burnin=0.5, seed=7)
    out[f"gibbs_{name}_Z_samples"] = g["Z_samples"]
    ew, eh = ref.compute_expectations(g["Z_samples"], V, al, be, ga)
    out[f"gibbs_{name}_E_W"] = ew; out[f"gibbs_{name}_E_H"] = eh
    # sibling samplers (collapsed_gibbs_z.cpp:441-557, :275-359; the DP one is undefined with NA)
    dd = ref.gibbs_dirdir(V, Z, K, al, ga, iter=13, burnin=0.5, seed=9)
    out[f"dirdir_{name}_Z_samples"] = dd["Z_samples"]; out[f"dirdir_{name}_C_samples"] = dd["C_samples"]
    if na == 0.0:
        dp = ref.gibbs_dirber_dp(V, Z, al, be, ga, iter=13, burnin=0.5, seed=9)
        out[f"dp_{name}_Z_samples"] = dp["Z_samples"]
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_vectors.npz"), **out)
print("wrote", len(out), "arrays")
