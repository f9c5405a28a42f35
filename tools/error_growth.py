"""Error of the tensor variants vs the FP64 path as a function of the number of VB updates
(checkpoints of the statistics inside one VB session; VB_Aspect_Rcpp's default is iter = 100,
collapsed_gibbs_z_VB.cpp:376).

    python tools/error_growth.py SIZE K CHECKPOINTS [variants] [chunks]
      variants  comma list of tensor,lo8,fast,fp32,pert   (default tensor,lo8,fast)
                pert = the FP64 path again from statistics perturbed by a relative 1e-10: how much the VB
                map itself amplifies a difference (what any arithmetic other than the reference's shows)
      chunks    comma list of NBMF_TP_CHUNK_STEPS values  (default: library default only)
      gamma     Dirichlet prior (default 1/K, the reference's experiments' choice)

The FP64 trajectory is computed once; every variant is compared with it at each checkpoint
(relative Frobenius error of E_W and E_H, max relative ELBO error so far).  Output lines are what
profiles/r2_error_growth_*.log hold and what tp_auto_mode's table (tensor_pass.cu) is fitted to."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbmf_b200 import Engine, _lib

F = N = int(sys.argv[1]); K = int(sys.argv[2]); cps = [int(x) for x in sys.argv[3].split(",")]
variants = (sys.argv[4] if len(sys.argv) > 4 else "tensor,lo8,fast").split(",")
chunks = [int(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else [0]
gamma = float(sys.argv[6]) if len(sys.argv) > 6 else 1.0 / K
PREC = {"fp64": _lib.PREC_FP64, "tensor": _lib.PREC_TENSOR, "fast": _lib.PREC_TENSOR_FAST, "lo8": _lib.PREC_TENSOR_LO8,
        "fp32": _lib.PREC_FP32}
rel = lambda x, y: np.linalg.norm(x - y) / np.linalg.norm(y)


def run(prec, ref=None, perturb=0.0):
    e = Engine(precision=prec)
    e.generate_synthetic(F, N, 8, 1.0, 1.0, 3); e.init_random_labels(K, 4)
    if perturb > 0:
        rng = np.random.default_rng(1)
        st = [x * (1.0 + perturb * rng.standard_normal(x.shape)) for x in e.get_stats(K)]
        e.set_stats(*st)
        del st
    e.vb_begin(K, 1.0, 1.0, gamma)
    out = {}
    t0 = time.time()
    for i in range(1, max(cps) + 1):
        e.vb_step(True)
        if i in cps:
            nt, fo, fz = e.get_stats(K)
            g = gamma + nt
            EW, EH = g / g.sum(1, keepdims=True), (1 + fo) / (2 + fo + fz)
            if ref is None:
                out[i] = (EW.astype(np.float64), EH.astype(np.float64))
            else:
                out[i] = (rel(EW, ref[i][0]), rel(EH, ref[i][1]))
    e.synchronize()
    wall = time.time() - t0
    _, _, lb = e.vb_end(max(cps), want_outputs=False)
    tm = e.timing()
    e.close()
    return out, lb, tm, wall


ref, lb0, tm0, wall0 = run(PREC["fp64"])
print(f"F=N={F} K={K} gamma={gamma:.4g} fp64 reference: {wall0 / max(cps) * 1e3:.1f} ms/update wall (path {tm0['path']})", flush=True)
for ch in chunks:
    if ch > 0:
        os.environ["NBMF_TP_CHUNK_STEPS"] = str(ch)
    else:
        os.environ.pop("NBMF_TP_CHUNK_STEPS", None)
    for name in variants:
        if name == "pert":
            if ch != chunks[0]:
                continue
            got, lb, tm, wall = run(PREC["fp64"], ref, perturb=1e-10)
        else:
            got, lb, tm, wall = run(PREC[name], ref)
        for i in cps:
            print(f"F=N={F} K={K} chunk={ch or 'default'} {name:6s} after {i:3d} updates: relF E_W {got[i][0]:.2e} E_H {got[i][1]:.2e} "
                  f"ELBO(max over first {i}) {np.max(np.abs(lb[:i] - lb0[:i]) / np.abs(lb0[:i])):.2e}", flush=True)
        print(f"F=N={F} K={K} chunk={ch or 'default'} {name:6s} last step: rows {tm['pass_rows_ms']:.3f} ms cols {tm['pass_cols_ms']:.3f} ms "
              f"params {tm['params_ms']:.3f} ms path {tm['path']}", flush=True)
