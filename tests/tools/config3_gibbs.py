"""BASELINE configs[2] at its size: uncollapsed Gibbs DirBer on data/mnistbinary (784 x 60000, tests/golden fixture),
K = 16, gamma = 1/16, 2000 sweeps, burn-in 0.5, seeds 1-5, 10 % of the pixels held out.

Part A (parity, same data on both sides): the engine on the first 2000 columns -- what the reference itself can hold
 -- against the oracle's restatement of Gibbs_DirBer_finite_Rcpp (collapsed_gibbs_z.cpp:363-437, pinned
draw for draw to the reference's own code) on those columns: predictive log-likelihood per held-out pixel and the
RMS difference of the posterior-mean E_V, beside the seed-to-seed spread of each side.
Part B (the config at its size): the engine on all 60000 columns, per-seed predictive LL and sweep time.
    python tests/tools/config3_gibbs.py [sweeps] [ref_iters]      (ref_iters = 0: part B only)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nbmf_b200 import Engine, _lib
from oracle import oracle
from tests import datasets

NA = datasets.NA
sweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
ref_iters = int(sys.argv[2]) if len(sys.argv) > 2 else 400
K, gamma, seeds = 16, 1.0 / 16, (1, 2, 3, 4, 5)


def ll_of(p, vt):
    p = np.clip(p, 1e-12, 1 - 1e-12)
    return float(np.mean(np.log(np.where(vt == 1, p, 1 - p))))


def engine_run(tr, idx, seed):
    e = Engine(precision=_lib.PREC_FP64)
    e.set_data(tr)
    Z = np.asfortranarray(np.random.default_rng(seed).integers(0, K, size=tr.shape).astype(np.int32))
    Z[tr == NA] = NA
    e.init_from_labels(Z, K)
    t0 = time.time()
    W, H, EV, _ = e.gibbs_dirber(K, 1.0, 1.0, gamma, iter=sweeps + 1, burnin=0.5, seed=seed, test_idx=idx)
    wall = time.time() - t0
    tm = e.timing()
    e.close()
    return EV, tm["total_ms"] / max(tm["iters"], 1), wall


V = datasets.mnistbinary_full()
F, N = V.shape
tr_full, te_full = datasets.hold_out(V, 0.1, 4)
rms = lambda a, b: float(np.sqrt(np.mean((a - b) ** 2)))

# ---- part A: first 2000 columns, engine vs the reference's collapsed chain
tr, te = np.asfortranarray(tr_full[:, :2000]), np.asfortranarray(te_full[:, :2000])
test = te != NA
fi, ni = np.nonzero(test)
idx = (fi + F * ni).astype(np.int64)
vt = V[:, :2000][test]
ev_e, ev_r, ll_e, ll_r = [], [], [], []
for s in seeds:
    EV, ms, _ = engine_run(tr, idx, s)
    ev_e.append(EV); ll_e.append(ll_of(EV, vt))
    if ref_iters <= 0:
        print(f"A seed {s}: engine {sweeps} sweeps ({ms:.3f} ms/sweep) LL {ll_e[-1]:.4f}", flush=True)
        continue
    Z = oracle.random_labels(tr, K, s)
    t0 = time.time()
    c = oracle.gibbs_dirber_finite(tr, Z, K, 1.0, 1.0, gamma, iter=ref_iters, burnin=0.5, seed=s)
    _, _, EVr = oracle.gibbs_expectation(tr, c["Z_samples"][:, :, :c["written"]], K, 1.0, 1.0, gamma)
    ev_r.append(EVr[test]); ll_r.append(ll_of(EVr[test], vt))
    print(f"A seed {s}: engine {sweeps} sweeps ({ms:.3f} ms/sweep) LL {ll_e[-1]:.4f} | reference chain {ref_iters} sweeps ({time.time() - t0:.0f} s CPU) LL {ll_r[-1]:.4f}", flush=True)
ev_e, ev_r = np.array(ev_e), np.array(ev_r)
if ref_iters <= 0:
  print(f"A 784x2000 K=16: predictive LL per held-out pixel  engine {np.mean(ll_e):.4f} +- {np.std(ll_e, ddof=1):.4f}; seed-to-seed E_V RMS {rms(ev_e[0], ev_e[1]):.4f}")
else:
  print(f"A 784x2000 K=16: predictive LL per held-out pixel  engine {np.mean(ll_e):.4f} +- {np.std(ll_e, ddof=1):.4f}   reference {np.mean(ll_r):.4f} +- {np.std(ll_r, ddof=1):.4f}")
  print(f"A E_V on the {idx.size} held-out pixels: RMS(engine mean - reference mean) {rms(ev_e.mean(0), ev_r.mean(0)):.4f}; "
      f"seed-to-seed RMS engine {rms(ev_e[0], ev_e[1]):.4f}, reference {rms(ev_r[0], ev_r[1]):.4f}", flush=True)

# ---- part B: all 60000 columns
test = te_full != NA
fi, ni = np.nonzero(test)
idx = (fi + F * ni).astype(np.int64)
vt = V[test]
ll_b, evs = [], []
for s in seeds:
    EV, ms, wall = engine_run(tr_full, idx, s)
    ll_b.append(ll_of(EV, vt)); evs.append(EV)
    print(f"B seed {s}: 784x60000 K=16 {sweeps} sweeps: {ms:.3f} ms/sweep incl. the held-out E_V accumulation over {idx.size} pixels "
          f"({F * N * K / (ms * 1e-3):.3e} entries*K/s), call {wall:.1f} s, predictive LL {ll_b[-1]:.4f}", flush=True)
print(f"B 784x60000 K=16: predictive LL per held-out pixel {np.mean(ll_b):.4f} +- {np.std(ll_b, ddof=1):.4f} over seeds {seeds}; seed-to-seed E_V RMS {rms(evs[0], evs[1]):.4f}")
