"""Where does the per-update error of the tensor path come from?  One VB update from the same statistics on
the FP64 path and on a tensor variant; the element-wise relative error of the new statistics (= of the G
accumulators, up to the exact per-lane renormalisation) is split into
  * its correlation with phi(m) = (m - 2/3) / m^2, m = mantissa of the accumulator: the signature of fp32
    accumulation that truncates (a chain of n additions of a linearly growing positive sum loses
    n * ulp/2 * phi(m) relative: DESIGN.md 5.1), which is systematic -- the same sign every update;
  * per-component (column k) means: a bias that follows the streamed operand;
  * the remaining element-wise noise.
    python tools/err_anatomy.py SIZE K variant [chunk ...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbmf_b200 import Engine, _lib

F = N = int(sys.argv[1]); K = int(sys.argv[2]); variant = sys.argv[3]
chunks = [int(x) for x in sys.argv[4:]] or [0]
PREC = {"tensor": _lib.PREC_TENSOR, "fast": _lib.PREC_TENSOR_FAST, "lo8": _lib.PREC_TENSOR_LO8}
gamma = 1.0 / K


def phi(x):
    m = x / 2.0 ** np.floor(np.log2(x))
    return (m - 2.0 / 3.0) / m ** 2


# reference: 3 FP64 warm-up updates (so that the operands are not the flat initial ones), then one more
e = Engine(precision=_lib.PREC_FP64)
e.generate_synthetic(F, N, 8, 1.0, 1.0, 3); e.init_random_labels(K, 4)
e.vb_begin(K, 1.0, 1.0, gamma)
for _ in range(3):
    e.vb_step(False)
e.vb_end(0, want_outputs=False)
st0 = e.get_stats(K)
e.vb_begin(K, 1.0, 1.0, gamma); e.vb_step(False); e.vb_end(0, want_outputs=False)
ref = e.get_stats(K)
e.close()
for ch in chunks:
    if ch > 0:
        os.environ["NBMF_TP_CHUNK_STEPS"] = str(ch)
    t = Engine(precision=PREC[variant])
    t.generate_synthetic(F, N, 8, 1.0, 1.0, 3)
    t.set_stats(*st0)
    t.vb_begin(K, 1.0, 1.0, gamma); t.vb_step(False); t.vb_end(0, want_outputs=False)
    got = t.get_stats(K)
    t.close()
    for name, g, r in (("n_total (rows pass)", got[0], ref[0]), ("f_ones (cols pass)", got[1], ref[1]), ("f_zeros (cols pass)", got[2], ref[2])):
        big = r > 1e-3 * r.max()                      # elements that matter for the Frobenius norm
        err = np.where(big, (g - r) / np.where(big, r, 1.0), 0.0)
        cnt = np.maximum(big.sum(1, keepdims=True), 1)
        err_c = np.where(big, err - err.sum(1, keepdims=True) / cnt, 0.0)      # row-constant part removed (it is renormalised away)
        p = np.where(big, phi(np.where(big, r, 1.0)), 0.0)
        p_c = np.where(big, p - p.sum(1, keepdims=True) / cnt, 0.0)
        slope = (err_c * p_c).sum() / (p_c ** 2).sum()
        resid = err_c - slope * p_c
        colmean = np.array([resid[:, k][big[:, k]].mean() if big[:, k].any() else 0.0 for k in range(K)])
        resid2 = resid - np.where(big, colmean[None, :], 0.0)
        rms = lambda a: np.sqrt((a[big] ** 2).mean())
        print(f"F=N={F} K={K} {variant} chunk={ch or 'default'} {name:20s}: relF {np.linalg.norm(g - r) / np.linalg.norm(r):.2e}  "
              f"elementwise rms {rms(err_c):.2e} = phi-part {rms(slope * p_c):.2e} (slope {slope:.2e}) + per-component part {rms(np.where(big, colmean[None, :], 0.0)):.2e} "
              f"+ rest {rms(resid2):.2e}; corr(err, phi) {np.corrcoef(err_c[big], p_c[big])[0, 1]:+.3f}", flush=True)
