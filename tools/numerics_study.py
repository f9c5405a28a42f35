"""CPU study of where the tensor path's W/H error comes from and how it grows over VB iterations.

NumPy emulation of the arithmetic of nbmf_b200/csrc/tensor_pass.cu (f16 operands scaled by a power of two,
fp32 accumulation, R rounded to f16, exact-count row renormalisation) run as a full VB_aspect trajectory
beside the fp64 batch form, with individual error sources switched on and off:

  pert     fp64 arithmetic, one relative perturbation of size `eps` injected into the initial statistics
           (measures the amplification of the VB map itself: anything injected early grows by this factor)
  hi       stage 1 hi.hi, stage 2 R16 . W_hi                    (NBMF_PREC_TENSOR_FAST)
  hilo     stage 1 hi.hi, stage 2 R16 . (W_hi + W_lo)           (NBMF_PREC_TENSOR)
  hilo+U   hilo with the owner operand's lo part added to stage 1 (S = (U_hi+U_lo).W_hi)
  hilo+S   hilo with an exact stage 1 (only R16, W rounding left)
  R16only  everything exact except R rounded to f16
  S16only  everything exact except stage 1 (hi.hi)

    python tools/numerics_study.py [F N K iters]
"""
import sys
import numpy as np
import scipy.special as sp


def gen(F, N, Kt, seed):
    rng = np.random.default_rng(seed)
    W = rng.dirichlet(np.full(Kt, 1.0), size=F)
    H = rng.beta(1.0, 1.0, size=(N, Kt))
    V = rng.random((F, N)) < W @ H.T
    return V


def init_stats(V, K, seed):
    rng = np.random.default_rng(seed)
    F, N = V.shape
    Z = rng.integers(0, K, size=(F, N))
    nt = np.zeros((F, K)); fo = np.zeros((N, K)); fz = np.zeros((N, K))
    for k in range(K):
        m = Z == k
        nt[:, k] = m.sum(1); fo[:, k] = (m & V).sum(0); fz[:, k] = (m & ~V).sum(0)
    return nt, fo, fz


def params(nt, fo, fz, gamma):
    g = gamma + nt; a = 1.0 + fo; b = 1.0 + fz
    A = np.exp(sp.digamma(g) - sp.digamma(g.sum(1, keepdims=True)))
    B1 = np.exp(sp.digamma(a) - sp.digamma(a + b)); B0 = np.exp(sp.digamma(b) - sp.digamma(a + b))
    return A, B1, B0, g / g.sum(1, keepdims=True), a / (a + b)


def split(X):
    e = 14 - int(np.floor(np.log2(X.max())))
    Xs = np.ldexp(X, e)
    hi = Xs.astype(np.float16).astype(np.float64)
    lo = (Xs - hi).astype(np.float16).astype(np.float64)
    return hi, lo, e


def f16(x):
    return np.minimum(x, 65504.0).astype(np.float16).astype(np.float64)


def step(V, nt, fo, fz, gamma, mode):
    """one VB update of the statistics; mode: dict of switches"""
    F, N = V.shape
    A, B1, B0, EW, EH = params(nt, fo, fz, gamma)
    if mode.get("exact"):
        D = np.where(V, A @ B1.T, A @ B0.T)
        R = 1.0 / D
        R1 = np.where(V, R, 0.0); R0 = np.where(V, 0.0, R)
        return A * (R1 @ B1 + R0 @ B0), B1 * (R1.T @ A), B0 * (R0.T @ A), EW, EH
    Ah, Al, eA = split(A)
    BBs = np.concatenate([B1, B0], 0)
    Bh, Bl, eB = split(BBs)
    B1h, B0h, B1l, B0l = Bh[:N], Bh[N:], Bl[:N], Bl[N:]
    f32 = np.float32
    sc = 2.0 ** (eA + eB)

    def S_of(Uh, Ul, W1h, W0h, W1l, W0l, owner_is_A):
        # stage 1 as the pass with this owner computes it (fp32 accumulate)
        if mode.get("s_exact"):
            S1 = (A @ B1.T) * sc; S0 = (A @ B0.T) * sc
            return (S1, S0) if owner_is_A else (S1, S0)
        U = Uh + (Ul if mode.get("u_lo") else 0.0)
        S1 = (U.astype(f32) @ W1h.astype(f32).T).astype(np.float64)
        S0 = (U.astype(f32) @ W0h.astype(f32).T).astype(np.float64)
        return S1, S0

    def Rf(S1, S0):
        r = (2.0 ** (eA + eB - 14)) / np.where(V, S1, S0)
        if mode.get("r_exact"):
            return r
        if mode.get("r_diffuse"):   # sigma-delta along the reduction axis handled by the caller
            return r
        return f16(r.astype(f32).astype(np.float64))

    # rows pass: owner A (hi [+lo]), streamed B
    S1, S0 = S_of(Ah, Al, B1h, B0h, B1l, B0l, True)
    R = Rf(S1, S0)
    R1 = np.where(V, R, 0.0); R0 = np.where(V, 0.0, R)
    if mode.get("w_exact"):
        W1 = B1 * 2.0 ** eB; W0 = B0 * 2.0 ** eB
    else:
        W1 = B1h + (B1l if mode.get("w_lo") else 0.0); W0 = B0h + (B0l if mode.get("w_lo") else 0.0)
    GF = (R1.astype(f32) @ W1.astype(f32) + R0.astype(f32) @ W0.astype(f32)).astype(np.float64) * 2.0 ** (14 - eB)
    n_new = A * GF
    n_new *= (N / n_new.sum(1, keepdims=True))
    # cols pass: owner BB (hi [+lo]), streamed A
    if mode.get("s_exact"):
        S1c, S0c = S1, S0
    else:
        U1 = B1h + (B1l if mode.get("u_lo") else 0.0); U0 = B0h + (B0l if mode.get("u_lo") else 0.0)
        S1c = (Ah.astype(f32) @ U1.astype(f32).T).astype(np.float64)
        S0c = (Ah.astype(f32) @ U0.astype(f32).T).astype(np.float64)
    Rc = Rf(S1c, S0c)
    R1 = np.where(V, Rc, 0.0); R0 = np.where(V, 0.0, Rc)
    Wa = (A * 2.0 ** eA) if mode.get("w_exact") else (Ah + (Al if mode.get("w_lo") else 0.0))
    G1 = (R1.T.astype(f32) @ Wa.astype(f32)).astype(np.float64) * 2.0 ** (14 - eA)
    G0 = (R0.T.astype(f32) @ Wa.astype(f32)).astype(np.float64) * 2.0 ** (14 - eA)
    o_new = B1 * G1; z_new = B0 * G0
    c1 = V.sum(0)[:, None].astype(np.float64); c0 = (~V).sum(0)[:, None].astype(np.float64)
    o_new *= c1 / np.maximum(o_new.sum(1, keepdims=True), 1e-300)
    z_new *= c0 / np.maximum(z_new.sum(1, keepdims=True), 1e-300)
    return n_new, o_new, z_new, EW, EH


MODES = {
    "hi": dict(),
    "hilo": dict(w_lo=True),
    "hilo+U": dict(w_lo=True, u_lo=True),
    "hilo+S": dict(w_lo=True, s_exact=True),
    "R16only": dict(w_exact=True, s_exact=True),
    "S16only": dict(w_exact=True, r_exact=True),
    "S16+Ulo": dict(w_exact=True, r_exact=True, u_lo=True),
}


def main():
    F, N, K, iters = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (1024, 1024, 64, 100)))
    which = sys.argv[5].split(",") if len(sys.argv) > 5 else list(MODES) + ["pert"]
    gamma = 1.0 / K
    V = gen(F, N, 8, 5)
    st0 = init_stats(V, K, 6)
    ref = [tuple(x.copy() for x in st0)]
    EWs, EHs = [], []
    st = st0
    for i in range(iters):
        n, o, z, EW, EH = step(V, *st, gamma, dict(exact=True))
        EWs.append(EW); EHs.append(EH)
        st = (n, o, z)
    rel = lambda x, y: np.linalg.norm(x - y) / np.linalg.norm(y)
    marks = [m for m in (1, 2, 5, 10, 20, 40, 60, 80, 100, 150, 200) if m <= iters]
    print(f"F={F} N={N} K={K} gamma=1/K; relF error of (E_W, E_H) vs fp64 after i updates", flush=True)
    print("mode      " + "".join(f"{m:>20d}" for m in marks))
    for name in which:
        if name == "pert":
            rng = np.random.default_rng(9)
            eps = 1e-7
            st = tuple(x * (1 + eps * rng.standard_normal(x.shape)) for x in st0)
            mode = dict(exact=True)
        else:
            st = st0; mode = MODES[name]
        out = []
        for i in range(iters):
            n, o, z, EW, EH = step(V, *st, gamma, mode)
            st = (n, o, z)
            if i in marks:   # E_W at the top of iteration i = after i updates
                out.append((rel(EW, EWs[i]), rel(EH, EHs[i])))
        if iters in marks:
            _, _, _, EW, EH = step(V, *st, gamma, mode)
        print(f"{name:10s}" + "".join(f"   {a:.1e}/{b:.1e}" for a, b in out), flush=True)


if __name__ == "__main__":
    main()
