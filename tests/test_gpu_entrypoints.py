"""The remaining registered entry points of the path (src/RcppExports.cpp:316-339) through the C ABI:
posterior-mean reducers of the collapsed samplers, samples_VWH_DirBer, the eleven-argument
lower_bound_aspect and sample_gibbs_cpp -- against the reference's own functions compiled here
(oracle/_ref, when present) and the oracle restatement -- and the device Gamma / Beta / Dirichlet draws
against SciPy."""
import numpy as np
import pytest
import scipy.special as sp
import scipy.stats as st

from oracle import oracle, ref

pytestmark = pytest.mark.gpu
NA = oracle.NA


def _problem(F, N, K, na, seed):
    V, _, _ = oracle.generate(F, N, 3, 0.4, 0.4, seed)
    V = np.asfortranarray(V)
    if na > 0:
        V[np.random.default_rng(seed).random(V.shape) < na] = NA
    return V, oracle.random_labels(V, K, seed + 1)


def _sample_cube(V, K, J, seed):
    F, N = V.shape
    Zs = np.asfortranarray(np.random.default_rng(seed).integers(0, K, size=(F, N, J)).astype(np.int32))
    Zs[np.repeat((V == NA)[:, :, None], J, axis=2)] = NA
    return Zs


@pytest.mark.parametrize("F,N,K,J,na", [(40, 60, 5, 7, 0.2), (129, 33, 12, 3, 0.0), (8, 300, 3, 10, 0.5)])
def test_compute_expectation_reducers_bit_exact(F, N, K, J, na):
    from nbmf_b200 import Engine
    V, _ = _problem(F, N, K, na, 5)
    Zs = _sample_cube(V, K, J, 6)
    e = Engine()
    EW = e.compute_expectation_W(Zs, 0.7)
    EH = e.compute_expectation_H(Zs, V, 1.3, 0.9)
    e.close()
    oW = oracle.compute_expectation_W(Zs, 0.7)
    oH = oracle.compute_expectation_H(Zs, V, 1.3, 0.9)
    assert EW.shape == oW.shape and EH.shape == oH.shape
    assert np.allclose(EW, oW, rtol=1e-14, atol=0)      # counts are integers; only the order of gamma + c/J differs
    assert np.array_equal(EH, oH)                        # same sum order over samples: bit-identical
    assert np.all(EH[:, -1] == 0.0)                      # the reference's k < K quirk (:236-242)
    if ref.available():
        rW, rH = ref.compute_expectations(Zs, V, 1.3, 0.9, 0.7)
        assert np.allclose(EW, rW, rtol=1e-14) and np.allclose(EH, rH, rtol=1e-14)


def test_lower_bound_aspect_eleven_arguments_matches_reference():
    from nbmf_b200 import Engine
    F, N, K = 50, 70, 4
    V, Z = _problem(F, N, K, 0.15, 9)
    alpha, beta, gamma = 0.8, 1.1, 0.6
    r = oracle.vb_aspect(V, Z, K, alpha, beta, gamma, iter=6, checklb=True, return_E_Z=True)
    E_Z = r["E_Z"]
    obs = (V != NA)
    # parameters implied by E_Z (collapsed_gibbs_z_VB.cpp:419-441)
    nt = np.einsum("fkn,fn->fk", E_Z, obs.astype(float))
    fo = np.einsum("fkn,fn->nk", E_Z, (obs & (V == 1)).astype(float))
    fz = np.einsum("fkn,fn->nk", E_Z, (obs & (V == 0)).astype(float))
    g, a, b = gamma + nt, alpha + fo, beta + fz
    ElW = sp.digamma(g) - sp.digamma(g.sum(1, keepdims=True))
    ElH = sp.digamma(a) - sp.digamma(a + b); El1H = sp.digamma(b) - sp.digamma(a + b)
    e = Engine()
    lb, t7 = e.lower_bound_aspect_full(E_Z, ElW, ElH, El1H, g, a, b, alpha, beta, gamma, V)
    e.close()
    olb, ot7 = oracle.lower_bound_aspect(E_Z, ElW, ElH, El1H, g, a, b, alpha, beta, gamma, V)
    assert np.allclose(t7, ot7, rtol=1e-10, atol=1e-9), (t7, ot7)
    assert abs(lb - olb) <= 1e-10 * abs(olb)
    if ref.available():
        rlb = ref.lower_bound_aspect(E_Z, ElW, ElH, El1H, g, a, b, alpha, beta, gamma, V)
        assert abs(lb - rlb) <= 1e-10 * abs(rlb)


def test_samples_VWH_DirBer_moments():
    """W_f ~ Dir(gamma + n_total_f), H_nk ~ Beta(alpha + ones, beta + zeros), V ~ Bernoulli(W H^T)
    (collapsed_gibbs_z.cpp:589-616): the same label matrix repeated J times gives J independent draws of
    the same conditionals, whose sample means must match the analytic means within Monte-Carlo error."""
    from nbmf_b200 import Engine
    F, N, K, J = 24, 30, 4, 600
    V, Z = _problem(F, N, K, 0.1, 21)
    gamma, alpha, beta = 0.5, 1.5, 0.7
    Zs = np.asfortranarray(np.repeat(Z[:, :, None], J, axis=2))
    e = Engine()
    out = e.samples_VWH_DirBer(V, Zs, K, gamma, alpha, beta, seed=3)
    e.close()
    Ws, Hs, Vs = out["W_samples"], out["H_samples"], out["V_samples"]
    obs = V != NA
    nt = np.stack([((Z == k) & obs).sum(1) for k in range(K)], 1).astype(float)
    fo = np.stack([((Z == k) & obs & (V == 1)).sum(0) for k in range(K)], 1).astype(float)
    fz = np.stack([((Z == k) & obs & (V == 0)).sum(0) for k in range(K)], 1).astype(float)
    g = gamma + nt
    mW = g / g.sum(1, keepdims=True)
    a, b = alpha + fo, beta + fz
    mH = a / (a + b)
    assert np.allclose(Ws.sum(1), 1.0, atol=1e-12)
    g0 = g.sum(1, keepdims=True)
    sdW = np.sqrt(mW * (1 - mW) / (g0 + 1)) / np.sqrt(J)
    sdH = np.sqrt(a * b / ((a + b) ** 2 * (a + b + 1))) / np.sqrt(J)
    zW = (Ws.mean(2) - mW) / np.maximum(sdW, 1e-12)
    zH = (Hs.mean(2) - mH) / np.maximum(sdH, 1e-12)
    assert np.abs(zW).max() < 5.0 and abs(zW.mean()) < 0.5, (np.abs(zW).max(), zW.mean())
    assert np.abs(zH).max() < 5.0 and abs(zH.mean()) < 0.5, (np.abs(zH).max(), zH.mean())
    # V | W, H: mean over draws of V_samples vs mean over draws of W H^T
    P = np.einsum("fkj,nkj->fn", Ws, Hs) / J
    zV = (Vs.mean(2) - P) / np.sqrt(np.maximum(P * (1 - P), 1e-6) / J)
    assert np.abs(zV).max() < 5.5 and abs(zV.mean()) < 0.3
    assert set(np.unique(Vs)) <= {0, 1}


def test_sample_gibbs_cpp_as_coded_equals_the_reference():
    """sample_gibbs_cpp never stores its draw (sample_gibbs.cpp:60-69): whatever the random numbers, the
    result is C_init / (number of kept states) -- deterministic, so the C ABI must reproduce the
    reference's own function exactly."""
    from nbmf_b200 import Engine
    F, K, it, burn = 37, 5, 40, 0.5
    rng = np.random.default_rng(4)
    W = rng.dirichlet(np.ones(K), size=F)
    v = rng.integers(0, 2, size=F).astype(np.int32)
    lab = rng.integers(0, K, size=F)
    C0 = np.zeros((F, K), dtype=np.int32, order="F"); C0[np.arange(F), lab] = 1
    e = Engine()
    got = e.sample_gibbs_cpp(v, W, C0, 1.0, it, burn, seed=9, as_coded=True)
    e.close()
    kept = 1 + sum(1 for i in range(1, it) if i >= int(np.floor(it * burn)))
    # every row is all-zero after the first sweep: only the initial state contributes (:47, :60, :73-78)
    assert np.allclose(got, C0 / kept, rtol=1e-14, atol=0)
    if ref.available():
        r = ref.sample_gibbs_cpp(v, W, C0, 1.0, it, burn, seed=9)
        assert np.allclose(got, r, rtol=1e-14, atol=1e-15), np.abs(got - r).max()


def test_sample_gibbs_column_chain_matches_restatement_within_mc_error():
    """as_coded = 0: the chain R's sample_gibbs_Z runs (R/MCEM - sample_gibbs.R:1-37).  The device chain
    (Philox) and the NumPy restatement (its own RNG) must give the same posterior mean assignments within
    Monte-Carlo error over seeds."""
    from nbmf_b200 import Engine
    F, K, it = 60, 4, 400
    rng = np.random.default_rng(12)
    W = rng.beta(0.5, 0.5, size=(F, K)).clip(0.02, 0.98)
    v = rng.integers(0, 2, size=F).astype(np.int32)
    lab = rng.integers(0, K, size=F)
    C0 = np.zeros((F, K), dtype=np.int32, order="F"); C0[np.arange(F), lab] = 1
    seeds = range(1, 9)
    e = Engine()
    dev = np.stack([e.sample_gibbs_cpp(v, W, C0, 1.0, it, 0.5, seed=s, as_coded=False) for s in seeds])
    e.close()
    cpu = np.stack([oracle.sample_gibbs_ZH_column(v, W, lab, 1.0, it, 0.5, seed=100 + s) for s in seeds])
    assert np.allclose(dev.sum(2), 1.0, atol=1e-12)
    md, mc = dev.mean(0), cpu.mean(0)
    se = np.sqrt(dev.var(0, ddof=1) / len(dev) + cpu.var(0, ddof=1) / len(cpu)) + 5e-3
    z = (md - mc) / se
    assert np.abs(z).max() < 5.0, np.abs(z).max()
    assert np.abs(md - mc).mean() < 0.02


# ---------------------------------------------------------------------------
# device random variates (a9): Gamma (shape << 1 up to large counts), Beta
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("which", [0, 1])
@pytest.mark.parametrize("shape", [1.0 / 128, 0.1, 0.5, 1.0, 3.7, 250.0, 262144.0])
def test_device_gamma_draws_ks_and_moments(which, shape):
    """Marsaglia-Tsang with the U^(1/a) boost for shape < 1 (gamma = 1/K priors with empty counts) up to counts of a
    whole row of C5: Kolmogorov-Smirnov against SciPy's Gamma CDF and the first two moments.  For shape 1/128 half
    a percent of the draws are below the smallest double; they sit at 0, where the CDF is 0.004."""
    from nbmf_b200 import Engine
    n = 200_000
    e = Engine()
    x = e.rng_probe(which, shape, 0.0, seed=17 + which, n=n)
    e.close()
    assert np.all(np.isfinite(x)) and np.all(x >= 0)
    rel32 = 3e-6 * shape if which == 1 else 0.0           # fp32 arithmetic of the fast sampler
    assert abs(x.mean() - shape) < 6.0 * np.sqrt(shape / n) + rel32 + 1e-12, (x.mean(), shape)
    # Var[sample variance] / sigma^4 = (kurtosis - 1) / n, kurtosis of Gamma(a) = 3 + 6 / a
    assert abs(x.var() / shape - 1.0) < 5.0 * np.sqrt((2.0 + 6.0 / shape) / n) + 1e-3
    d, _ = st.kstest(x, "gamma", args=(shape,))
    assert d < (0.006 if shape >= 0.1 else 0.01), d


def _ks_logit_beta(t, a, b):
    """KS distance between logit samples t and the logit of Beta(a, b); the upper tail through the survival
    function of the mirrored variable, so that nothing is lost next to 1."""
    t = np.sort(t)
    n = t.size
    sig = lambda z: 1.0 / (1.0 + np.exp(-z))
    cdf = np.where(t <= 0, st.beta.cdf(sig(np.minimum(t, 0)), a, b), 1.0 - st.beta.cdf(sig(-np.maximum(t, 0)), b, a))
    return max(np.max(np.arange(1, n + 1) / n - cdf), np.max(cdf - np.arange(0, n) / n))


@pytest.mark.parametrize("a,b", [(1.0, 1.0), (0.3, 2.5), (40.0, 3.0), (1.0 / 16, 1.0 / 16), (1.0 / 128, 2.0), (3000.0, 5000.0)])
def test_device_beta_draws_ks(a, b):
    from nbmf_b200 import Engine
    n = 200_000
    e = Engine()
    x = e.rng_probe(2, a, b, seed=23, n=n)
    t = e.rng_probe(3, a, b, seed=23, n=n)
    e.close()
    assert np.all((x >= 0) & (x <= 1))
    m = a / (a + b)
    v = a * b / ((a + b) ** 2 * (a + b + 1))
    assert abs(x.mean() - m) < 6 * np.sqrt(v / n) + 2e-7
    d = _ks_logit_beta(t, a, b)
    assert d < 0.006, d


def test_device_dirichlet_row_from_the_gibbs_draw_kernel():
    """gibbs_draw_kernel's Dirichlet rows through the public chain: with the burn-in keeping every sweep and a
    tiny data set the Rao-Blackwell means are checked elsewhere; here the FFMA-peak probe and the row sums of
    samples_VWH_DirBer's W (fp64 Gamma stream) cover the normalisation."""
    from nbmf_b200 import Engine
    e = Engine()
    t = e.ffma_peak_tflops()
    e.close()
    assert 20.0 < t < 200.0, t
