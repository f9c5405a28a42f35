/*
 * nbmf_oracle.c -- CPU restatement (fp64, single thread) of the hot path of
 * alumbreras/NBMF (R package rMMLEDirBer).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke check
 * in __graft_entry__.py and the cpu_baseline / --impl reference legs of
 * bench.py may load it.  libnbmf_cuda never links or calls anything here.
 *
 * Parity status: the reference ships no golden vectors, known-answer tests or
 * fixtures for this path, and its own sources need
 * Rcpp/RcppArmadillo/Boost which are absent from the image.  The oracle is
 * pinned instead against the reference's *own source files* compiled where
 * they lie against a minimal Armadillo/Rcpp shim (oracle/ref_shim, built into
 * oracle/_ref by oracle/Makefile); tests/test_oracle_pin.py compares the two
 * and tests/golden/ holds vectors produced by that build.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).  Conventions are the reference's: column-major matrices,
 * V and Z are int32 with NA = INT_MIN (R's NA_INTEGER), Z labels 0-based,
 * E_Z cube element (f,k,n) at f + F*k + F*Kc*n (arma::cube layout).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

#define NA_INT INT_MIN

/* ------------------------------------------------------------------------ */
/* special functions                                                         */
/* ------------------------------------------------------------------------ */

/* digamma: stands in for boost::math::digamma (collapsed_gibbs_z_VB.cpp:5,430).
 * Recurrence psi(x) = psi(x+1) - 1/x up to x >= 10, then the asymptotic series
 * with Bernoulli numbers through B12; |error| < 1e-15 for x > 0.            */
double oracle_digamma(double x)
{
    double r = 0.0;
    if (!(x > 0.0)) {
        if (x == 0.0) return -INFINITY;
        /* reflection: psi(1-x) - psi(x) = pi*cot(pi*x) */
        if (x == floor(x)) return NAN;
        return oracle_digamma(1.0 - x) - M_PI / tan(M_PI * x);
    }
    while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
    {
        double inv = 1.0 / x, inv2 = inv * inv;
        double s = inv2 * (1.0 / 12.0 - inv2 * (1.0 / 120.0 - inv2 * (1.0 / 252.0 -
                   inv2 * (1.0 / 240.0 - inv2 * (1.0 / 132.0 - inv2 * (691.0 / 32760.0 -
                   inv2 * (1.0 / 12.0)))))));
        r += log(x) - 0.5 * inv - s;
    }
    return r;
}

/* lgamma: stands in for boost::math::lgamma (collapsed_gibbs_z_VB.cpp:6,177). */
double oracle_lgamma(double x) { return lgamma(x); }

/* ------------------------------------------------------------------------ */
/* RNG for the sampling oracles.  The reference draws from R's global RNG    */
/* (rmultinom / Rcpp::rbeta / arma::randg), which does not exist here; Gibbs */
/* parity is statistical by construction.      */
/* ------------------------------------------------------------------------ */
typedef struct { uint64_t s[4]; } orng;

static uint64_t splitmix64(uint64_t *x)
{
    uint64_t z = (*x += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static void orng_seed(orng *g, uint64_t seed)
{
    for (int i = 0; i < 4; i++) g->s[i] = splitmix64(&seed);
}
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t orng_next(orng *g) /* xoshiro256** */
{
    uint64_t *s = g->s, r = rotl64(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl64(s[3], 45);
    return r;
}
static double orng_unif(orng *g) /* (0,1) */
{
    return ((double)(orng_next(g) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
static double orng_norm(orng *g)
{
    double u = orng_unif(g), v = orng_unif(g);
    return sqrt(-2.0 * log(u)) * cos(2.0 * M_PI * v);
}
/* Gamma(shape a, scale 1): Marsaglia-Tsang, with the U^(1/a) boost for a<1.
 * Stands in for arma::randg (collapsed_gibbs_z.cpp:593).                     */
static double orng_gamma(orng *g, double a)
{
    if (a < 1.0) {
        double u = orng_unif(g);
        return orng_gamma(g, a + 1.0) * pow(u, 1.0 / a);
    }
    double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (;;) {
        double x = orng_norm(g), v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        double u = orng_unif(g);
        if (u < 1.0 - 0.0331 * x * x * x * x) return d * v;
        if (log(u) < 0.5 * x * x + d * (1.0 - v + log(v))) return d * v;
    }
}
/* one categorical draw from normalised probs; stands in for
 * rmultinom(1, probs, K, ans) + index_max (collapsed_gibbs_z.cpp:411-413).   */
static int orng_categorical(orng *g, const double *p, int K)
{
    double u = orng_unif(g), c = 0.0;
    for (int k = 0; k < K; k++) { c += p[k]; if (u < c) return k; }
    for (int k = K - 1; k >= 0; k--) if (p[k] > 0.0) return k;
    return K - 1;
}

/* ------------------------------------------------------------------------ */
/* vbcounts  (collapsed_gibbs_z_VB.cpp:14-35)                                */
/* ------------------------------------------------------------------------ */
typedef struct {
    int F, N, K;
    double *f_ones;   /* N x K */
    double *f_zeros;  /* N x K */
    double *f_total;  /* N x K */
    double *n_total;  /* F x K */
    double *fn_total; /* K     */
} vbcounts;

static int vbcounts_alloc(vbcounts *c, int F, int K, int N)
{
    c->F = F; c->N = N; c->K = K;
    c->f_ones = calloc((size_t)N * K, sizeof(double));
    c->f_zeros = calloc((size_t)N * K, sizeof(double));
    c->f_total = calloc((size_t)N * K, sizeof(double));
    c->n_total = calloc((size_t)F * K, sizeof(double));
    c->fn_total = calloc((size_t)K, sizeof(double));
    return (c->f_ones && c->f_zeros && c->f_total && c->n_total && c->fn_total) ? 0 : -1;
}
static void vbcounts_free(vbcounts *c)
{
    free(c->f_ones); free(c->f_zeros); free(c->f_total); free(c->n_total); free(c->fn_total);
}

#define EZ(f, k, n) E_Z[(size_t)(f) + (size_t)F * (k) + (size_t)F * Kc * (n)]

/* vb_matrix_to_tensor (commons.cpp:29-42): one-hot E_Z from labels, NA -> 0 */
static void vb_matrix_to_tensor(const int32_t *Z, int F, int N, int Kc, double *E_Z)
{
    memset(E_Z, 0, sizeof(double) * (size_t)F * Kc * N);
    for (int n = 0; n < N; n++)
        for (int f = 0; f < F; f++) {
            int k = Z[(size_t)f + (size_t)F * n];
            if (k != NA_INT) EZ(f, k, n) = 1.0;
        }
}

/* compute_vbcounts (collapsed_gibbs_z_VB.cpp:42-66) */
static void compute_vbcounts(const int32_t *V, const double *E_Z, int Kc, vbcounts *c)
{
    int F = c->F, N = c->N;
    memset(c->f_ones, 0, sizeof(double) * (size_t)N * Kc);
    memset(c->f_zeros, 0, sizeof(double) * (size_t)N * Kc);
    memset(c->f_total, 0, sizeof(double) * (size_t)N * Kc);
    memset(c->n_total, 0, sizeof(double) * (size_t)F * Kc);
    memset(c->fn_total, 0, sizeof(double) * (size_t)Kc);
    for (int n = 0; n < N; n++)
        for (int f = 0; f < F; f++) {
            int v = V[(size_t)f + (size_t)F * n];
            if (v != NA_INT)
                for (int k = 0; k < Kc; k++) {
                    double p = EZ(f, k, n);
                    c->f_ones[n + (size_t)N * k] += p * v;
                    c->f_zeros[n + (size_t)N * k] += p * (1 - v);
                    c->f_total[n + (size_t)N * k] += p;
                    c->n_total[f + (size_t)F * k] += p;
                    c->fn_total[k] += p;
                }
        }
}

/* set_vbassignment (collapsed_gibbs_z_VB.cpp:69-78) */
static inline void set_vbassignment(int f, int n, const double *probs, double *E_Z, int v,
                                    int Kc, vbcounts *c)
{
    int F = c->F, N = c->N;
    for (int k = 0; k < Kc; k++) {
        double p = probs[k];
        EZ(f, k, n) = p;
        c->f_ones[n + (size_t)N * k] += p * v;
        c->f_zeros[n + (size_t)N * k] += p * (1 - v);
        c->f_total[n + (size_t)N * k] += p;
        c->n_total[f + (size_t)F * k] += p;
        c->fn_total[k] += p;
    }
}

/* remove_vbassignment (collapsed_gibbs_z_VB.cpp:81-92) */
static inline void remove_vbassignment(int f, int n, double *E_Z, int v, int Kc, vbcounts *c)
{
    int F = c->F, N = c->N;
    for (int k = 0; k < Kc; k++) {
        double p = EZ(f, k, n);
        EZ(f, k, n) = 0.0;
        c->f_ones[n + (size_t)N * k] -= p * v;
        c->f_zeros[n + (size_t)N * k] -= p * (1 - v);
        c->f_total[n + (size_t)N * k] -= p;
        c->n_total[f + (size_t)F * k] -= p;
        c->fn_total[k] -= p;
    }
}

/* ------------------------------------------------------------------------ */
/* ELBO terms  (collapsed_gibbs_z_VB.cpp:173-294; second statement in        */
/* R/algo_VB_aspect.R:187-266)                                               */
/* ------------------------------------------------------------------------ */

/* E_log_p_W_Rcpp (:173-182) */
double oracle_E_log_p_W(const double *E_log_W, int F, int K, double gamma)
{
    double x = F * (oracle_lgamma(K * gamma) - K * oracle_lgamma(gamma));
    for (int f = 0; f < F; f++) {
        double s = 0.0;
        for (int k = 0; k < K; k++) s += E_log_W[f + (size_t)F * k];
        x += (gamma - 1) * s;
    }
    return x;
}

/* E_log_q_W_Rcpp (:185-197) */
double oracle_E_log_q_W(const double *E_log_W, const double *gamma_vb, int F, int K)
{
    double x = 0.0;
    for (int f = 0; f < F; f++) {
        double s = 0.0, t = 0.0;
        for (int k = 0; k < K; k++) s += gamma_vb[f + (size_t)F * k];
        x += oracle_lgamma(s);
        for (int k = 0; k < K; k++) x -= oracle_lgamma(gamma_vb[f + (size_t)F * k]);
        for (int k = 0; k < K; k++)
            t += (gamma_vb[f + (size_t)F * k] - 1) * E_log_W[f + (size_t)F * k];
        x += t;
    }
    return x;
}

/* E_log_p_Z_Rcpp (:201-222) */
double oracle_E_log_p_Z(const double *E_Z, const double *E_log_W, const int32_t *V, int F, int Kc,
                        int N)
{
    double x = 0.0;
    for (int n = 0; n < N; n++)
        for (int f = 0; f < F; f++)
            if (V[(size_t)f + (size_t)F * n] != NA_INT)
                for (int k = 0; k < Kc; k++) x += EZ(f, k, n) * E_log_W[f + (size_t)F * k];
    return x;
}

/* E_log_q_Z_Rcpp (:225-240); 0*log(0) is NaN exactly as in the reference (:234) */
double oracle_E_log_q_Z(const double *E_Z, const int32_t *V, int F, int Kc, int N)
{
    double x = 0.0;
    for (int n = 0; n < N; n++)
        for (int f = 0; f < F; f++)
            if (V[(size_t)f + (size_t)F * n] != NA_INT)
                for (int k = 0; k < Kc; k++) x += EZ(f, k, n) * log(EZ(f, k, n));
    return x;
}

/* E_log_p_H_Rcpp (:243-253) */
double oracle_E_log_p_H(const double *E_log_H, const double *E_log_1_H, int N, int K, double alpha,
                        double beta)
{
    double s1 = 0.0, s0 = 0.0;
    for (size_t i = 0; i < (size_t)N * K; i++) { s1 += E_log_H[i]; s0 += E_log_1_H[i]; }
    return (double)N * K * (oracle_lgamma(alpha + beta) - (oracle_lgamma(alpha) + oracle_lgamma(beta))) +
           (alpha - 1) * s1 + (beta - 1) * s0;
}

/* E_log_q_H_Rcpp (:256-273) */
double oracle_E_log_q_H(const double *E_log_H, const double *E_log_1_H, const double *alpha_vb,
                        const double *beta_vb, int N, int K)
{
    double x = 0.0;
    for (int n = 0; n < N; n++)
        for (int k = 0; k < K; k++) {
            size_t i = n + (size_t)N * k;
            x += oracle_lgamma(alpha_vb[i] + beta_vb[i]) -
                 (oracle_lgamma(alpha_vb[i]) + oracle_lgamma(beta_vb[i])) +
                 (alpha_vb[i] - 1) * E_log_H[i] + (beta_vb[i] - 1) * E_log_1_H[i];
        }
    return x;
}

/* E_log_p_V_Rcpp (:276-294) */
double oracle_E_log_p_V(const int32_t *V, const double *E_Z, const double *E_log_H,
                        const double *E_log_1_H, int F, int Kc, int N)
{
    double x = 0.0;
    for (int n = 0; n < N; n++)
        for (int f = 0; f < F; f++) {
            int v = V[(size_t)f + (size_t)F * n];
            if (v == NA_INT) continue;
            for (int k = 0; k < Kc; k++)
                x += EZ(f, k, n) *
                     (v * E_log_H[n + (size_t)N * k] + (1 - v) * E_log_1_H[n + (size_t)N * k]);
        }
    return x;
}

/* lower_bound_aspect (:339-369).  terms7 (optional) receives
 * pW,pZ,pH,pV,qW,qZ,qH in that order.                                       */
double oracle_lower_bound_aspect(const double *E_Z, const double *E_log_W, const double *E_log_H,
                                 const double *E_log_1_H, const double *gamma_vb,
                                 const double *alpha_vb, const double *beta_vb, double alpha,
                                 double beta, double gamma, const int32_t *V, int F, int Kc, int N,
                                 double *terms7)
{
    double pW = oracle_E_log_p_W(E_log_W, F, Kc, gamma);
    double qW = oracle_E_log_q_W(E_log_W, gamma_vb, F, Kc);
    double pZ = oracle_E_log_p_Z(E_Z, E_log_W, V, F, Kc, N);
    double qZ = oracle_E_log_q_Z(E_Z, V, F, Kc, N);
    double pH = oracle_E_log_p_H(E_log_H, E_log_1_H, N, Kc, alpha, beta);
    double qH = oracle_E_log_q_H(E_log_H, E_log_1_H, alpha_vb, beta_vb, N, Kc);
    double pV = oracle_E_log_p_V(V, E_Z, E_log_H, E_log_1_H, F, Kc, N);
    if (terms7) {
        terms7[0] = pW; terms7[1] = pZ; terms7[2] = pH; terms7[3] = pV;
        terms7[4] = qW; terms7[5] = qZ; terms7[6] = qH;
    }
    return pW + pZ + pH + pV - qW - qZ - qH;
}

/* ------------------------------------------------------------------------ */
/* VB_Aspect_Rcpp  (collapsed_gibbs_z_VB.cpp:373-526) -- literal sequential  */
/* restatement, E_Z cube and in-place counters included.                     */
/* Outputs: E_W F x K, E_H N x K, lowerbounds[iter] (0 when !checklb, as the */
/* reference :474-477), E_Z_out optional F x K x N.  Returns 0, -1 on OOM,   */
/* 1 if a negative value was found by the reference's assertions.            */
/* ------------------------------------------------------------------------ */
int oracle_vb_aspect(const int32_t *V, const int32_t *Z, int F, int N, int K, double alpha,
                     double beta, double gamma, int iter, int checklb, double *E_W, double *E_H,
                     double *lowerbounds, double *E_Z_out)
{
    const int Kc = K; /* :379 */
    vbcounts cnt;
    int rc = 0;
    double *E_Z = malloc(sizeof(double) * (size_t)F * Kc * N);
    double *gamma_vb = calloc((size_t)F * Kc, sizeof(double));
    double *alpha_vb = calloc((size_t)N * Kc, sizeof(double));
    double *beta_vb = calloc((size_t)N * Kc, sizeof(double));
    double *E_log_W = calloc((size_t)F * Kc, sizeof(double));
    double *E_log_H = calloc((size_t)N * Kc, sizeof(double));
    double *E_log_1_H = calloc((size_t)N * Kc, sizeof(double));
    double *probs = malloc(sizeof(double) * Kc);
    if (!E_Z || !gamma_vb || !alpha_vb || !beta_vb || !E_log_W || !E_log_H || !E_log_1_H ||
        !probs || vbcounts_alloc(&cnt, F, Kc, N))
        return -1;

    vb_matrix_to_tensor(Z, F, N, Kc, E_Z); /* :406-407 */
    compute_vbcounts(V, E_Z, Kc, &cnt);    /* :411-412 */

    for (int i = 0; i < iter; i++) { /* :415 */
        for (int f = 0; f < F; f++) /* :419-421 */
            for (int k = 0; k < Kc; k++)
                gamma_vb[f + (size_t)F * k] = gamma + cnt.n_total[f + (size_t)F * k];
        for (int n = 0; n < N; n++) /* :422-425 */
            for (int k = 0; k < Kc; k++) {
                alpha_vb[n + (size_t)N * k] = alpha + cnt.f_ones[n + (size_t)N * k];
                beta_vb[n + (size_t)N * k] = beta + cnt.f_zeros[n + (size_t)N * k];
            }
        for (int f = 0; f < F; f++) { /* q(W) :428-434 */
            double s = 0.0;
            for (int k = 0; k < Kc; k++) s += gamma_vb[f + (size_t)F * k];
            double ps = oracle_digamma(s);
            for (int k = 0; k < Kc; k++)
                E_log_W[f + (size_t)F * k] = oracle_digamma(gamma_vb[f + (size_t)F * k]) - ps;
        }
        for (int n = 0; n < N; n++) /* q(H) :437-444 */
            for (int k = 0; k < Kc; k++) {
                size_t j = n + (size_t)N * k;
                double pab = oracle_digamma(alpha_vb[j] + beta_vb[j]);
                E_log_H[j] = oracle_digamma(alpha_vb[j]) - pab;
                E_log_1_H[j] = oracle_digamma(beta_vb[j]) - pab;
            }
        for (int n = 0; n < N; n++) /* q(Z) :447-465 */
            for (int f = 0; f < F; f++) {
                int v = V[(size_t)f + (size_t)F * n];
                if (v != NA_INT) {
                    remove_vbassignment(f, n, E_Z, v, Kc, &cnt); /* :450 */
                    double s = 0.0;
                    for (int k = 0; k < Kc; k++) { /* :453-457 */
                        probs[k] = exp(E_log_W[f + (size_t)F * k] +
                                       v * E_log_H[n + (size_t)N * k] +
                                       (1 - v) * E_log_1_H[n + (size_t)N * k]);
                        s += probs[k];
                    }
                    for (int k = 0; k < Kc; k++) probs[k] /= s;
                    set_vbassignment(f, n, probs, E_Z, v, Kc, &cnt); /* :460 */
                } else {
                    for (int k = 0; k < Kc; k++) EZ(f, k, n) = 0.0; /* :462 */
                }
            }
        if (checklb) /* :469-477 */
            lowerbounds[i] = oracle_lower_bound_aspect(E_Z, E_log_W, E_log_H, E_log_1_H, gamma_vb,
                                                       alpha_vb, beta_vb, alpha, beta, gamma, V, F,
                                                       Kc, N, NULL);
        else
            lowerbounds[i] = 0.0;
    }
    /* :501-502, :507-518 */
    for (size_t j = 0; j < (size_t)N * Kc; j++)
        if (alpha_vb[j] < 0 || beta_vb[j] < 0) rc = 1;
    for (int f = 0; f < F; f++) {
        double s = 0.0;
        for (int k = 0; k < Kc; k++) s += gamma_vb[f + (size_t)F * k];
        for (int k = 0; k < Kc; k++) E_W[f + (size_t)F * k] = gamma_vb[f + (size_t)F * k] / s;
    }
    for (size_t j = 0; j < (size_t)N * Kc; j++) E_H[j] = alpha_vb[j] / (alpha_vb[j] + beta_vb[j]);
    if (E_Z_out) memcpy(E_Z_out, E_Z, sizeof(double) * (size_t)F * Kc * N);
    free(E_Z); free(gamma_vb); free(alpha_vb); free(beta_vb); free(E_log_W); free(E_log_H);
    free(E_log_1_H); free(probs); vbcounts_free(&cnt);
    return rc;
}

/* ------------------------------------------------------------------------ */
/* E_Z-free batch form of VB_Aspect_Rcpp: the same     */
/* iteration written as the contraction the engine computes.  Used where the */
/* F x K x N cube no longer fits; tests prove it equal to oracle_vb_aspect.   */
/* stats_io (optional): n_total F x K, f_ones N x K, f_zeros N x K; when     */
/* use_stats_in != 0 the statistics are read from it instead of being        */
/* initialised from Z.  On return it holds the statistics after the last     */
/* Z-update.                                                                 */
/* ------------------------------------------------------------------------ */
int oracle_vb_aspect_batch(const int32_t *V, const int32_t *Z, int F, int N, int K, double alpha,
                           double beta, double gamma, int iter, int checklb, double *E_W,
                           double *E_H, double *lowerbounds, double *stats_io, int use_stats_in)
{
    size_t FK = (size_t)F * K, NK = (size_t)N * K;
    double *nt = calloc(FK, sizeof(double)), *fo = calloc(NK, sizeof(double)),
           *fz = calloc(NK, sizeof(double));
    double *g = malloc(sizeof(double) * FK), *a = malloc(sizeof(double) * NK),
           *b = malloc(sizeof(double) * NK);
    double *A = malloc(sizeof(double) * FK), *B1 = malloc(sizeof(double) * NK),
           *B0 = malloc(sizeof(double) * NK);
    double *nt2 = malloc(sizeof(double) * FK), *fo2 = malloc(sizeof(double) * NK),
           *fz2 = malloc(sizeof(double) * NK);
    double *elw = malloc(sizeof(double) * FK), *elh = malloc(sizeof(double) * NK),
           *el1h = malloc(sizeof(double) * NK);
    if (!nt || !fo || !fz || !g || !a || !b || !A || !B1 || !B0 || !nt2 || !fo2 || !fz2 || !elw ||
        !elh || !el1h)
        return -1;
    if (use_stats_in && stats_io) {
        memcpy(nt, stats_io, sizeof(double) * FK);
        memcpy(fo, stats_io + FK, sizeof(double) * NK);
        memcpy(fz, stats_io + FK + NK, sizeof(double) * NK);
    } else {
        for (int n = 0; n < N; n++)
            for (int f = 0; f < F; f++) {
                int v = V[(size_t)f + (size_t)F * n], k = Z[(size_t)f + (size_t)F * n];
                if (v == NA_INT || k == NA_INT) continue;
                nt[f + (size_t)F * k] += 1.0;
                if (v) fo[n + (size_t)N * k] += 1.0; else fz[n + (size_t)N * k] += 1.0;
            }
    }
    for (int i = 0; i < iter; i++) {
        for (size_t j = 0; j < FK; j++) g[j] = gamma + nt[j];
        for (size_t j = 0; j < NK; j++) { a[j] = alpha + fo[j]; b[j] = beta + fz[j]; }
        for (int f = 0; f < F; f++) {
            double s = 0.0;
            for (int k = 0; k < K; k++) s += g[f + (size_t)F * k];
            double ps = oracle_digamma(s);
            for (int k = 0; k < K; k++) {
                elw[f + (size_t)F * k] = oracle_digamma(g[f + (size_t)F * k]) - ps;
                A[f + (size_t)F * k] = exp(elw[f + (size_t)F * k]);
            }
        }
        for (size_t j = 0; j < NK; j++) {
            double pab = oracle_digamma(a[j] + b[j]);
            elh[j] = oracle_digamma(a[j]) - pab; el1h[j] = oracle_digamma(b[j]) - pab;
            B1[j] = exp(elh[j]); B0[j] = exp(el1h[j]);
        }
        memset(nt2, 0, sizeof(double) * FK); memset(fo2, 0, sizeof(double) * NK);
        memset(fz2, 0, sizeof(double) * NK);
        double sumlogD = 0.0;
        for (int n = 0; n < N; n++)
            for (int f = 0; f < F; f++) {
                int v = V[(size_t)f + (size_t)F * n];
                if (v == NA_INT) continue;
                const double *Bv = v ? B1 : B0;
                double D = 0.0;
                for (int k = 0; k < K; k++) D += A[f + (size_t)F * k] * Bv[n + (size_t)N * k];
                double R = 1.0 / D;
                sumlogD += log(D);
                double *fv = v ? fo2 : fz2;
                for (int k = 0; k < K; k++) {
                    double p = A[f + (size_t)F * k] * Bv[n + (size_t)N * k] * R;
                    nt2[f + (size_t)F * k] += p;
                    fv[n + (size_t)N * k] += p;
                }
            }
        if (checklb) {
            /* pZ + pV - qZ == sum_obs log D at the softmax optimum (Appendix A) */
            lowerbounds[i] = oracle_E_log_p_W(elw, F, K, gamma) - oracle_E_log_q_W(elw, g, F, K) +
                             oracle_E_log_p_H(elh, el1h, N, K, alpha, beta) -
                             oracle_E_log_q_H(elh, el1h, a, b, N, K) + sumlogD;
        } else if (lowerbounds) lowerbounds[i] = 0.0;
        memcpy(nt, nt2, sizeof(double) * FK); memcpy(fo, fo2, sizeof(double) * NK);
        memcpy(fz, fz2, sizeof(double) * NK);
    }
    for (int f = 0; f < F; f++) {
        double s = 0.0;
        for (int k = 0; k < K; k++) s += g[f + (size_t)F * k];
        for (int k = 0; k < K; k++) E_W[f + (size_t)F * k] = g[f + (size_t)F * k] / s;
    }
    for (size_t j = 0; j < NK; j++) E_H[j] = a[j] / (a[j] + b[j]);
    if (stats_io) {
        memcpy(stats_io, nt, sizeof(double) * FK);
        memcpy(stats_io + FK, fo, sizeof(double) * NK);
        memcpy(stats_io + FK + NK, fz, sizeof(double) * NK);
    }
    free(nt); free(fo); free(fz); free(g); free(a); free(b); free(A); free(B1); free(B0);
    free(nt2); free(fo2); free(fz2); free(elw); free(elh); free(el1h);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* VB_DirBer_Rcpp  (collapsed_gibbs_z_VB.cpp:97-168) -- sequential CVB0.     */
/* Kc = K+1 (:101); iter-1 sweeps (:119); E_W / E_H from the stale rows of   */
/* gamma_vb / alpha_vb / beta_vb (:151-159).  Rows / columns that never see  */
/* an observed entry are uninitialised in the reference; here they are 0     */
/* (=> 0/0 = NaN in E_W / E_H), documented in DESIGN.md.                     */
/* ------------------------------------------------------------------------ */
int oracle_vb_dirber(const int32_t *V, const int32_t *Z, int F, int N, int K, double alpha,
                     double beta, double gamma, int iter, double *E_W, double *E_H,
                     double *E_Z_out)
{
    const int Kc = K + 1;
    vbcounts cnt;
    int rc = 0;
    double *E_Z = malloc(sizeof(double) * (size_t)F * Kc * N);
    double *gamma_vb = calloc((size_t)F * Kc, sizeof(double));
    double *alpha_vb = calloc((size_t)N * Kc, sizeof(double));
    double *beta_vb = calloc((size_t)N * Kc, sizeof(double));
    double *probs = malloc(sizeof(double) * Kc);
    if (!E_Z || !gamma_vb || !alpha_vb || !beta_vb || !probs || vbcounts_alloc(&cnt, F, Kc, N))
        return -1;
    vb_matrix_to_tensor(Z, F, N, Kc, E_Z); /* :110-111 */
    compute_vbcounts(V, E_Z, Kc, &cnt);    /* :115-116 */
    for (int i = 1; i < iter; i++)         /* :119 */
        for (int n = 0; n < N; n++)
            for (int f = 0; f < F; f++) {
                int v = V[(size_t)f + (size_t)F * n];
                if (v == NA_INT) continue;
                remove_vbassignment(f, n, E_Z, v, Kc, &cnt); /* :125 */
                double s = 0.0;
                for (int k = 0; k < Kc; k++) { /* :128-133 */
                    double gv = gamma + cnt.n_total[f + (size_t)F * k];
                    double av = alpha + cnt.f_ones[n + (size_t)N * k];
                    double bv = beta + cnt.f_zeros[n + (size_t)N * k];
                    gamma_vb[f + (size_t)F * k] = gv;
                    alpha_vb[n + (size_t)N * k] = av;
                    beta_vb[n + (size_t)N * k] = bv;
                    probs[k] = gv * (v * av + (1 - v) * bv) / (av + bv);
                    s += probs[k];
                }
                for (int k = 0; k < Kc; k++) probs[k] /= s; /* :134 */
                set_vbassignment(f, n, probs, E_Z, v, Kc, &cnt); /* :137 */
            }
    for (size_t j = 0; j < (size_t)N * Kc; j++)
        if (alpha_vb[j] < 0 || beta_vb[j] < 0) rc = 1;
    for (int f = 0; f < F; f++) { /* :154-156 */
        double s = 0.0;
        for (int k = 0; k < Kc; k++) s += gamma_vb[f + (size_t)F * k];
        for (int k = 0; k < Kc; k++) E_W[f + (size_t)F * k] = gamma_vb[f + (size_t)F * k] / s;
    }
    for (size_t j = 0; j < (size_t)N * Kc; j++) /* :157-159 */
        E_H[j] = alpha_vb[j] / (alpha_vb[j] + beta_vb[j]);
    if (E_Z_out) memcpy(E_Z_out, E_Z, sizeof(double) * (size_t)F * Kc * N);
    free(E_Z); free(gamma_vb); free(alpha_vb); free(beta_vb); free(probs); vbcounts_free(&cnt);
    return rc;
}

/* ------------------------------------------------------------------------ */
/* integer counts (collapsed_gibbs_z.cpp:11-85)                              */
/* ------------------------------------------------------------------------ */
typedef struct {
    int F, N, K;
    int32_t *f_ones, *f_zeros, *f_total, *n_total, *fn_total;
} icounts;

static int icounts_alloc(icounts *c, int F, int K, int N)
{
    c->F = F; c->N = N; c->K = K;
    c->f_ones = calloc((size_t)N * K, 4); c->f_zeros = calloc((size_t)N * K, 4);
    c->f_total = calloc((size_t)N * K, 4); c->n_total = calloc((size_t)F * K, 4);
    c->fn_total = calloc((size_t)K, 4);
    return (c->f_ones && c->f_zeros && c->f_total && c->n_total && c->fn_total) ? 0 : -1;
}
static void icounts_free(icounts *c)
{
    free(c->f_ones); free(c->f_zeros); free(c->f_total); free(c->n_total); free(c->fn_total);
}
/* compute_counts (:38-60) */
static void compute_counts(const int32_t *V, const int32_t *Z, icounts *c)
{
    int F = c->F, N = c->N, K = c->K;
    memset(c->f_ones, 0, 4 * (size_t)N * K); memset(c->f_zeros, 0, 4 * (size_t)N * K);
    memset(c->f_total, 0, 4 * (size_t)N * K); memset(c->n_total, 0, 4 * (size_t)F * K);
    memset(c->fn_total, 0, 4 * (size_t)K);
    for (int n = 0; n < N; n++)
        for (int f = 0; f < F; f++) {
            int v = V[(size_t)f + (size_t)F * n];
            if (v == NA_INT) continue;
            int k = Z[(size_t)f + (size_t)F * n];
            c->f_ones[n + (size_t)N * k] += v;
            c->f_zeros[n + (size_t)N * k] += 1 - v;
            c->f_total[n + (size_t)N * k] += 1;
            c->n_total[f + (size_t)F * k] += 1;
            c->fn_total[k] += 1;
        }
}

/* ------------------------------------------------------------------------ */
/* Gibbs_DirBer_finite_Rcpp  (collapsed_gibbs_z.cpp:363-437) -- collapsed    */
/* sweep.  Z is updated in place.  Z_samples: F x N x nsamples with          */
/* nsamples = iter - floor(iter*burnin) slices, zero-filled, of which only   */
/* those with i >= nburnin, i in 1..iter-1, are written (Appendix B.10).     */
/* Returns the number of slices written, or -1.                              */
/* ------------------------------------------------------------------------ */
int oracle_gibbs_dirber_finite(const int32_t *V, int32_t *Z, int F, int N, int K, double alpha,
                               double beta, double gamma, int iter, double burnin, uint64_t seed,
                               int32_t *Z_samples)
{
    icounts cnt;
    orng g;
    int nburnin = (int)floor(iter * burnin);
    int nsamples = iter - (int)floor(iter * burnin);
    int j = 0;
    double *probs = malloc(sizeof(double) * K);
    if (!probs || icounts_alloc(&cnt, F, K, N)) return -1;
    orng_seed(&g, seed);
    if (Z_samples) memset(Z_samples, 0, 4 * (size_t)F * N * (nsamples > 0 ? nsamples : 0));
    compute_counts(V, Z, &cnt); /* :388-389 */
    for (int i = 1; i < iter; i++) { /* :392 */
        for (int n = 0; n < N; n++)
            for (int f = 0; f < F; f++) {
                size_t e = (size_t)f + (size_t)F * n;
                int v = V[e];
                if (v == NA_INT) continue;
                int k0 = Z[e]; /* remove_assignment :75-85 */
                cnt.f_ones[n + (size_t)N * k0] -= v; cnt.f_zeros[n + (size_t)N * k0] -= 1 - v;
                cnt.f_total[n + (size_t)N * k0] -= 1; cnt.n_total[f + (size_t)F * k0] -= 1;
                cnt.fn_total[k0] -= 1;
                double s = 0.0;
                for (int k = 0; k < K; k++) { /* :402-407 */
                    double gp = gamma + cnt.n_total[f + (size_t)F * k];
                    double ap = alpha + cnt.f_ones[n + (size_t)N * k];
                    double bp = beta + cnt.f_zeros[n + (size_t)N * k];
                    probs[k] = gp * (v * ap + (1 - v) * bp) / (ap + bp);
                    s += probs[k];
                }
                for (int k = 0; k < K; k++) probs[k] /= s; /* :408 */
                int k1 = orng_categorical(&g, probs, K);   /* :411-413 */
                Z[e] = k1; /* set_assignment :63-72 */
                cnt.f_ones[n + (size_t)N * k1] += v; cnt.f_zeros[n + (size_t)N * k1] += 1 - v;
                cnt.f_total[n + (size_t)N * k1] += 1; cnt.n_total[f + (size_t)F * k1] += 1;
                cnt.fn_total[k1] += 1;
            }
        if (i >= nburnin && Z_samples && j < nsamples) { /* :423-426 */
            memcpy(Z_samples + (size_t)F * N * j, Z, 4 * (size_t)F * N);
            j++;
        }
    }
    free(probs); icounts_free(&cnt);
    return j;
}

/* ------------------------------------------------------------------------ */
/* The same sweep with a counter-based uniform per entry instead of a        */
/* sequential stream: u(f, n, i) = Philox4x32-10 keyed (seed; f, n, i)       */
/* (Salmon et al. 2011).  Loop order, arithmetic and bookkeeping are those   */
/* of Gibbs_DirBer_finite_Rcpp (collapsed_gibbs_z.cpp:392-427); only the     */
/* source of the uniform differs, which makes the chain independent of the   */
/* traversal order -- the engine's wavefront kernel must reproduce Z and     */
/* Z_samples bit for bit.                                                    */
/* ------------------------------------------------------------------------ */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += W0; k1 += W1;
    }
}

/* one entry of the sweep (:396-418) */
static void gibbs_keyed_entry(const int32_t *V, int32_t *Z, icounts *cnt, double *probs, int F, int N, int K,
                              double alpha, double beta, double gamma, uint64_t seed, int i, int f, int n)
{
    size_t e = (size_t)f + (size_t)F * n;
    int v = V[e];
    if (v == NA_INT) return;
    int k0 = Z[e]; /* remove_assignment :75-85 */
    cnt->f_ones[n + (size_t)N * k0] -= v; cnt->f_zeros[n + (size_t)N * k0] -= 1 - v;
    cnt->f_total[n + (size_t)N * k0] -= 1; cnt->n_total[f + (size_t)F * k0] -= 1;
    cnt->fn_total[k0] -= 1;
    double s = 0.0;
    for (int k = 0; k < K; k++) { /* :402-407 */
        double gp = gamma + cnt->n_total[f + (size_t)F * k];
        double ap = alpha + cnt->f_ones[n + (size_t)N * k];
        double bp = beta + cnt->f_zeros[n + (size_t)N * k];
        probs[k] = gp * (v * ap + (1 - v) * bp) / (ap + bp);
        s += probs[k];
    }
    for (int k = 0; k < K; k++) probs[k] /= s; /* :408 */
    uint32_t c[4] = {(uint32_t)f, (uint32_t)n, (uint32_t)i, 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    uint64_t bits = ((uint64_t)c[0] << 32) | c[1];
    double u = ((double)(bits >> 11) + 0.5) * (1.0 / 9007199254740992.0), acc = 0.0;
    int k1 = -1; /* :411-413 */
    for (int k = 0; k < K && k1 < 0; k++) { acc += probs[k]; if (u < acc) k1 = k; }
    if (k1 < 0) { k1 = K - 1; for (int k = K - 1; k >= 0; k--) if (probs[k] > 0.0) { k1 = k; break; } }
    Z[e] = k1; /* set_assignment :63-72 */
    cnt->f_ones[n + (size_t)N * k1] += v; cnt->f_zeros[n + (size_t)N * k1] += 1 - v;
    cnt->f_total[n + (size_t)N * k1] += 1; cnt->n_total[f + (size_t)F * k1] += 1;
    cnt->fn_total[k1] += 1;
}

/* wavefront = 0: the reference's column-major order (:394-395); 1: anti-diagonals d = f + n
 * (the engine's order; must give the same labels, tests/test_oracle.py)      */
static int gibbs_keyed(const int32_t *V, int32_t *Z, int F, int N, int K, double alpha, double beta, double gamma,
                       int iter, double burnin, uint64_t seed, int32_t *Z_samples, int wavefront)
{
    icounts cnt;
    int nburnin = (int)floor(iter * burnin);
    int nsamples = iter - (int)floor(iter * burnin);
    int j = 0;
    double *probs = malloc(sizeof(double) * K);
    if (!probs || icounts_alloc(&cnt, F, K, N)) return -1;
    if (Z_samples) memset(Z_samples, 0, 4 * (size_t)F * N * (nsamples > 0 ? nsamples : 0));
    compute_counts(V, Z, &cnt); /* :388-389 */
    for (int i = 1; i < iter; i++) { /* :392 */
        if (!wavefront) {
            for (int n = 0; n < N; n++)
                for (int f = 0; f < F; f++) gibbs_keyed_entry(V, Z, &cnt, probs, F, N, K, alpha, beta, gamma, seed, i, f, n);
        } else {
            for (int d = 0; d <= F + N - 2; d++) {
                int f_lo = d - (N - 1) > 0 ? d - (N - 1) : 0, f_hi = d < F - 1 ? d : F - 1;
                for (int f = f_hi; f >= f_lo; f--) /* any order within a diagonal */
                    gibbs_keyed_entry(V, Z, &cnt, probs, F, N, K, alpha, beta, gamma, seed, i, f, d - f);
            }
        }
        if (i >= nburnin && Z_samples && j < nsamples) { /* :423-426 */
            memcpy(Z_samples + (size_t)F * N * j, Z, 4 * (size_t)F * N);
            j++;
        }
    }
    free(probs); icounts_free(&cnt);
    return j;
}

int oracle_gibbs_dirber_finite_keyed(const int32_t *V, int32_t *Z, int F, int N, int K, double alpha,
                                     double beta, double gamma, int iter, double burnin, uint64_t seed,
                                     int32_t *Z_samples)
{
    return gibbs_keyed(V, Z, F, N, K, alpha, beta, gamma, iter, burnin, seed, Z_samples, 0);
}

int oracle_gibbs_dirber_finite_keyed_wavefront(const int32_t *V, int32_t *Z, int F, int N, int K, double alpha,
                                               double beta, double gamma, int iter, double burnin,
                                               uint64_t seed, int32_t *Z_samples)
{
    return gibbs_keyed(V, Z, F, N, K, alpha, beta, gamma, iter, burnin, seed, Z_samples, 1);
}

/* ------------------------------------------------------------------------ */
/* Sibling collapsed samplers with the same keyed uniform. */
/* ------------------------------------------------------------------------ */
/* stream != NULL: the next draw of the sequential stream (what the reference does with R's RNG and
 * the reference build does with the shim's); else the counter-based one */
static double keyed_unif(uint64_t seed, int f, int n, int i, int second);
static double next_unif(orng *stream, uint64_t seed, int f, int n, int i, int second)
{
    return stream ? orng_unif(stream) : keyed_unif(seed, f, n, i, second);
}
static double keyed_unif(uint64_t seed, int f, int n, int i, int second)
{
    uint32_t c[4] = {(uint32_t)f, (uint32_t)n, (uint32_t)i, 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    uint64_t bits = second ? (((uint64_t)c[2] << 32) | c[3]) : (((uint64_t)c[0] << 32) | c[1]);
    return ((double)(bits >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
/* probs / accu(probs), then the inverse-CDF draw that stands in for rmultinom + index_max */
static int normalise_and_draw(double *probs, int K, double u)
{
    double s = 0.0, acc = 0.0;
    for (int k = 0; k < K; k++) s += probs[k];
    for (int k = 0; k < K; k++) probs[k] /= s;
    for (int k = 0; k < K; k++) { acc += probs[k]; if (u < acc) return k; }
    for (int k = K - 1; k >= 0; k--) if (probs[k] > 0.0) return k;
    return K - 1;
}
static void iremove(icounts *c, int f, int n, int k, int v)
{   /* remove_assignment :75-85 */
    c->f_ones[n + (size_t)c->N * k] -= v; c->f_zeros[n + (size_t)c->N * k] -= 1 - v;
    c->f_total[n + (size_t)c->N * k] -= 1; c->n_total[f + (size_t)c->F * k] -= 1; c->fn_total[k] -= 1;
}
static void iset(icounts *c, int f, int n, int k, int v)
{   /* set_assignment :63-72 */
    c->f_ones[n + (size_t)c->N * k] += v; c->f_zeros[n + (size_t)c->N * k] += 1 - v;
    c->f_total[n + (size_t)c->N * k] += 1; c->n_total[f + (size_t)c->F * k] += 1; c->fn_total[k] += 1;
}

/* Gibbs_DirBer_DP_Rcpp (collapsed_gibbs_z.cpp:275-359): Kmax = 200 components (:282), the
 * Beta prior replaced by epsilon = 1e-8 on both counts (:281,317-318; alpha and beta are
 * unused), weight n_total (no gamma) per occupied component (:321) and gamma/2 on the first
 * empty one (:325-329).  The reference has no NA test (:308-310) and indexes probs[Kmax] when
 * no component is empty: here V must be NA-free and a full row simply opens no new component. */
static void dp_entry(const int32_t *V, int32_t *Z, icounts *cnt, double *probs, int F, int N, int K,
                     double gamma, orng *stream, uint64_t seed, int i, int f, int n)
{
    const double epsilon = 0.00000001;
    size_t e = (size_t)f + (size_t)F * n;
    int v = V[e];
    iremove(cnt, f, n, Z[e], v);
    int k_free = -1;
    for (int k = 0; k < K; k++) {
        double ap = cnt->f_ones[n + (size_t)N * k] + epsilon;
        double bp = cnt->f_zeros[n + (size_t)N * k] + epsilon;
        double bpr = (v * ap + (1 - v) * bp) / (ap + bp);
        probs[k] = cnt->n_total[f + (size_t)F * k] * bpr;
        if (k_free < 0 && cnt->n_total[f + (size_t)F * k] == 0) k_free = k;
    }
    if (k_free >= 0) probs[k_free] = gamma * 0.5;
    int k1 = normalise_and_draw(probs, K, next_unif(stream, seed, f, n, i, 0));
    Z[e] = k1;
    iset(cnt, f, n, k1, v);
}

/* Gibbs_DirDir_finite_Rcpp (collapsed_gibbs_z.cpp:441-557): two assignments per entry.
 * V = 1: Z and C drawn together from (gamma + n_total^Z_f) (alpha + f_total^C_n) (:481-497);
 * V = 0: Z from gamma + n_total^Z_f with C's component forbidden (:500-512), then C from
 * alpha + f_total^C_n with the new Z component forbidden (:514-531).                       */
static void dirdir_entry(const int32_t *V, int32_t *Z, int32_t *C, icounts *cz, icounts *cc, double *probs,
                         int F, int N, int K, double alpha, double gamma, orng *stream, uint64_t seed, int i, int f, int n)
{
    size_t e = (size_t)f + (size_t)F * n;
    int v = V[e];
    if (v == NA_INT) return;
    if (v) {
        iremove(cz, f, n, Z[e], v); iremove(cc, f, n, C[e], v);
        for (int k = 0; k < K; k++)
            probs[k] = (gamma + cz->n_total[f + (size_t)F * k]) * (alpha + cc->f_total[n + (size_t)N * k]);
        int k = normalise_and_draw(probs, K, next_unif(stream, seed, f, n, i, 0));
        Z[e] = k; C[e] = k;
        iset(cz, f, n, k, v); iset(cc, f, n, k, v);
    } else {
        iremove(cz, f, n, Z[e], v);
        for (int k = 0; k < K; k++) probs[k] = gamma + cz->n_total[f + (size_t)F * k];
        probs[C[e]] = 0;
        int kz = normalise_and_draw(probs, K, next_unif(stream, seed, f, n, i, 0));
        iremove(cc, f, n, C[e], v);
        for (int k = 0; k < K; k++) probs[k] = alpha + cc->f_total[n + (size_t)N * k];
        probs[kz] = 0;
        int kc = normalise_and_draw(probs, K, next_unif(stream, seed, f, n, i, 1));
        Z[e] = kz; C[e] = kc;
        iset(cz, f, n, kz, v); iset(cc, f, n, kc, v);
    }
}

/* variant 1 = DP (K is forced to 200, C ignored), 2 = Dir-Dir.  Same sweep bookkeeping as
 * Gibbs_DirBer_finite_Rcpp; wavefront as in gibbs_keyed. */
static int siblings_keyed(int variant, const int32_t *V, int32_t *Z, int32_t *C, int F, int N, int K, double alpha,
                          double gamma, int iter, double burnin, uint64_t seed, int32_t *Z_samples,
                          int32_t *C_samples, int wavefront, int use_stream)
{
    icounts cz, cc;
    orng g, *stream = use_stream ? &g : NULL;
    if (use_stream) { orng_seed(&g, seed); wavefront = 0; }
    int nburnin = (int)floor(iter * burnin);
    int nsamples = iter - (int)floor(iter * burnin);
    int j = 0;
    if (variant == 1) K = 200;
    double *probs = malloc(sizeof(double) * K);
    if (!probs || icounts_alloc(&cz, F, K, N) || icounts_alloc(&cc, F, K, N)) return -1;
    if (Z_samples) memset(Z_samples, 0, 4 * (size_t)F * N * (nsamples > 0 ? nsamples : 0));
    if (C_samples) memset(C_samples, 0, 4 * (size_t)F * N * (nsamples > 0 ? nsamples : 0));
    compute_counts(V, Z, &cz);
    if (variant == 2) { memcpy(C, Z, 4 * (size_t)F * N); compute_counts(V, C, &cc); } /* :465-473 */
    for (int i = 1; i < iter; i++) {
        for (int d = 0; d <= (wavefront ? F + N - 2 : N - 1); d++) {
            int f_lo = wavefront ? (d - (N - 1) > 0 ? d - (N - 1) : 0) : 0;
            int f_hi = wavefront ? (d < F - 1 ? d : F - 1) : F - 1;
            for (int f = f_lo; f <= f_hi; f++) {
                int n = wavefront ? d - f : d;
                if (variant == 1) dp_entry(V, Z, &cz, probs, F, N, K, gamma, stream, seed, i, f, n);
                else dirdir_entry(V, Z, C, &cz, &cc, probs, F, N, K, alpha, gamma, stream, seed, i, f, n);
            }
        }
        if (i >= nburnin && j < nsamples) {
            if (Z_samples) memcpy(Z_samples + (size_t)F * N * j, Z, 4 * (size_t)F * N);
            if (C_samples) memcpy(C_samples + (size_t)F * N * j, C, 4 * (size_t)F * N);
            j++;
        }
    }
    free(probs); icounts_free(&cz); icounts_free(&cc);
    return j;
}

/* wavefront: 0 column-major / 1 anti-diagonal with keyed uniforms; 2 = column-major on the sequential
 * stream (the form that is pinned against the reference's own sources, tests/test_oracle_pin.py) */
int oracle_gibbs_dirber_dp_keyed(const int32_t *V, int32_t *Z, int F, int N, double gamma, int iter,
                                 double burnin, uint64_t seed, int32_t *Z_samples, int wavefront)
{
    for (size_t e = 0; e < (size_t)F * N; e++) if (V[e] == NA_INT) return -2; /* undefined in the reference */
    return siblings_keyed(1, V, Z, NULL, F, N, 200, 0.0, gamma, iter, burnin, seed, Z_samples, NULL, wavefront == 1, wavefront == 2);
}

int oracle_gibbs_dirdir_keyed(const int32_t *V, int32_t *Z, int32_t *C, int F, int N, int K, double alpha,
                              double gamma, int iter, double burnin, uint64_t seed, int32_t *Z_samples,
                              int32_t *C_samples, int wavefront)
{
    return siblings_keyed(2, V, Z, C, F, N, K, alpha, gamma, iter, burnin, seed, Z_samples, C_samples, wavefront == 1, wavefront == 2);
}

/* ------------------------------------------------------------------------ */
/* expectation.GibbsDirBer  (R/generic_Gibbs_DirBer.R:54-101): per-sample    */
/* E_W (Dirichlet mean) and E_H (Beta mean) from the counts of sample j,     */
/* E_V = mean_j E_W E_H^T.  Also returns the across-sample means of E_W and  */
/* E_H (what compute_expectation_W_Rcpp :178-197 and a quirk-free            */
/* compute_expectation_H_Rcpp :230-272 give).  E_V may be NULL.              */
/* ------------------------------------------------------------------------ */
int oracle_gibbs_expectation(const int32_t *V, const int32_t *Z_samples, int F, int N, int K,
                             int J, double alpha, double beta, double gamma, double *E_W,
                             double *E_H, double *E_V)
{
    icounts cnt;
    double *w = malloc(sizeof(double) * (size_t)F * K), *h = malloc(sizeof(double) * (size_t)N * K);
    if (!w || !h || icounts_alloc(&cnt, F, K, N)) return -1;
    memset(E_W, 0, sizeof(double) * (size_t)F * K);
    memset(E_H, 0, sizeof(double) * (size_t)N * K);
    if (E_V) memset(E_V, 0, sizeof(double) * (size_t)F * N);
    for (int j = 0; j < J; j++) {
        compute_counts(V, Z_samples + (size_t)F * N * j, &cnt);
        for (int f = 0; f < F; f++) { /* R:82-85 */
            double s = 0.0;
            for (int k = 0; k < K; k++) s += gamma + cnt.n_total[f + (size_t)F * k];
            for (int k = 0; k < K; k++)
                w[f + (size_t)F * k] = (gamma + cnt.n_total[f + (size_t)F * k]) / s;
        }
        for (size_t i = 0; i < (size_t)N * K; i++) /* R:88-90 */
            h[i] = (alpha + cnt.f_ones[i]) / (alpha + beta + cnt.f_total[i]);
        for (size_t i = 0; i < (size_t)F * K; i++) E_W[i] += w[i] / J;
        for (size_t i = 0; i < (size_t)N * K; i++) E_H[i] += h[i] / J;
        if (E_V)
            for (int n = 0; n < N; n++)
                for (int f = 0; f < F; f++) { /* R:93 */
                    double s = 0.0;
                    for (int k = 0; k < K; k++) s += w[f + (size_t)F * k] * h[n + (size_t)N * k];
                    E_V[(size_t)f + (size_t)F * n] += s / J;
                }
    }
    free(w); free(h); icounts_free(&cnt);
    return 0;
}

/* compute_expectation_W_Rcpp (collapsed_gibbs_z.cpp:178-197): normalise
 * gamma + mean_j n_total_j over Kq = max label + 1 columns.                  */
int oracle_compute_expectation_W(const int32_t *Z_samples, int F, int N, int J, double gamma,
                                 int Kq, double *E_W)
{
    memset(E_W, 0, sizeof(double) * (size_t)F * Kq);
    for (int j = 0; j < J; j++)
        for (int n = 0; n < N; n++)
            for (int f = 0; f < F; f++) {
                int k = Z_samples[(size_t)f + (size_t)F * n + (size_t)F * N * j];
                if (k != NA_INT) E_W[f + (size_t)F * k] += 1.0 / J;
            }
    for (int f = 0; f < F; f++) {
        double s = 0.0;
        for (int k = 0; k < Kq; k++) { E_W[f + (size_t)F * k] += gamma; s += E_W[f + (size_t)F * k]; }
        for (int k = 0; k < Kq; k++) E_W[f + (size_t)F * k] /= s;
    }
    return 0;
}

/* compute_expectation_H_Rcpp (collapsed_gibbs_z.cpp:230-272) with its quirk:
 * the k loop stops at K = max label, so column Kq-1 stays 0 (:236-242).     */
int oracle_compute_expectation_H(const int32_t *Z_samples, const int32_t *V, int F, int N, int J,
                                 double alpha, double beta, int Kq, double *E_H)
{
    memset(E_H, 0, sizeof(double) * (size_t)N * Kq);
    for (int k = 0; k < Kq - 1; k++)
        for (int n = 0; n < N; n++) {
            double acc = 0.0;
            for (int j = 0; j < J; j++) {
                int ones = 0, tot = 0;
                for (int f = 0; f < F; f++) {
                    int v = V[(size_t)f + (size_t)F * n];
                    if (v == NA_INT) continue;
                    int z = Z_samples[(size_t)f + (size_t)F * n + (size_t)F * N * j];
                    ones += (z == k) * v; tot += (z == k);
                }
                acc += (alpha + ones) / (alpha + beta + tot);
            }
            E_H[n + (size_t)N * k] = acc / J;
        }
    return 0;
}

/* ------------------------------------------------------------------------ */
/* Uncollapsed (blocked) Gibbs sampler: the chain the engine runs.           */
/*   Z_fn | W,H ~ Cat(W_fk H_nk^v (1-H_nk)^(1-v))   (R/MCEM - sample_gibbs.R */
/*                 :40-77 sample_gibbs_ZH; sample_gibbs.cpp:43 likelihood)   */
/*   W_f | Z ~ Dir(gamma + n_total_f), H_nk | Z,V ~ Beta(alpha+ones,         */
/*                 beta+zeros)    (collapsed_gibbs_z.cpp:589-605)            */
/* iter-1 sweeps, keep those with i >= floor(iter*burnin) (same bookkeeping  */
/* as :392,423).  Accumulates the Rao-Blackwell means of                     */
/* expectation.GibbsDirBer over kept sweeps.  Returns kept count.            */
/* ------------------------------------------------------------------------ */
int oracle_gibbs_uncollapsed(const int32_t *V, int32_t *Z, int F, int N, int K, double alpha,
                             double beta, double gamma, int iter, double burnin, uint64_t seed,
                             double *E_W, double *E_H, double *E_V)
{
    icounts cnt;
    orng g;
    int nburnin = (int)floor(iter * burnin), kept = 0;
    double *W = malloc(sizeof(double) * (size_t)F * K), *H = malloc(sizeof(double) * (size_t)N * K);
    double *w = malloc(sizeof(double) * (size_t)F * K), *h = malloc(sizeof(double) * (size_t)N * K);
    double *probs = malloc(sizeof(double) * K);
    if (!W || !H || !w || !h || !probs || icounts_alloc(&cnt, F, K, N)) return -1;
    orng_seed(&g, seed);
    memset(E_W, 0, sizeof(double) * (size_t)F * K);
    memset(E_H, 0, sizeof(double) * (size_t)N * K);
    if (E_V) memset(E_V, 0, sizeof(double) * (size_t)F * N);
    compute_counts(V, Z, &cnt);
    for (int i = 1; i < iter; i++) {
        for (int f = 0; f < F; f++) { /* W | Z :590-596 */
            double s = 0.0;
            for (int k = 0; k < K; k++) {
                W[f + (size_t)F * k] = orng_gamma(&g, gamma + cnt.n_total[f + (size_t)F * k]);
                s += W[f + (size_t)F * k];
            }
            for (int k = 0; k < K; k++) W[f + (size_t)F * k] /= s;
        }
        for (size_t j = 0; j < (size_t)N * K; j++) { /* H | Z,V :599-605 */
            double x = orng_gamma(&g, alpha + cnt.f_ones[j]);
            double y = orng_gamma(&g, beta + cnt.f_zeros[j]);
            H[j] = x / (x + y);
        }
        for (int n = 0; n < N; n++) /* Z | W,H */
            for (int f = 0; f < F; f++) {
                size_t e = (size_t)f + (size_t)F * n;
                int v = V[e];
                if (v == NA_INT) continue;
                double s = 0.0;
                for (int k = 0; k < K; k++) {
                    double hk = H[n + (size_t)N * k];
                    probs[k] = W[f + (size_t)F * k] * (v ? hk : 1.0 - hk);
                    s += probs[k];
                }
                for (int k = 0; k < K; k++) probs[k] /= s;
                Z[e] = orng_categorical(&g, probs, K);
            }
        compute_counts(V, Z, &cnt);
        if (i >= nburnin) {
            kept++;
            for (int f = 0; f < F; f++) {
                double s = 0.0;
                for (int k = 0; k < K; k++) s += gamma + cnt.n_total[f + (size_t)F * k];
                for (int k = 0; k < K; k++)
                    w[f + (size_t)F * k] = (gamma + cnt.n_total[f + (size_t)F * k]) / s;
            }
            for (size_t j = 0; j < (size_t)N * K; j++)
                h[j] = (alpha + cnt.f_ones[j]) / (alpha + beta + cnt.f_total[j]);
            for (size_t j = 0; j < (size_t)F * K; j++) E_W[j] += w[j];
            for (size_t j = 0; j < (size_t)N * K; j++) E_H[j] += h[j];
            if (E_V)
                for (int n = 0; n < N; n++)
                    for (int f = 0; f < F; f++) {
                        double s = 0.0;
                        for (int k = 0; k < K; k++)
                            s += w[f + (size_t)F * k] * h[n + (size_t)N * k];
                        E_V[(size_t)f + (size_t)F * n] += s;
                    }
        }
    }
    if (kept > 0) {
        for (size_t j = 0; j < (size_t)F * K; j++) E_W[j] /= kept;
        for (size_t j = 0; j < (size_t)N * K; j++) E_H[j] /= kept;
        if (E_V) for (size_t j = 0; j < (size_t)F * N; j++) E_V[j] /= kept;
    }
    free(W); free(H); free(w); free(h); free(probs); icounts_free(&cnt);
    return kept;
}

/* ------------------------------------------------------------------------ */
/* loglikelihood.Bernoulli / expectation.VBBernoulli (R/commons.R:17-33):    */
/* sum over non-NA test entries of log(v E_V + (1-v)(1-E_V)), E_V = E_W E_H^T*/
/* ------------------------------------------------------------------------ */
double oracle_loglikelihood(const double *E_W, const double *E_H, const int32_t *V_test, int F,
                            int N, int K, int64_t *n_test)
{
    double ll = 0.0;
    int64_t cntt = 0;
    for (int n = 0; n < N; n++)
        for (int f = 0; f < F; f++) {
            int v = V_test[(size_t)f + (size_t)F * n];
            if (v == NA_INT) continue;
            double p = 0.0;
            for (int k = 0; k < K; k++) p += E_W[f + (size_t)F * k] * E_H[n + (size_t)N * k];
            ll += log(v * p + (1 - v) * (1 - p));
            cntt++;
        }
    if (n_test) *n_test = cntt;
    return ll;
}

/* same, from a dense E_V (Gibbs: R/generic_Gibbs_DirBer.R:54-101 feeds commons.R:17-24) */
double oracle_loglikelihood_EV(const double *E_V, const int32_t *V_test, int F, int N)
{
    double ll = 0.0;
    for (size_t e = 0; e < (size_t)F * N; e++) {
        int v = V_test[e];
        if (v == NA_INT) continue;
        ll += log(v * E_V[e] + (1 - v) * (1 - E_V[e]));
    }
    return ll;
}

/* ------------------------------------------------------------------------ */
/* Synthetic generator (R/utils.R:7-16 sample_V; experiments/                */
/* xp_paper_basic_example.R:12-30): W_f ~ Dir(a 1_Kt), H_nk ~ Beta(a,b),     */
/* V_fn ~ Bernoulli(W_f . H_n).  W is F x Kt, H is N x Kt (code orientation).*/
/* ------------------------------------------------------------------------ */
int oracle_generate(int F, int N, int Kt, double a, double b, uint64_t seed, int32_t *V, double *W,
                    double *H)
{
    orng g;
    orng_seed(&g, seed);
    double *Wl = W ? W : malloc(sizeof(double) * (size_t)F * Kt);
    double *Hl = H ? H : malloc(sizeof(double) * (size_t)N * Kt);
    if (!Wl || !Hl) return -1;
    for (int f = 0; f < F; f++) {
        double s = 0.0;
        for (int k = 0; k < Kt; k++) { Wl[f + (size_t)F * k] = orng_gamma(&g, a); s += Wl[f + (size_t)F * k]; }
        for (int k = 0; k < Kt; k++) Wl[f + (size_t)F * k] /= s;
    }
    for (size_t j = 0; j < (size_t)N * Kt; j++) {
        double x = orng_gamma(&g, a), y = orng_gamma(&g, b);
        Hl[j] = x / (x + y);
    }
    for (int n = 0; n < N; n++)
        for (int f = 0; f < F; f++) {
            double p = 0.0;
            for (int k = 0; k < Kt; k++) p += Wl[f + (size_t)F * k] * Hl[n + (size_t)N * k];
            V[(size_t)f + (size_t)F * n] = orng_unif(&g) < p ? 1 : 0;
        }
    if (!W) free(Wl);
    if (!H) free(Hl);
    return 0;
}

/* wavefront restatement of VB_DirBer_Rcpp: entries ordered by anti-diagonal
 * d = f+n (SURVEY.md Appendix A proves it exact).  Same arithmetic as
 * oracle_vb_dirber, different traversal; tests assert bit equality.         */
int oracle_vb_dirber_wavefront(const int32_t *V, const int32_t *Z, int F, int N, int K,
                               double alpha, double beta, double gamma, int iter, double *E_W,
                               double *E_H)
{
    const int Kc = K + 1;
    vbcounts cnt;
    double *E_Z = malloc(sizeof(double) * (size_t)F * Kc * N);
    double *gamma_vb = calloc((size_t)F * Kc, sizeof(double));
    double *alpha_vb = calloc((size_t)N * Kc, sizeof(double));
    double *beta_vb = calloc((size_t)N * Kc, sizeof(double));
    double *probs = malloc(sizeof(double) * Kc);
    if (!E_Z || !gamma_vb || !alpha_vb || !beta_vb || !probs || vbcounts_alloc(&cnt, F, Kc, N))
        return -1;
    vb_matrix_to_tensor(Z, F, N, Kc, E_Z);
    compute_vbcounts(V, E_Z, Kc, &cnt);
    for (int i = 1; i < iter; i++)
        for (int d = 0; d <= F + N - 2; d++) {
            int f_lo = d - (N - 1) > 0 ? d - (N - 1) : 0, f_hi = d < F - 1 ? d : F - 1;
            for (int f = f_hi; f >= f_lo; f--) { /* any order inside a diagonal */
                int n = d - f;
                int v = V[(size_t)f + (size_t)F * n];
                if (v == NA_INT) continue;
                remove_vbassignment(f, n, E_Z, v, Kc, &cnt);
                double s = 0.0;
                for (int k = 0; k < Kc; k++) {
                    double gv = gamma + cnt.n_total[f + (size_t)F * k];
                    double av = alpha + cnt.f_ones[n + (size_t)N * k];
                    double bv = beta + cnt.f_zeros[n + (size_t)N * k];
                    gamma_vb[f + (size_t)F * k] = gv;
                    alpha_vb[n + (size_t)N * k] = av;
                    beta_vb[n + (size_t)N * k] = bv;
                    probs[k] = gv * (v * av + (1 - v) * bv) / (av + bv);
                    s += probs[k];
                }
                for (int k = 0; k < Kc; k++) probs[k] /= s;
                set_vbassignment(f, n, probs, E_Z, v, Kc, &cnt);
            }
        }
    for (int f = 0; f < F; f++) {
        double s = 0.0;
        for (int k = 0; k < Kc; k++) s += gamma_vb[f + (size_t)F * k];
        for (int k = 0; k < Kc; k++) E_W[f + (size_t)F * k] = gamma_vb[f + (size_t)F * k] / s;
    }
    for (size_t j = 0; j < (size_t)N * Kc; j++) E_H[j] = alpha_vb[j] / (alpha_vb[j] + beta_vb[j]);
    free(E_Z); free(gamma_vb); free(alpha_vb); free(beta_vb); free(probs); vbcounts_free(&cnt);
    return 0;
}
