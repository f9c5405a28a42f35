// kernels_basic.cuh -- packing, histogram, parameter, FP64 contraction, CVB0
// wavefront, Gibbs and predictive-LL kernels of libnbmf_cuda (sm_100a).
// See common.cuh for the HBM layout.  Reference citations are relative to
// /root/reference.
#pragma once
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

// ---------------------------------------------------------------------------
// packing: int32 column-major chunk (NA = INT_MIN) -> value / mask bit planes
// packed along f.  One warp per (column, word).  bad[0] counts entries that are
// neither 0, 1 nor NA (the engine is binary; the reference would silently use
// them as numbers).  nobs[0] accumulates the observed count.
// ---------------------------------------------------------------------------
__global__ void pack_cols_kernel(const int32_t *__restrict__ V, int64_t F, int64_t ncols,
                                 int64_t n0, int64_t Fw, uint32_t *__restrict__ vc,
                                 uint32_t *__restrict__ mc, unsigned long long *nobs,
                                 unsigned long long *bad)
{
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    int64_t total = ncols * Fw;
    if (warp >= total) return;
    int64_t c = warp / Fw, w = warp % Fw;
    int64_t f = w * 32 + lane;
    int32_t v = NBMF_NA_INTEGER;
    if (f < F) v = V[f + F * c];
    bool obs = (v != NBMF_NA_INTEGER);
    bool isbad = obs && (v != 0 && v != 1);
    uint32_t vw = __ballot_sync(0xffffffffu, obs && v == 1);
    uint32_t mw = __ballot_sync(0xffffffffu, obs);
    uint32_t bw = __ballot_sync(0xffffffffu, isbad);
    if (lane == 0) {
        vc[(n0 + c) * Fw + w] = vw;
        mc[(n0 + c) * Fw + w] = mw;
        if (mw) atomicAdd(nobs, (unsigned long long)__popc(mw));
        if (bw) atomicAdd(bad, (unsigned long long)__popc(bw));
    }
}

// 32x32 bit transpose across a warp: lane l holds row l (bit c = element (l, c)) and ends up
// with column l (bit r = element (r, l)).  Five butterfly steps, each swapping the off-diagonal
// j x j blocks between lanes l and l ^ j.
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, int lane)
{
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        // m: j ones, j zeros, ... from the least significant bit
        const uint32_t m = (j == 16) ? 0x0000ffffu : (j == 8) ? 0x00ff00ffu : (j == 4) ? 0x0f0f0f0fu
                         : (j == 2) ? 0x33333333u : 0x55555555u;
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
        const bool upper = (lane & j) == 0;
        const uint32_t t = upper ? (y << j) : (y >> j);
        const uint32_t keep = upper ? m : ~m;
        x = (x & keep) | (t & ~keep);
    }
    return x;
}

// Bit-plane transpose: src [R][Ws] packed along c  ->  dst [C][Wd] packed along r.
// One block = 8 row groups (256 rows) x 16 source words: the tile is read as 64-byte row
// segments, every warp transposes the 8 32x32 bit tiles of one source word, and the result
// leaves as 32-byte runs (8 consecutive destination words per destination row).
#define TB_GROUPS 8
#define TB_WORDS 16
__global__ void __launch_bounds__(TB_WORDS * 32)
transpose_bits_kernel(const uint32_t *__restrict__ src, int64_t R, int64_t Ws,
                      int64_t C, uint32_t *__restrict__ dst, int64_t Wd)
{
    __shared__ uint32_t tin[TB_GROUPS * 32][TB_WORDS + 1];
    __shared__ uint32_t tout[TB_WORDS * 32][TB_GROUPS + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t cw = (C + 31) / 32;
    const int64_t w0 = (int64_t)blockIdx.x * TB_WORDS, g0 = (int64_t)blockIdx.y * TB_GROUPS;
    for (int i = threadIdx.x; i < TB_GROUPS * 32 * TB_WORDS; i += TB_WORDS * 32) {
        const int rr = i / TB_WORDS, ww = i % TB_WORDS;
        const int64_t r = g0 * 32 + rr, w = w0 + ww;
        tin[rr][ww] = (r < R && w < cw) ? src[r * Ws + w] : 0u;
    }
    __syncthreads();
    // warp = source word w0 + warp
#pragma unroll
    for (int g = 0; g < TB_GROUPS; g++)
        tout[warp * 32 + lane][g] = warp_transpose32(tin[g * 32 + lane][warp], lane);
    __syncthreads();
    const int64_t gw = (R + 31) / 32;
    for (int i = threadIdx.x; i < TB_WORDS * 32 * TB_GROUPS; i += TB_WORDS * 32) {
        const int row = i / TB_GROUPS, g = i % TB_GROUPS;
        const int64_t c = w0 * 32 + row;
        if (c < C && g0 + g < gw) dst[c * Wd + g0 + g] = tout[row][g];
    }
}

// a bounded pause on the stream (no flag is polled): lets a kernel that another stream has just
// been allowed to start get its CTAs placed before the next persistent kernel takes every SM
__global__ void pause_kernel(unsigned ns)
{
    const long long t0 = clock64();
    for (unsigned waited = 0; waited < ns; waited += 1000) {
        __nanosleep(1000);
        if (clock64() - t0 > (long long)ns * 4) break;   // never more than ~ns at >= 250 MHz
    }
}

__global__ void popcount_kernel(const uint32_t *__restrict__ m, int64_t count, unsigned long long *out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned c = (i < count) ? __popc(m[i]) : 0u;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (unsigned long long)c);
}

// fill the mask planes for data without NA: bit set iff index < limit
__global__ void fill_mask_kernel(uint32_t *m, int64_t rows, int64_t words, int64_t limit)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * words) return;
    int64_t w = i % words;
    int64_t lo = w * 32;
    uint32_t v = 0;
    if (lo + 32 <= limit) v = 0xffffffffu;
    else if (lo < limit) v = (1u << (limit - lo)) - 1u;
    m[i] = v;
}

// ---------------------------------------------------------------------------
// initial statistics = histogram of the labels (vb_matrix_to_tensor +
// compute_vbcounts, commons.cpp:29-42, collapsed_gibbs_z_VB.cpp:42-66).
// ---------------------------------------------------------------------------
__global__ void hist_labels_kernel(const int32_t *__restrict__ Z, const uint32_t *__restrict__ vc,
                                   const uint32_t *__restrict__ mc, int64_t F, int64_t ncols,
                                   int64_t n0, int64_t Fw, int Kc, int Kp, double *statF,
                                   double *statN, unsigned long long *bad)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * ncols) return;
    int64_t f = i % F, c = i / F, n = n0 + c;
    uint32_t mw = mc ? mc[n * Fw + (f >> 5)] : 0xffffffffu;
    if (!((mw >> (f & 31)) & 1u)) return;
    int32_t k = Z[i];
    if (k == NBMF_NA_INTEGER) return;
    if (k < 0 || k >= Kc) { atomicAdd(bad, 1ull); return; }
    int v = (vc[n * Fw + (f >> 5)] >> (f & 31)) & 1u;
    atomicAdd(&statF[f * Kp + k], 1.0);
    atomicAdd(&statN[(2 * n + v) * Kp + k], 1.0);
}

__device__ __forceinline__ uint32_t label_hash(uint64_t seed, int64_t f, int64_t ng)
{
    philox4 r = philox4x32_10((uint32_t)(f >> 2), (uint32_t)ng, (uint32_t)(ng >> 32), 0x5A17u,
                              (uint32_t)seed, (uint32_t)(seed >> 32));
    int s = (int)(f & 3);
    return s == 0 ? r.x : s == 1 ? r.y : s == 2 ? r.z : r.w;
}

// random labels Z_fn = Philox(seed, f, global n) mod Kc, histogrammed without
// storing Z: one block per column builds (f_ones, f_zeros)(n,:) in shared
// memory; a second launch with one block per row builds n_total(f,:).
__global__ void hist_random_cols_kernel(const uint32_t *__restrict__ vc,
                                        const uint32_t *__restrict__ mc, int64_t F, int64_t Fw,
                                        int64_t n_offset, int Kc, int Kp, uint64_t seed,
                                        double *statN, int32_t *Zout)
{
    extern __shared__ int sh_hist[]; // [2][Kc]
    int64_t n = blockIdx.x;
    for (int i = threadIdx.x; i < 2 * Kc; i += blockDim.x) sh_hist[i] = 0;
    __syncthreads();
    for (int64_t f = threadIdx.x; f < F; f += blockDim.x) {
        uint32_t mw = mc ? mc[n * Fw + (f >> 5)] : 0xffffffffu;
        bool obs = (mw >> (f & 31)) & 1u;
        if (obs) {
            int v = (vc[n * Fw + (f >> 5)] >> (f & 31)) & 1u;
            int k = (int)(label_hash(seed, f, n + n_offset) % (uint32_t)Kc);
            atomicAdd(&sh_hist[v * Kc + k], 1);
            if (Zout) Zout[f + F * n] = k;
        } else if (Zout) Zout[f + F * n] = NBMF_NA_INTEGER;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * Kc; i += blockDim.x) {
        int v = i / Kc, k = i % Kc;
        statN[(2 * n + v) * Kp + k] = (double)sh_hist[i];
    }
}

__global__ void hist_random_rows_kernel(const uint32_t *__restrict__ mr, int64_t N, int64_t Nw,
                                        int64_t n_offset, int Kc, int Kp, uint64_t seed,
                                        double *statF)
{
    extern __shared__ int sh_hist[]; // [Kc]
    int64_t f = blockIdx.x;
    for (int i = threadIdx.x; i < Kc; i += blockDim.x) sh_hist[i] = 0;
    __syncthreads();
    for (int64_t n = threadIdx.x; n < N; n += blockDim.x) {
        uint32_t mw = mr ? mr[f * Nw + (n >> 5)] : 0xffffffffu;
        if ((mw >> (n & 31)) & 1u) {
            int k = (int)(label_hash(seed, f, n + n_offset) % (uint32_t)Kc);
            atomicAdd(&sh_hist[k], 1);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Kc; i += blockDim.x) statF[f * Kp + i] = (double)sh_hist[i];
}

// ---------------------------------------------------------------------------
// layout conversion host column-major [rows x K] <-> device row-major [rows][Kp]
// (optionally interleaved: device row = rows_mul*r + rows_add)
// ---------------------------------------------------------------------------
__global__ void colmajor_to_rows_kernel(const double *__restrict__ src, int64_t rows, int K,
                                        double *__restrict__ dst, int Kp, int mul, int add)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * K) return;
    int64_t r = i % rows;
    int k = (int)(i / rows);
    dst[(mul * r + add) * Kp + k] = src[i];
}
__global__ void rows_to_colmajor_kernel(const double *__restrict__ src, int64_t rows, int K,
                                        double *__restrict__ dst, int Kp, int mul, int add)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * K) return;
    int64_t r = i % rows;
    int k = (int)(i / rows);
    dst[i] = src[(mul * r + add) * Kp + k];
}

// ---------------------------------------------------------------------------
// parameter kernels: statistics -> operands of the contraction
//   aspect mode (VB_Aspect_Rcpp :419-444):
//     gamma_vb = gamma + n_total;  A = exp(psi(gamma_vb) - psi(sum_k gamma_vb))
//     alpha_vb = alpha + f_ones, beta_vb = beta + f_zeros
//     B1 = exp(psi(alpha_vb) - psi(alpha_vb+beta_vb)), B0 likewise with beta_vb
//   mean mode (CVB0 without the self term):
//     A = gamma_vb / sum_k gamma_vb (the row scale cancels in E_Z), B1 = alpha_vb/(alpha_vb+beta_vb), B0 = beta_vb/(...)
//   E_W = gamma_vb / sum_k gamma_vb, E_H = alpha_vb/(alpha_vb+beta_vb) (:507-514)
//   want_lb: adds pW - qW (rows) / pH - qH (columns) to scal[SC_PRIOR]
//   (E_log_p_W_Rcpp :173-182, E_log_q_W_Rcpp :185-197, E_log_p_H_Rcpp :243-253,
//    E_log_q_H_Rcpp :256-273).
// One warp per row.
// ---------------------------------------------------------------------------
__global__ void strided_copy_kernel(const double *__restrict__ src, int sstride, double *__restrict__ dst, int dstride, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[(size_t)i * dstride] = src[(size_t)i * sstride];
}

// rows [f0, F) are computed (f0 = first row of the launch); rows in [lb0, lb1) contribute to the prior
// term of the ELBO (row-sliced sharded sessions sum that term over the ranks)
__global__ void params_rows_kernel(const double *__restrict__ statF, int64_t f0, int64_t F, int K, int Kp,
                                   double gamma, int mean_mode, int want_lb, int64_t lb0, int64_t lb1,
                                   double *__restrict__ A, double *__restrict__ EW, double *scal)
{
    __shared__ double sh[32];
    int64_t f = f0 + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    int lane = threadIdx.x & 31;
    double lb = 0.0, neg = 0.0;
    if (f < F) {
        double s = 0.0;
        for (int k = lane; k < K; k += 32) s += gamma + statF[f * Kp + k];
        s = warp_sum(s);
        double ps = dev_digamma(s);
        double sum_elw = 0.0, q = 0.0;
        for (int k = lane; k < Kp; k += 32) {
            if (k < K) {
                double g = gamma + statF[f * Kp + k];
                if (g < 0) neg = 1.0;
                double elw = dev_digamma(g) - ps;
                A[f * Kp + k] = mean_mode ? g / s : exp(elw); // mean mode: E_W, so that D <= 1 in both modes
                EW[f * Kp + k] = g / s;
                if (want_lb) { sum_elw += elw; q += -lgamma(g) + (g - 1.0) * elw; }
            } else {
                A[f * Kp + k] = 0.0;
                EW[f * Kp + k] = 0.0;
            }
        }
        if (want_lb) {
            sum_elw = warp_sum(sum_elw);
            q = warp_sum(q);
            if (lane == 0 && f >= lb0 && f < lb1)
                lb = (lgamma(K * gamma) - K * lgamma(gamma)) + (gamma - 1.0) * sum_elw -
                     (lgamma(s) + q);
        }
    }
    if (want_lb) {
        double t = block_sum(lb, sh);
        if (threadIdx.x == 0 && t != 0.0) atomicAdd(&scal[SC_PRIOR_ROWS], t);
    }
    if (neg != 0.0) scal[SC_NEG] = 1.0;
}

__global__ void params_cols_kernel(const double *__restrict__ statN, int64_t N, int K, int Kp,
                                   double alpha, double beta, int mean_mode, int want_lb,
                                   double *__restrict__ BB, double *__restrict__ EH, double *scal)
{
    __shared__ double sh[32];
    int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    double lb = 0.0, neg = 0.0;
    if (n < N) {
        for (int k = lane; k < Kp; k += 32) {
            if (k < K) {
                double a = alpha + statN[(2 * n + 1) * Kp + k];
                double b = beta + statN[(2 * n) * Kp + k];
                if (a < 0 || b < 0) neg = 1.0;
                double pab = dev_digamma(a + b);
                double elh = dev_digamma(a) - pab, el1h = dev_digamma(b) - pab;
                BB[(2 * n + 1) * Kp + k] = mean_mode ? a / (a + b) : exp(elh);
                BB[(2 * n) * Kp + k] = mean_mode ? b / (a + b) : exp(el1h);
                EH[n * Kp + k] = a / (a + b);
                if (want_lb)
                    lb += (lgamma(alpha + beta) - lgamma(alpha) - lgamma(beta)) +
                          (alpha - 1.0) * elh + (beta - 1.0) * el1h -
                          (lgamma(a + b) - lgamma(a) - lgamma(b) + (a - 1.0) * elh +
                           (b - 1.0) * el1h);
            } else {
                BB[(2 * n + 1) * Kp + k] = 0.0;
                BB[(2 * n) * Kp + k] = 0.0;
                EH[n * Kp + k] = 0.0;
            }
        }
    }
    if (want_lb) {
        double t = block_sum(lb, sh);
        if (threadIdx.x == 0 && t != 0.0) atomicAdd(&scal[SC_PRIOR_COLS], t);
    }
    if (neg != 0.0) scal[SC_NEG] = 1.0;
}

// stat = U (.) G   (n_total = A (.) (R1 B1 + R0 B0), f_ones = B1 (.) (R1^T A), ...)
__global__ void finalize_stats_kernel(const double *__restrict__ U, const double *__restrict__ G,
                                      int64_t count, double *__restrict__ stat)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) stat[i] = U[i] * G[i];
}

// ---------------------------------------------------------------------------
// FP64 contraction pass ("owner / stream" form shared by both statistics).
//   owner lane l  (L lanes):  operand u_l = U[l][:]      accumulator G[l][:]
//   streamed row  c:          operand w_c = W[c][:]
//   S_lc = u_l . w_c ;  R_lc = sel(l,c) / S_lc ;  G[l][:] += sum_c R_lc w_c
//   sum log D accumulates log S_lc over selected pairs.
// rows pass  (Dirichlet statistic): owner = row f (ob=1), U = A, streamed =
//   (n,b) pairs (sb=2), W = BB, bits packed along n; sel = mask & (v == b).
// cols pass  (Beta statistics): owner lane = 2n+b (ob=2), U = BB, streamed =
//   row f (sb=1), W = A, bits packed along f; sel = mask & (v == b).
// This is VB_Aspect_Rcpp's q(Z) loop (collapsed_gibbs_z_VB.cpp:447-465) with
// the E_Z cube eliminated.
// ---------------------------------------------------------------------------
struct PassArgs {
    const double *U, *W;
    double *G;
    const uint32_t *vbits, *mbits;
    int64_t words, L, S, s_per_split;
    int ob, sb, Kp, TR;
    double *sumlogD;
};

// SIMT form (FP64 or FP32 FMA pipe + warp shuffles): NKC = Kp/KC consecutive lanes share an owner
// lane and split k, the dot product is finished with shfl.xor.  T = double: any Kp up to 512 (the
// DMMA kernel of pass_mma64.cuh takes Kp <= 128); T = float: the small-K FFMA path of the north
// star (NBMF_PREC_FP32) -- operands and accumulation in fp32, results handed back in fp64.
template <typename T, int KC>
__global__ void __launch_bounds__(256) pass_simt_kernel(PassArgs a)
{
    extern __shared__ unsigned char sm_w_raw[];
    T *sm_w = reinterpret_cast<T *>(sm_w_raw); // [TR*sb][Kp]
    __shared__ double sh_red[32];
    const int Kp = a.Kp, NKC = Kp / KC, tid = threadIdx.x;
    const int kc = tid % NKC, ol = tid / NKC, OL = 256 / NKC;
    const int64_t l = (int64_t)blockIdx.x * OL + ol;
    const bool lane_ok = l < a.L;
    const int64_t orow = lane_ok ? l / a.ob : 0;
    const int bown = (int)(l % a.ob);
    T u[KC], acc[KC];
#pragma unroll
    for (int i = 0; i < KC; i++) {
        u[i] = lane_ok ? (T)a.U[l * Kp + kc * KC + i] : (T)0;
        acc[i] = (T)0;
    }
    double ell = 0.0;
    const int64_t s_begin = (int64_t)blockIdx.y * a.s_per_split;
    const int64_t s_end = min(a.S, s_begin + a.s_per_split);
    const int TR = a.TR, sb = a.sb;
    const int tile_elems = TR * sb * Kp;
    for (int64_t r0 = s_begin; r0 < s_end; r0 += TR) {
        __syncthreads();
        const double *src = a.W + r0 * sb * Kp;
        const int64_t valid = min((int64_t)tile_elems, (a.S - r0) * sb * Kp);
        for (int i = tid; i < tile_elems; i += 256) sm_w[i] = (i < valid) ? (T)src[i] : (T)0;
        __syncthreads();
        uint32_t wv = 0, wm = 0;
        if (lane_ok) {
            wv = a.vbits[orow * a.words + (r0 >> 5)];
            wm = a.mbits ? a.mbits[orow * a.words + (r0 >> 5)] : 0xffffffffu;
        }
        const int sh0 = (int)(r0 & 31);
        for (int j = 0; j < TR; j++) {
            const int bit = sh0 + j;
            const int v = (wv >> bit) & 1u;
            bool act = ((wm >> bit) & 1u) && (r0 + j < s_end);
            int c;
            if (sb == 2) c = 2 * j + v;
            else { c = j; act = act && (v == bown); }
            const T *wrow = sm_w + c * Kp + kc * KC;
            T d = (T)0;
#pragma unroll
            for (int i = 0; i < KC; i++) d = fma(u[i], wrow[i], d);
            for (int o = NKC >> 1; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
            const T R = act ? (T)1 / d : (T)0;
            if (a.sumlogD != nullptr && act && kc == 0) ell += (double)log(d);
#pragma unroll
            for (int i = 0; i < KC; i++) acc[i] = fma(R, wrow[i], acc[i]);
        }
    }
    if (lane_ok) {
        if (gridDim.y == 1) {
#pragma unroll
            for (int i = 0; i < KC; i++) a.G[l * Kp + kc * KC + i] = (double)acc[i];
        } else {
#pragma unroll
            for (int i = 0; i < KC; i++) atomicAdd(&a.G[l * Kp + kc * KC + i], (double)acc[i]);
        }
    }
    if (a.sumlogD != nullptr) {
        double t = block_sum(ell, sh_red);
        if (tid == 0) atomicAdd(a.sumlogD, t);
    }
}

// E_Z(f,k,n) = A_fk B^{v}_nk / D_fn for observed entries, 0 otherwise
// (collapsed_gibbs_z_VB.cpp:453-462), arma::cube layout f + F k + F K n.
__global__ void ez_aspect_kernel(const double *__restrict__ A, const double *__restrict__ BB,
                                 const uint32_t *__restrict__ vc, const uint32_t *__restrict__ mc,
                                 int64_t F, int64_t N, int64_t Fw, int K, int Kp,
                                 double *__restrict__ EZ)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * N) return;
    int64_t f = i % F, n = i / F;
    uint32_t mw = mc ? mc[n * Fw + (f >> 5)] : 0xffffffffu;
    bool obs = (mw >> (f & 31)) & 1u;
    int v = (vc[n * Fw + (f >> 5)] >> (f & 31)) & 1u;
    const double *a = A + f * Kp, *b = BB + (2 * n + v) * Kp;
    double d = 0.0;
    for (int k = 0; k < K; k++) d += a[k] * b[k];
    for (int k = 0; k < K; k++) EZ[f + F * k + F * (int64_t)K * n] = obs ? a[k] * b[k] / d : 0.0;
}

// Z_fn <- which.max(E_Z[f,,n]) (BernoulliNMF.R:45-52, the Gibbs warm start) without the
// E_Z cube: argmax_k A_fk B^{v}_nk from the operands of the last VB iteration; first maximum
// wins, as which.max does.  NA where V is NA.
__global__ void argmax_labels_kernel(const double *__restrict__ A, const double *__restrict__ BB,
                                     const uint32_t *__restrict__ vc, const uint32_t *__restrict__ mc,
                                     int64_t F, int64_t N, int64_t Fw, int K, int Kp, int32_t *__restrict__ Z)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * N) return;
    int64_t f = i % F, n = i / F;
    uint32_t mw = mc ? mc[n * Fw + (f >> 5)] : 0xffffffffu;
    if (!((mw >> (f & 31)) & 1u)) { Z[i] = NBMF_NA_INTEGER; return; }
    int v = (vc[n * Fw + (f >> 5)] >> (f & 31)) & 1u;
    const double *a = A + f * Kp, *b = BB + (2 * n + v) * Kp;
    int best = 0;
    double bv = a[0] * b[0];
    for (int k = 1; k < K; k++) {
        double p = a[k] * b[k];
        if (p > bv) { bv = p; best = k; }
    }
    Z[i] = best;
}

__global__ void f64_stats_to_counts_kernel(const double *__restrict__ statF, const double *__restrict__ statN,
                                           int64_t F, int64_t N, int K, int Kp, int32_t *__restrict__ cF,
                                           int32_t *__restrict__ cN)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < F * K) cF[i] = (int32_t)llrint(statF[(i / K) * Kp + (i % K)]);
    if (i < 2 * N * K) cN[i] = (int32_t)llrint(statN[(i / K) * Kp + (i % K)]);
}
__global__ void i32_to_f64_kernel(const int32_t *__restrict__ a, int64_t n, double *__restrict__ b)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = (double)a[i];
}
__global__ void f64_to_i32_kernel(const double *__restrict__ a, int64_t n, int32_t *__restrict__ b)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = (int32_t)llrint(a[i]);
}

// ---------------------------------------------------------------------------
// VB_DirBer_Rcpp sweep (collapsed_gibbs_z_VB.cpp:119-141) on the anti-diagonal
// wavefront d = f + n: entries of one diagonal touch disjoint counter rows, and
// ordering by d preserves the order of every conflicting pair of the
// reference's column-major sweep, so the result is bit-identical (SURVEY.md
// Appendix A).  Arithmetic is written with explicit _rn intrinsics so that no
// FMA contraction changes a rounding.  E_Z is [n][f][Kc] on the device.
// One thread per entry; a grid barrier (or __syncthreads for one block)
// separates diagonals.
// ---------------------------------------------------------------------------
struct WaveArgs {
    const uint32_t *vc, *mc;
    int64_t F, N, Fw;
    int Kc, sweeps;
    double alpha, beta, gamma;
    double *EZ;     // [N][F][Kc]
    double *cF;     // [F][Kc] n_total
    double *cN1;    // [N][Kc] f_ones
    double *cN0;    // [N][Kc] f_zeros
    double *gvb;    // [F][Kc] gamma_vb (stale rows)
    double *avb;    // [N][Kc]
    double *bvb;    // [N][Kc]
};

// G lanes (a power of two <= 32) share one entry and split k; everything that the reference does
// independently per k (counter updates, the divisions) runs in parallel, and only the two
// order-sensitive reductions stay sequential: the running sum over k is carried through the
// group with shuffles in the reference's order k = 0, 1, ..., so the result is bit-identical to
// the one-thread-per-entry form while the critical path per entry drops from 2 Kc divisions to
// two divisions and Kc additions.
template <bool SINGLE_BLOCK, int G>
__global__ void __launch_bounds__(1024) cvb0_wavefront_kernel(WaveArgs a)
{
    cg::grid_group grid = cg::this_grid();
    const int64_t F = a.F, N = a.N;
    const int Kc = a.Kc;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / G;
    const int64_t gid = tid / G;
    const int gl = (int)(tid & (G - 1));                 // lane within the group
    const int gbase = (threadIdx.x & 31) & ~(G - 1);     // first warp lane of the group
    for (int sweep = 0; sweep < a.sweeps; sweep++) {
        for (int64_t d = 0; d <= F + N - 2; d++) {
            const int64_t f_lo = d - (N - 1) > 0 ? d - (N - 1) : 0;
            const int64_t f_hi = d < F - 1 ? d : F - 1;
            const int64_t cnt = f_hi - f_lo + 1;
            // uniform trip count: the shuffles below need every lane of the warp
            for (int64_t base = 0; base < cnt; base += ngroups) {
                const int64_t f = f_lo + base + gid;
                const int64_t n = d - f;
                bool act = (base + gid) < cnt;
                int v = 0;
                if (act) {
                    const uint32_t mw = a.mc ? a.mc[n * a.Fw + (f >> 5)] : 0xffffffffu;
                    act = (mw >> (f & 31)) & 1u;
                    v = (a.vc[n * a.Fw + (f >> 5)] >> (f & 31)) & 1u;
                }
                const int64_t fs = act ? f : 0, ns = act ? n : 0;    // inactive groups compute on row 0, store nothing
                double *ez = a.EZ + (ns * F + fs) * Kc;
                double *nt = a.cF + fs * Kc, *fo = a.cN1 + ns * Kc, *fz = a.cN0 + ns * Kc;
                double *gv = a.gvb + fs * Kc, *av = a.avb + ns * Kc, *bv = a.bvb + ns * Kc;
                double s = 0.0;
                for (int k0 = 0; k0 < Kc; k0 += G) {
                    const int k = k0 + gl;
                    double pr = 0.0;
                    if (k < Kc) {
                        const double p = ez[k];
                        // remove_vbassignment (:81-92)
                        const double o1 = __dsub_rn(fo[k], v ? p : 0.0);
                        const double o0 = __dsub_rn(fz[k], v ? 0.0 : p);
                        const double t = __dsub_rn(nt[k], p);
                        // :128-133
                        const double g = __dadd_rn(a.gamma, t);
                        const double al = __dadd_rn(a.alpha, o1);
                        const double be = __dadd_rn(a.beta, o0);
                        pr = __ddiv_rn(__dmul_rn(g, v ? al : be), __dadd_rn(al, be));
                        if (act) { fo[k] = o1; fz[k] = o0; nt[k] = t; gv[k] = g; av[k] = al; bv[k] = be; ez[k] = pr; }
                    }
                    const int kn = Kc - k0 < G ? Kc - k0 : G;
                    for (int j = 0; j < kn; j++) s = __dadd_rn(s, __shfl_sync(0xffffffffu, pr, gbase + j));
                }
                if (act) {
                    for (int k = gl; k < Kc; k += G) {
                        const double pr = __ddiv_rn(ez[k], s); // :134
                        ez[k] = pr;                            // set_vbassignment (:69-78)
                        fo[k] = __dadd_rn(fo[k], v ? pr : 0.0);
                        fz[k] = __dadd_rn(fz[k], v ? 0.0 : pr);
                        nt[k] = __dadd_rn(nt[k], pr);
                    }
                }
            }
            if (SINGLE_BLOCK) __syncthreads();
            else grid.sync();
        }
    }
}

// ---------------------------------------------------------------------------
// Collapsed Gibbs sweep of Gibbs_DirBer_finite_Rcpp (collapsed_gibbs_z.cpp:392-421) in the
// same anti-diagonal wavefront order as the CVB0 kernel above: an entry reads and writes only
// the count rows of its own f and n, so ordering by d = f + n keeps every conflicting pair in
// the reference's column-major order.  The reference draws from R's global stream
// (rmultinom, :411); here the uniform of entry (f, n) in sweep i is Philox4x32-10 keyed
// (seed; f, n, i), which makes the chain independent of the traversal order -- the oracle
// restates the reference loop with the same key, and the labels agree bit for bit.
// Integer counts; the probabilities follow :402-408 in fp64 without FMA contraction.
// ---------------------------------------------------------------------------
struct GibbsWaveArgs {
    const uint32_t *vc, *mc;
    int64_t F, N, Fw;
    int K, iter, nburnin, nsamples;
    double alpha, beta, gamma;
    uint64_t seed;
    int32_t *Z;          // [N][F] = f + F n, updated in place
    int32_t *cF;         // [F][K]      n_total
    int32_t *cN;         // [2N][K]     row 2n+1 f_ones, row 2n f_zeros
    int32_t *Z_samples;  // [nsamples][N][F] or NULL (:423-426)
    int *n_written;
    // Dir-Dir only: the second assignment C with its own counts
    int32_t *C, *cF2, *cN2, *C_samples;
};

// Categorical draw shared by the G lanes of a group: q(k) is the unnormalised weight of
// component k.  s = sum_k q(k) and the inverse-CDF scan over q(k)/s are accumulated in the
// order k = 0, 1, ... (shuffles carry the running sums), exactly as probs/accu(probs) followed
// by a sequential scan would; every lane returns the same k.
template <int G, typename QF>
__device__ __forceinline__ int group_draw(int K, int gl, int gbase, double u, QF q)
{
    double s = 0.0;
    for (int k0 = 0; k0 < K; k0 += G) {
        const int k = k0 + gl;
        const double qk = k < K ? q(k) : 0.0;
        const int kn = K - k0 < G ? K - k0 : G;
        for (int j = 0; j < kn; j++) s = __dadd_rn(s, __shfl_sync(0xffffffffu, qk, gbase + j));
    }
    double c = 0.0;
    int k1 = -1, last_pos = K - 1;
    for (int k0 = 0; k0 < K; k0 += G) {
        const int k = k0 + gl;
        const double pr = k < K ? __ddiv_rn(q(k), s) : 0.0;
        const int kn = K - k0 < G ? K - k0 : G;
        for (int j = 0; j < kn; j++) {
            const double pj = __shfl_sync(0xffffffffu, pr, gbase + j);
            c = __dadd_rn(c, pj);
            if (pj > 0.0) last_pos = k0 + j;
            if (k1 < 0 && u < c) k1 = k0 + j;
        }
    }
    return k1 < 0 ? last_pos : k1;
}

// VARIANT 0: Gibbs_DirBer_finite_Rcpp (collapsed_gibbs_z.cpp:363-437)
//         1: Gibbs_DirBer_DP_Rcpp     (:275-359; K = 200, V without NA)
//         2: Gibbs_DirDir_finite_Rcpp (:441-557; assignments Z and C)
template <bool SINGLE_BLOCK, int G, int VARIANT>
__global__ void __launch_bounds__(1024) gibbs_collapsed_wavefront_kernel(GibbsWaveArgs a)
{
    cg::grid_group grid = cg::this_grid();
    const int64_t F = a.F, N = a.N;
    const int K = a.K;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t ngroups = nthreads / G;
    const int64_t gid = tid / G;
    const int gl = (int)(tid & (G - 1));
    const int gbase = (threadIdx.x & 31) & ~(G - 1);
    int jw = 0;
    for (int i = 1; i < a.iter; i++) {   // :392
        for (int64_t d = 0; d <= F + N - 2; d++) {
            const int64_t f_lo = d - (N - 1) > 0 ? d - (N - 1) : 0;
            const int64_t f_hi = d < F - 1 ? d : F - 1;
            const int64_t cnt = f_hi - f_lo + 1;
            for (int64_t base = 0; base < cnt; base += ngroups) {   // uniform trip count (shuffles)
                const int64_t f = f_lo + base + gid;
                const int64_t n = d - f;
                bool act = (base + gid) < cnt;
                int v = 0;
                if (act) {
                    const uint32_t mw = a.mc ? a.mc[n * a.Fw + (f >> 5)] : 0xffffffffu;
                    act = (mw >> (f & 31)) & 1u;
                    v = (a.vc[n * a.Fw + (f >> 5)] >> (f & 31)) & 1u;
                }
                // inactive groups run the arithmetic on row 0 and store nothing
                const int64_t fs = act ? f : 0, ns = act ? n : 0;
                int32_t *nF = a.cF + fs * K, *nV = a.cN + (2 * ns + v) * K;
                const int32_t *n1 = a.cN + (2 * ns + 1) * K, *n0 = a.cN + (2 * ns) * K;
                const int k_old = act ? a.Z[fs + F * ns] : -1;
                const philox4 r = philox4x32_10((uint32_t)fs, (uint32_t)ns, (uint32_t)i, 0u, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
                const double u0 = u64_to_unif(r.x, r.y);
                // The counts in memory are touched once, at the end, by the group's first lane; until
                // then "the entry's own assignment removed" (:396) is a -1 on the component it holds.
                if (VARIANT == 0) {
                    const int k1 = group_draw<G>(K, gl, gbase, u0, [&](int k) {   // :402-408
                        const int own = (k == k_old) ? 1 : 0;
                        const double gp = __dadd_rn(a.gamma, (double)(nF[k] - own));
                        const double ap = __dadd_rn(a.alpha, (double)(n1[k] - (v ? own : 0)));
                        const double bp = __dadd_rn(a.beta, (double)(n0[k] - (v ? 0 : own)));
                        return __ddiv_rn(__dmul_rn(gp, v ? ap : bp), __dadd_rn(ap, bp));
                    });
                    if (act && gl == 0) {
                        nF[k_old] -= 1; nV[k_old] -= 1;   // remove_assignment (:75-85)
                        a.Z[fs + F * ns] = k1;            // set_assignment (:63-72)
                        nF[k1] += 1; nV[k1] += 1;
                    }
                } else if (VARIANT == 1) {
                    // first empty component of row f (:325-328), found chunk by chunk with a ballot
                    int k_free = -1;
                    for (int k0 = 0; k0 < K; k0 += G) {   // every lane of the warp takes every ballot
                        const int k = k0 + gl;
                        const bool empty = k < K && (nF[k] - (k == k_old ? 1 : 0)) == 0;
                        const uint32_t m = (__ballot_sync(0xffffffffu, empty) >> gbase) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u));
                        if (m && k_free < 0) k_free = k0 + __ffs(m) - 1;
                    }
                    const int k1 = group_draw<G>(K, gl, gbase, u0, [&](int k) {   // :317-329
                        if (k == k_free) return __dmul_rn(a.gamma, 0.5);
                        const int own = (k == k_old) ? 1 : 0;
                        const double ap = __dadd_rn((double)(n1[k] - (v ? own : 0)), 1e-8);
                        const double bp = __dadd_rn((double)(n0[k] - (v ? 0 : own)), 1e-8);
                        return __dmul_rn((double)(nF[k] - own), __ddiv_rn(v ? ap : bp, __dadd_rn(ap, bp)));
                    });
                    if (act && gl == 0) {
                        nF[k_old] -= 1; nV[k_old] -= 1;
                        a.Z[fs + F * ns] = k1;
                        nF[k1] += 1; nV[k1] += 1;
                    }
                } else {
                    int32_t *cF2 = a.cF2 + fs * K, *cV2 = a.cN2 + (2 * ns + v) * K;
                    const int32_t *c1 = a.cN2 + (2 * ns + 1) * K, *c0 = a.cN2 + (2 * ns) * K;
                    const int c_old = act ? a.C[fs + F * ns] : -1;
                    const double u1 = u64_to_unif(r.z, r.w);
                    // Both draws are executed by every group (the shuffles need the whole warp); which
                    // formula a group uses, and whether it keeps the second draw, depends on its v.
                    // V = 1: Z and C together from (gamma + n^Z_f)(alpha + t^C_n) (:481-497).
                    // V = 0: Z with C's component forbidden, then C with the new Z component forbidden (:500-531).
                    const int kz = group_draw<G>(K, gl, gbase, u0, [&](int k) {
                        const double gp = __dadd_rn(a.gamma, (double)(nF[k] - (k == k_old ? 1 : 0)));
                        if (v) return __dmul_rn(gp, __dadd_rn(a.alpha, (double)(c1[k] + c0[k] - (k == c_old ? 1 : 0))));
                        return k == c_old ? 0.0 : gp;
                    });
                    const int kc2 = group_draw<G>(K, gl, gbase, u1, [&](int k) {
                        return k == kz ? 0.0 : __dadd_rn(a.alpha, (double)(c1[k] + c0[k] - (k == c_old ? 1 : 0)));
                    });
                    const int kc = v ? kz : kc2;
                    if (act && gl == 0) {
                        nF[k_old] -= 1; nV[k_old] -= 1; cF2[c_old] -= 1; cV2[c_old] -= 1;
                        a.Z[fs + F * ns] = kz; a.C[fs + F * ns] = kc;
                        nF[kz] += 1; nV[kz] += 1; cF2[kc] += 1; cV2[kc] += 1;
                    }
                }
            }
            if (SINGLE_BLOCK) __syncthreads();
            else grid.sync();
        }
        if (i >= a.nburnin && jw < a.nsamples) {   // :423-426
            if (a.Z_samples) {
                int32_t *dst = a.Z_samples + (int64_t)jw * F * N;
                for (int64_t e = tid; e < F * N; e += nthreads) dst[e] = a.Z[e];
                if (VARIANT == 2 && a.C_samples) {
                    int32_t *dc = a.C_samples + (int64_t)jw * F * N;
                    for (int64_t e = tid; e < F * N; e += nthreads) dc[e] = a.C[e];
                }
                if (SINGLE_BLOCK) __syncthreads();
                else grid.sync();
            }
            jw++;
        }
    }
    if (tid == 0) *a.n_written = jw;
}

// one-hot E_Z [n][f][Kc] + counters from labels (commons.cpp:29-42,
// collapsed_gibbs_z_VB.cpp:42-66); counters come from the histogram kernel.
__global__ void onehot_ez_kernel(const int32_t *__restrict__ Z, const uint32_t *__restrict__ mc,
                                 int64_t F, int64_t N, int64_t Fw, int Kc, double *EZ)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * N) return;
    int64_t f = i % F, n = i / F;
    uint32_t mw = mc ? mc[n * Fw + (f >> 5)] : 0xffffffffu;
    bool obs = (mw >> (f & 31)) & 1u;
    int32_t z = Z[i];
    double *ez = EZ + i * Kc; // (n*F + f)*Kc
    for (int k = 0; k < Kc; k++) ez[k] = (obs && z == k) ? 1.0 : 0.0;
}

// [n][f][Kc] -> arma::cube layout f + F k + F Kc n
__global__ void ez_to_cube_kernel(const double *__restrict__ EZ, int64_t F, int64_t N, int Kc,
                                  double *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * N * Kc) return;
    int64_t f = i % F;
    int k = (int)((i / F) % Kc);
    int64_t n = i / (F * Kc);
    out[i] = EZ[(n * F + f) * Kc + k];
}

// stale-row outputs of VB_DirBer_Rcpp (:151-159): E_W = gamma_vb / rowsum,
// E_H = alpha_vb / (alpha_vb + beta_vb), straight to column-major.
__global__ void dirber_outputs_kernel(const double *__restrict__ gvb, const double *__restrict__ avb,
                                      const double *__restrict__ bvb, int64_t F, int64_t N, int Kc,
                                      double *__restrict__ EW, double *__restrict__ EH, double *scal)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < F) {
        double s = 0.0;
        for (int k = 0; k < Kc; k++) s += gvb[i * Kc + k];
        for (int k = 0; k < Kc; k++) EW[i + F * k] = gvb[i * Kc + k] / s;
    }
    if (i < N) {
        for (int k = 0; k < Kc; k++) {
            double al = avb[i * Kc + k], be = bvb[i * Kc + k];
            if (al < 0 || be < 0) scal[SC_NEG] = 1.0;
            EH[i + N * k] = al / (al + be);
        }
    }
}

// ---------------------------------------------------------------------------
// Gibbs: uncollapsed blocked sampler  Z|W,H -> counts -> W|Z, H|Z,V
// (collapsed_gibbs_z.cpp:589-605 conditionals; R/MCEM - sample_gibbs.R:40-77
//  sweep; sample_gibbs.cpp:43 likelihood form).
// ---------------------------------------------------------------------------
struct PhiloxStream {
    uint32_t c0, c1, c2, k0, k1, ctr;
    philox4 buf; int have;
    __device__ PhiloxStream(uint64_t seed, uint32_t a, uint32_t b, uint32_t c)
        : c0(a), c1(b), c2(c), k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), ctr(0), have(0) {}
    __device__ uint32_t next()
    {
        if (have == 0) { buf = philox4x32_10(c0, c1, c2, ctr++, k0, k1); have = 4; }
        uint32_t r = have == 4 ? buf.x : have == 3 ? buf.y : have == 2 ? buf.z : buf.w;
        have--;
        return r;
    }
    __device__ double unif() { uint32_t a = next(), b = next(); return u64_to_unif(a, b); }
    __device__ double normal()
    {
        double u = unif(), v = unif();
        return sqrt(-2.0 * log(u)) * cospi(2.0 * v);
    }
    // Marsaglia-Tsang with the U^(1/a) boost for shape < 1 (gamma = 1/K priors)
    __device__ double gamma(double a)
    {
        double boost = 1.0;
        if (a < 1.0) { boost = pow(unif(), 1.0 / a); a += 1.0; }
        const double d = a - 1.0 / 3.0, c = rsqrt(9.0 * d);
        for (int it = 0; it < 64; it++) {
            double x = normal(), v = 1.0 + c * x;
            if (v <= 0.0) continue;
            v = v * v * v;
            double u = unif();
            if (u < 1.0 - 0.0331 * x * x * x * x) return d * v * boost;
            if (log(u) < 0.5 * x * x + d * (1.0 - v + log(v))) return d * v * boost;
        }
        return d * boost;
    }
};

// W_f ~ Dir(gamma + n_total_f): one warp per row, normalised Gammas (:590-596).
// Every rank draws the same W from the shared key (seed, sweep) -- no broadcast.
__global__ void gibbs_draw_W_kernel(const int32_t *__restrict__ cF, int64_t F, int K, double gamma,
                                    uint64_t seed, uint32_t sweep, float *__restrict__ W)
{
    int64_t f = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (f >= F) return;
    double s = 0.0;
    for (int k0 = 0; k0 < K; k0 += 32) {
        int k = k0 + lane;
        double g = 0.0;
        if (k < K) {
            PhiloxStream rs(seed, (uint32_t)f, (uint32_t)k | 0x40000000u, sweep);
            g = rs.gamma(gamma + (double)cF[f * K + k]);
        }
        s += g;
        if (k < K) W[f * K + k] = (float)g; // overwritten below after normalisation
    }
    s = warp_sum(s);
    for (int k = lane; k < K; k += 32) {
        double g = (double)W[f * K + k];
        W[f * K + k] = (s > 0.0) ? (float)(g / s) : 1.0f / K;
    }
}

// H_nk ~ Beta(alpha + ones, beta + zeros) (:599-605), keyed by the global column
__global__ void gibbs_draw_H_kernel(const int32_t *__restrict__ cN, int64_t N, int64_t n_offset,
                                    int K, double alpha, double beta, uint64_t seed, uint32_t sweep,
                                    float *__restrict__ H)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * K) return;
    int64_t n = i / K;
    int k = (int)(i % K);
    int64_t ng = n + n_offset;
    PhiloxStream rs(seed, (uint32_t)ng, (uint32_t)k | 0x80000000u | ((uint32_t)(ng >> 32) << 16),
                    sweep);
    double x = rs.gamma(alpha + (double)cN[(2 * n + 1) * K + k]);
    double y = rs.gamma(beta + (double)cN[(2 * n) * K + k]);
    double h = x / (x + y);
    if (!(h > 0.0)) h = 1e-300;
    H[n * K + k] = (float)h;
}

// Z_fn ~ Cat(W_fk H_nk^v (1-H_nk)^(1-v)) for a 32(f) x 32(n) tile per block of
// 256 threads; counts reduced in shared memory, flushed with integer atomics.
// Counter-based draw: Philox(seed; f, global n, sweep).
__global__ void __launch_bounds__(256) gibbs_sweep_kernel(
    const uint32_t *__restrict__ vc, const uint32_t *__restrict__ mc, int64_t F, int64_t N,
    int64_t Fw, int64_t n_offset, int K, const float *__restrict__ W, const float *__restrict__ H,
    uint64_t seed, uint32_t sweep, int32_t *__restrict__ cF, int32_t *__restrict__ cN,
    int32_t *__restrict__ Zout)
{
    extern __shared__ unsigned char sm_raw[];
    const int KS = K + 1; // padded stride: lanes index different f rows
    float *sW = (float *)sm_raw;            // [32][KS]
    float *sH = sW + 32 * KS;               // [32][K]
    int *hF = (int *)(sH + 32 * K);         // [32][K]
    int *hN = hF + 32 * K;                  // [64][K]
    const int64_t f0 = (int64_t)blockIdx.x * 32, n0 = (int64_t)blockIdx.y * 32;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    for (int i = tid; i < 32 * K; i += 256) {
        int r = i / K, k = i % K;
        sW[r * KS + k] = (f0 + r < F) ? W[(f0 + r) * K + k] : 0.f;
        sH[i] = (n0 + r < N) ? H[(n0 + r) * K + k] : 0.f;
        hF[i] = 0; hN[i] = 0; hN[i + 32 * K] = 0;
    }
    __syncthreads();
    const int64_t f = f0 + tx;
    for (int jn = ty; jn < 32; jn += 8) {
        const int64_t n = n0 + jn;
        if (n >= N) break;
        uint32_t mw = mc ? mc[n * Fw + (f0 >> 5)] : 0xffffffffu;
        uint32_t vw = vc[n * Fw + (f0 >> 5)];
        bool obs = (f < F) && ((mw >> tx) & 1u);
        int v = (vw >> tx) & 1u;
        int kz = -1;
        if (obs) {
            const float *w = sW + tx * KS, *hh = sH + jn * K;
            float tot = 0.f;
            for (int k = 0; k < K; k++) tot += w[k] * (v ? hh[k] : 1.f - hh[k]);
            philox4 r = philox4x32_10((uint32_t)f, (uint32_t)(n + n_offset),
                                      (uint32_t)((n + n_offset) >> 32), sweep, (uint32_t)seed,
                                      (uint32_t)(seed >> 32) ^ 0x2545F491u);
            float target = u32_to_unif(r.x) * tot, c = 0.f;
            kz = K - 1;
            for (int k = 0; k < K; k++) {
                c += w[k] * (v ? hh[k] : 1.f - hh[k]);
                if (target < c) { kz = k; break; }
            }
            atomicAdd(&hF[tx * K + kz], 1);
            atomicAdd(&hN[(2 * jn + v) * K + kz], 1);
        }
        if (Zout && f < F) Zout[f + F * n] = obs ? kz : NBMF_NA_INTEGER;
    }
    __syncthreads();
    for (int i = tid; i < 32 * K; i += 256) {
        int r = i / K;
        if (hF[i] && f0 + r < F) atomicAdd(&cF[(f0 + r) * K + (i % K)], hF[i]);
    }
    for (int i = tid; i < 64 * K; i += 256) {
        int r = i / K;
        if (hN[i] && n0 + (r >> 1) < N) atomicAdd(&cN[(2 * n0 + r) * K + (i % K)], hN[i]);
    }
}

// integer counts from stored labels (compute_counts, collapsed_gibbs_z.cpp:38-60)
__global__ void counts_from_labels_kernel(const int32_t *__restrict__ Z,
                                          const uint32_t *__restrict__ vc,
                                          const uint32_t *__restrict__ mc, int64_t F, int64_t N,
                                          int64_t Fw, int K, int32_t *cF, int32_t *cN,
                                          unsigned long long *bad)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * N) return;
    int64_t f = i % F, n = i / F;
    uint32_t mw = mc ? mc[n * Fw + (f >> 5)] : 0xffffffffu;
    if (!((mw >> (f & 31)) & 1u)) return;
    int32_t k = Z[i];
    if (k < 0 || k >= K) { atomicAdd(bad, 1ull); return; }
    int v = (vc[n * Fw + (f >> 5)] >> (f & 31)) & 1u;
    atomicAdd(&cF[f * K + k], 1);
    atomicAdd(&cN[(2 * n + v) * K + k], 1);
}

// Rao-Blackwell accumulators of expectation.GibbsDirBer
// (R/generic_Gibbs_DirBer.R:82-90): w = (gamma + n_total)/sum, h = (alpha +
// ones)/(alpha + beta + total); also kept as the sweep's w, h for E_V.
__global__ void gibbs_accumulate_kernel(const int32_t *__restrict__ cF,
                                        const int32_t *__restrict__ cN, int64_t F, int64_t N, int K,
                                        double alpha, double beta, double gamma,
                                        double *__restrict__ accW, double *__restrict__ accH,
                                        float *__restrict__ w_cur, float *__restrict__ h_cur)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < F) {
        double s = 0.0;
        for (int k = 0; k < K; k++) s += gamma + cF[i * K + k];
        for (int k = 0; k < K; k++) {
            double w = (gamma + cF[i * K + k]) / s;
            accW[i * K + k] += w;
            w_cur[i * K + k] = (float)w;
        }
    }
    if (i < N) {
        for (int k = 0; k < K; k++) {
            double o = cN[(2 * i + 1) * K + k], z = cN[(2 * i) * K + k];
            double h = (alpha + o) / (alpha + beta + o + z);
            accH[i * K + k] += h;
            h_cur[i * K + k] = (float)h;
        }
    }
}

// E_V on held-out entries only: acc[t] += sum_k w_fk h_nk (R:93, test positions)
__global__ void gibbs_ev_test_kernel(const int64_t *__restrict__ idx, int64_t n_test, int64_t F,
                                     int K, const float *__restrict__ w, const float *__restrict__ h,
                                     double *__restrict__ acc)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_test) return;
    int64_t f = idx[t] % F, n = idx[t] / F;
    double s = 0.0;
    for (int k = 0; k < K; k++) s += (double)w[f * K + k] * (double)h[n * K + k];
    acc[t] += s;
}

__global__ void scale_rows_to_colmajor_kernel(const double *__restrict__ src, int64_t rows, int K,
                                              double scale, double *__restrict__ dst)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * K) return;
    int64_t r = i % rows;
    int k = (int)(i / rows);
    dst[i] = src[r * K + k] * scale;
}

// per-column sweep given a fixed dictionary (sample_gibbs.cpp:22-80 /
// R sample_gibbs_ZH R/MCEM - sample_gibbs.R:40-77), single block.
//   lik_fk = v_f W_fk + (1-v_f)(1-W_fk);  h ~ Dir(alpha + colSums C);
//   C_f ~ Cat(h (.) lik_f);  C_mean accumulates the one-hot draws.
__global__ void gibbs_given_W_kernel(const int32_t *__restrict__ v, const double *__restrict__ W,
                                     int64_t F, int K, double alpha, int iter, int nburnin,
                                     uint64_t seed, int32_t *__restrict__ lab,
                                     double *__restrict__ Cmean)
{
    extern __shared__ unsigned char sm_raw[];
    double *h = (double *)sm_raw; // [K]
    int *cnt = (int *)(h + K);    // [K]
    __shared__ double sh_s;
    const int tid = threadIdx.x;
    for (int k = tid; k < K; k += blockDim.x) cnt[k] = 0;
    __syncthreads();
    for (int64_t f = tid; f < F; f += blockDim.x) {
        int k = (int)(label_hash(seed ^ 0x1234567ull, f, 0) % (uint32_t)K);
        lab[f] = k;
        atomicAdd(&cnt[k], 1);
    }
    __syncthreads();
    int kept = 0;
    for (int i = 1; i < iter; i++) {
        for (int k = tid; k < K; k += blockDim.x) {
            PhiloxStream rs(seed, (uint32_t)k, 0xC0000000u, (uint32_t)i);
            h[k] = rs.gamma(alpha + (double)cnt[k]);
        }
        __syncthreads();
        if (tid == 0) { double s = 0; for (int k = 0; k < K; k++) s += h[k]; sh_s = s; }
        __syncthreads();
        for (int k = tid; k < K; k += blockDim.x) cnt[k] = 0;
        __syncthreads();
        for (int64_t f = tid; f < F; f += blockDim.x) {
            const int vf = v[f];
            double tot = 0.0;
            for (int k = 0; k < K; k++) {
                double w = W[f + F * k];
                tot += h[k] * (vf ? w : 1.0 - w);
            }
            philox4 r = philox4x32_10((uint32_t)f, (uint32_t)(f >> 32), 0xD0000000u, (uint32_t)i,
                                      (uint32_t)seed, (uint32_t)(seed >> 32));
            double target = u64_to_unif(r.x, r.y) * tot, c = 0.0;
            int kz = K - 1;
            for (int k = 0; k < K; k++) {
                double w = W[f + F * k];
                c += h[k] * (vf ? w : 1.0 - w);
                if (target < c) { kz = k; break; }
            }
            lab[f] = kz;
            atomicAdd(&cnt[kz], 1);
            if (i >= nburnin) Cmean[f + F * kz] += 1.0;
        }
        if (i >= nburnin) kept++;
        __syncthreads();
    }
    (void)sh_s;
    if (kept > 0)
        for (int64_t j = tid; j < F * K; j += blockDim.x) Cmean[j] /= kept;
}

// ---------------------------------------------------------------------------
// loglikelihood.Bernoulli (R/commons.R:17-24) over a column chunk of V_test
// ---------------------------------------------------------------------------
__global__ void predictive_ll_kernel(const int32_t *__restrict__ Vt, int64_t F, int64_t ncols,
                                     int64_t n0, int K, int Kp, const double *__restrict__ EW,
                                     const double *__restrict__ EH, double *ll,
                                     unsigned long long *cnt)
{
    __shared__ double sh[32];
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double t = 0.0;
    if (i < F * ncols) {
        int32_t v = Vt[i];
        if (v != NBMF_NA_INTEGER) {
            int64_t f = i % F, n = n0 + i / F;
            double p = 0.0;
            for (int k = 0; k < K; k++) p += EW[f * Kp + k] * EH[n * Kp + k];
            t = log(v * p + (1 - v) * (1.0 - p));
            atomicAdd(cnt, 1ull);
        }
    }
    t = block_sum(t, sh);
    if (threadIdx.x == 0 && t != 0.0) atomicAdd(ll, t);
}

// ---------------------------------------------------------------------------
// synthetic generator (R/utils.R:7-16, xp_paper_basic_example.R:22-30)
// ---------------------------------------------------------------------------
__global__ void synth_W_kernel(int64_t F, int Kt, double a, uint64_t seed, float *__restrict__ Wt)
{
    int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    double s = 0.0;
    for (int k = 0; k < Kt; k++) {
        PhiloxStream rs(seed, (uint32_t)f, (uint32_t)k, 0xA0000001u);
        double g = rs.gamma(a);
        Wt[f * Kt + k] = (float)g;
        s += g;
    }
    for (int k = 0; k < Kt; k++) Wt[f * Kt + k] = (float)((double)Wt[f * Kt + k] / s);
}
__global__ void synth_H_kernel(int64_t N, int64_t n_offset, int Kt, double a, double b,
                               uint64_t seed, float *__restrict__ Ht)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * Kt) return;
    int64_t n = i / Kt + n_offset;
    int k = (int)(i % Kt);
    PhiloxStream rs(seed, (uint32_t)n, (uint32_t)k | ((uint32_t)(n >> 32) << 16), 0xA0000002u);
    double x = rs.gamma(a), y = rs.gamma(b);
    Ht[i] = (float)(x / (x + y));
}
// one thread per (column, word): 32 Bernoulli draws
__global__ void synth_V_kernel(int64_t F, int64_t N, int64_t Fw, int64_t n_offset, int Kt,
                               const float *__restrict__ Wt, const float *__restrict__ Ht,
                               uint64_t seed, uint32_t *__restrict__ vc, unsigned long long *ones)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * Fw) return;
    int64_t n = i / Fw, w = i % Fw, ng = n + n_offset;
    float h[16];
    for (int k = 0; k < Kt; k++) h[k] = Ht[n * Kt + k];
    uint32_t word = 0;
    for (int q = 0; q < 8; q++) {
        philox4 r = philox4x32_10((uint32_t)(w * 8 + q), (uint32_t)ng, (uint32_t)(ng >> 32),
                                  0xA0000003u, (uint32_t)seed, (uint32_t)(seed >> 32));
        uint32_t rr[4] = {r.x, r.y, r.z, r.w};
        for (int j = 0; j < 4; j++) {
            int64_t f = w * 32 + q * 4 + j;
            if (f < F) {
                float p = 0.f;
                for (int k = 0; k < Kt; k++) p += Wt[f * Kt + k] * h[k];
                if (u32_to_unif(rr[j]) < p) word |= 1u << (q * 4 + j);
            }
        }
    }
    vc[i] = word;
    if (word) atomicAdd(ones, (unsigned long long)__popc(word));
}
