// nbmf_extra.cu -- the remaining registered entry points of the path (src/RcppExports.cpp:316-339) behind the
// C ABI: the posterior-mean reducers of the collapsed samplers, the conditional draws of
// samples_VWH_DirBer, the 11-argument lower_bound_aspect and sample_gibbs_cpp.  Same conventions as
// nbmf_cuda.cu: host matrices column-major, int32 with NA = INT_MIN, fp64 outputs, status codes.
// (included at the end of nbmf_cuda.cu: one translation unit with the kernels it reuses)
#pragma once

// ---------------------------------------------------------------------------
// compute_expectation_W_Rcpp (collapsed_gibbs_z.cpp:178-197)
// ---------------------------------------------------------------------------
__global__ void exw_hist_kernel(const int32_t *__restrict__ Z, int64_t F, int64_t count, int Kq, int *cnt, unsigned long long *bad)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int32_t k = Z[i];
    if (k == NBMF_NA_INTEGER) return;                       // matrix_to_tensor: an all-zero indicator (commons.cpp:21-24)
    if (k < 0 || k >= Kq) { atomicAdd(bad, 1ull); return; }
    atomicAdd(&cnt[(i % F) * Kq + k], 1);
}
__global__ void exw_finish_kernel(const int *__restrict__ cnt, int64_t F, int Kq, int J, double gamma, double *__restrict__ EW)
{
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    double s = 0.0;
    for (int k = 0; k < Kq; k++) s += gamma + (double)cnt[f * Kq + k] / J;   // :190-191: gamma + mean over samples of the row counts
    for (int k = 0; k < Kq; k++) EW[f + F * k] = (gamma + (double)cnt[f * Kq + k] / J) / s;
}

extern "C" int nbmf_compute_expectation_W(nbmf_handle *h, const int32_t *Z_samples, int64_t F, int64_t N, int J,
                                          double gamma, int Kq, double *E_W)
{
    if (!h) return NBMF_ERR_ARG;
    if (!Z_samples || !E_W || F <= 0 || N <= 0 || J <= 0 || Kq <= 0) return fail(h, NBMF_ERR_ARG, "nbmf_compute_expectation_W: bad arguments");
    CK(cudaSetDevice(h->device));
    DevBuf<int32_t> dz; DevBuf<int> cnt; DevBuf<double> out; DevBuf<unsigned long long> bad;
    const int64_t slice = F * N;
    CK(dz.alloc((size_t)slice)); CK(cnt.alloc((size_t)F * Kq)); CK(out.alloc((size_t)F * Kq)); CK(bad.alloc(1));
    CK(cudaMemsetAsync(cnt, 0, (size_t)F * Kq * 4, h->stream));
    CK(cudaMemsetAsync(bad, 0, 8, h->stream));
    for (int j = 0; j < J; j++) {
        CK(cudaMemcpyAsync(dz, Z_samples + (size_t)j * slice, (size_t)slice * 4, cudaMemcpyHostToDevice, h->stream));
        exw_hist_kernel<<<nblk(slice, 256), 256, 0, h->stream>>>(dz, F, slice, Kq, cnt, bad);
        LAUNCH_CHECK();
        CK(cudaStreamSynchronize(h->stream));              // the host slice may be pageable
    }
    unsigned long long nbad = 0;
    CK(cudaMemcpy(&nbad, bad, 8, cudaMemcpyDeviceToHost));
    if (nbad) return fail(h, NBMF_ERR_ARG, "Z_samples contains labels outside 0..Kq-1");
    exw_finish_kernel<<<nblk(F, 256), 256, 0, h->stream>>>(cnt, F, Kq, J, gamma, out);
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(E_W, out, (size_t)F * Kq * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return NBMF_OK;
}

// ---------------------------------------------------------------------------
// compute_expectation_H_Rcpp (collapsed_gibbs_z.cpp:230-272): E_H(n,k) = mean_j (alpha + ones_jnk) /
// (alpha + beta + total_jnk) for k < max label -- the last column stays 0 (:236-242).  One block per
// column builds the two histograms of a sample in shared memory; the sum over j runs in the
// reference's order, so the result is bit-identical.
// ---------------------------------------------------------------------------
__global__ void exh_accumulate_kernel(const int32_t *__restrict__ Z, const int32_t *__restrict__ V, int64_t F, int64_t N,
                                      int Kq, double alpha, double beta, double *__restrict__ acc)
{
    extern __shared__ int sh_cnt[];   // [2][Kq]: ones, total
    const int64_t n = blockIdx.x;
    for (int i = threadIdx.x; i < 2 * Kq; i += blockDim.x) sh_cnt[i] = 0;
    __syncthreads();
    for (int64_t f = threadIdx.x; f < F; f += blockDim.x) {
        const int32_t v = V[f + F * n];
        if (v == NBMF_NA_INTEGER) continue;
        const int32_t z = Z[f + F * n];
        if (z < 0 || z >= Kq) continue;                    // NA label: equal to no k
        atomicAdd(&sh_cnt[Kq + z], 1);
        if (v) atomicAdd(&sh_cnt[z], v);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < Kq - 1; k += blockDim.x)
        acc[n + N * k] += (alpha + (double)sh_cnt[k]) / (alpha + beta + (double)sh_cnt[Kq + k]);
}
__global__ void exh_finish_kernel(double *acc, int64_t count, int J)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) acc[i] = acc[i] / J;
}

extern "C" int nbmf_compute_expectation_H(nbmf_handle *h, const int32_t *Z_samples, const int32_t *V, int64_t F,
                                          int64_t N, int J, double alpha, double beta, int Kq, double *E_H)
{
    if (!h) return NBMF_ERR_ARG;
    if (!Z_samples || !V || !E_H || F <= 0 || N <= 0 || J <= 0 || Kq <= 0) return fail(h, NBMF_ERR_ARG, "nbmf_compute_expectation_H: bad arguments");
    CK(cudaSetDevice(h->device));
    DevBuf<int32_t> dz, dv; DevBuf<double> acc;
    const int64_t slice = F * N;
    CK(dz.alloc((size_t)slice)); CK(dv.alloc((size_t)slice)); CK(acc.alloc((size_t)N * Kq));
    CK(cudaMemsetAsync(acc, 0, (size_t)N * Kq * 8, h->stream));
    CK(cudaMemcpyAsync(dv, V, (size_t)slice * 4, cudaMemcpyHostToDevice, h->stream));
    for (int j = 0; j < J; j++) {
        CK(cudaMemcpyAsync(dz, Z_samples + (size_t)j * slice, (size_t)slice * 4, cudaMemcpyHostToDevice, h->stream));
        exh_accumulate_kernel<<<(unsigned)N, 256, 2 * Kq * sizeof(int), h->stream>>>(dz, dv, F, N, Kq, alpha, beta, acc);
        LAUNCH_CHECK();
        CK(cudaStreamSynchronize(h->stream));
    }
    exh_finish_kernel<<<nblk(N * (int64_t)(Kq - 1), 256), 256, 0, h->stream>>>(acc, N * (int64_t)(Kq - 1), J);
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(E_H, acc, (size_t)N * Kq * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return NBMF_OK;
}

// ---------------------------------------------------------------------------
// lower_bound_aspect (collapsed_gibbs_z_VB.cpp:339-369) with its eleven arguments, for sizes where the
// E_Z cube exists: the seven terms as the reference states them (:173-294), no identity used.
// terms7 = pW, pZ, pH, pV, qW, qZ, qH.
// ---------------------------------------------------------------------------
__global__ void lbfull_rows_kernel(const double *__restrict__ ElogW, const double *__restrict__ gvb, int64_t F, int K,
                                   double gamma, double *out /* [0] pW, [4] qW */)
{
    __shared__ double sh[32];
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double pw = 0.0, qw = 0.0;
    if (f < F) {
        double s = 0.0, se = 0.0, q = 0.0;
        for (int k = 0; k < K; k++) {
            const double g = gvb[f + F * k], e = ElogW[f + F * k];
            s += g; se += e; q += -lgamma(g) + (g - 1.0) * e;
        }
        pw = (lgamma(K * gamma) - K * lgamma(gamma)) + (gamma - 1.0) * se;   // :177-180
        qw = lgamma(s) + q;                                                   // :189-195
    }
    double t = block_sum(pw, sh);
    if (threadIdx.x == 0) atomicAdd(&out[0], t);
    t = block_sum(qw, sh);
    if (threadIdx.x == 0) atomicAdd(&out[4], t);
}
__global__ void lbfull_cols_kernel(const double *__restrict__ ElogH, const double *__restrict__ Elog1H,
                                   const double *__restrict__ avb, const double *__restrict__ bvb, int64_t N, int K,
                                   double alpha, double beta, double *out /* [2] pH, [6] qH */)
{
    __shared__ double sh[32];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double ph = 0.0, qh = 0.0;
    if (i < N * K) {
        const double a = avb[i], b = bvb[i], eh = ElogH[i], e1 = Elog1H[i];
        ph = (lgamma(alpha + beta) - lgamma(alpha) - lgamma(beta)) + (alpha - 1.0) * eh + (beta - 1.0) * e1;   // :249-251
        qh = lgamma(a + b) - lgamma(a) - lgamma(b) + (a - 1.0) * eh + (b - 1.0) * e1;                          // :262-270
    }
    double t = block_sum(ph, sh);
    if (threadIdx.x == 0) atomicAdd(&out[2], t);
    t = block_sum(qh, sh);
    if (threadIdx.x == 0) atomicAdd(&out[6], t);
}
// pZ (:201-222), qZ (:225-240, NaN on an exact zero as in the reference), pV (:276-294); E_Z in
// arma::cube layout f + F k + F K n
__global__ void lbfull_entries_kernel(const double *__restrict__ EZ, const int32_t *__restrict__ V,
                                      const double *__restrict__ ElogW, const double *__restrict__ ElogH,
                                      const double *__restrict__ Elog1H, int64_t F, int64_t N, int K, double *out)
{
    __shared__ double sh[32];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double pz = 0.0, qz = 0.0, pv = 0.0;
    if (i < F * N) {
        const int64_t f = i % F, n = i / F;
        const int32_t v = V[i];
        if (v != NBMF_NA_INTEGER) {
            for (int k = 0; k < K; k++) {
                const double z = EZ[f + F * k + F * (int64_t)K * n];
                pz += z * ElogW[f + F * k];
                qz += z * log(z);
                pv += z * (v * ElogH[n + N * k] + (1 - v) * Elog1H[n + N * k]);
            }
        }
    }
    double t = block_sum(pz, sh);
    if (threadIdx.x == 0) atomicAdd(&out[1], t);
    t = block_sum(qz, sh);
    if (threadIdx.x == 0) atomicAdd(&out[5], t);
    t = block_sum(pv, sh);
    if (threadIdx.x == 0) atomicAdd(&out[3], t);
}

extern "C" int nbmf_lower_bound_aspect_full(nbmf_handle *h, const double *E_Z, const double *E_log_W, const double *E_log_H,
                                            const double *E_log_1_H, const double *gamma_vb, const double *alpha_vb,
                                            const double *beta_vb, double alpha, double beta, double gamma,
                                            const int32_t *V, int64_t F, int64_t N, int K, double *lb, double *terms7)
{
    if (!h) return NBMF_ERR_ARG;
    if (!E_Z || !E_log_W || !E_log_H || !E_log_1_H || !gamma_vb || !alpha_vb || !beta_vb || !V || F <= 0 || N <= 0 || K <= 0)
        return fail(h, NBMF_ERR_ARG, "nbmf_lower_bound_aspect_full: bad arguments");
    if ((double)F * N * K * 8 > 32e9) return fail(h, NBMF_ERR_UNSUPPORTED, "E_Z does not fit on the device: use nbmf_lower_bound (E_Z-free)");
    CK(cudaSetDevice(h->device));
    DevBuf<double> dz, dw, dh, d1, dg, da, db, out;
    DevBuf<int32_t> dv;
    const size_t fk = (size_t)F * K, nk = (size_t)N * K, fnk = (size_t)F * N * K;
    CK(dz.alloc(fnk)); CK(dw.alloc(fk)); CK(dg.alloc(fk)); CK(dh.alloc(nk)); CK(d1.alloc(nk)); CK(da.alloc(nk)); CK(db.alloc(nk));
    CK(dv.alloc((size_t)F * N)); CK(out.alloc(8));
    CK(cudaMemsetAsync(out, 0, 64, h->stream));
    CK(cudaMemcpyAsync(dz, E_Z, fnk * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dw, E_log_W, fk * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dg, gamma_vb, fk * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dh, E_log_H, nk * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d1, E_log_1_H, nk * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(da, alpha_vb, nk * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(db, beta_vb, nk * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dv, V, (size_t)F * N * 4, cudaMemcpyHostToDevice, h->stream));
    lbfull_rows_kernel<<<nblk(F, 256), 256, 0, h->stream>>>(dw, dg, F, K, gamma, out);
    LAUNCH_CHECK();
    lbfull_cols_kernel<<<nblk(N * (int64_t)K, 256), 256, 0, h->stream>>>(dh, d1, da, db, N, K, alpha, beta, out);
    LAUNCH_CHECK();
    lbfull_entries_kernel<<<nblk(F * N, 256), 256, 0, h->stream>>>(dz, dv, dw, dh, d1, F, N, K, out);
    LAUNCH_CHECK();
    double t[8];
    CK(cudaMemcpyAsync(t, out, 64, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (terms7) for (int i = 0; i < 7; i++) terms7[i] = t[i];
    if (lb) *lb = t[0] + t[1] + t[2] + t[3] - t[4] - t[5] - t[6];   // :364-365
    return NBMF_OK;
}

// ---------------------------------------------------------------------------
// samples_VWH_DirBer (collapsed_gibbs_z.cpp:565-625): for each kept label matrix, W_f ~ Dir(gamma +
// n_total_f) as normalised Gammas, H_nk ~ Beta(alpha + ones, beta + zeros), V ~ Bernoulli(W H^T).
// R's stream is replaced by Philox keyed (seed; sample, row / column, component); outputs are the
// reference's cubes (column-major slices).
// ---------------------------------------------------------------------------
__global__ void svwh_draw_W_kernel(const int32_t *__restrict__ cF, int64_t F, int K, double gamma, uint64_t seed, uint32_t j,
                                   double *__restrict__ Wd /* F x K column-major */)
{
    const int64_t f = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (f >= F) return;
    double s = 0.0;
    for (int k0 = 0; k0 < K; k0 += 32) {
        const int k = k0 + lane;
        double g = 0.0;
        if (k < K) {
            PhiloxStream rs(seed, (uint32_t)f, (uint32_t)k | 0x40000000u, j);
            g = rs.gamma(gamma + (double)cF[f * K + k]);
            Wd[f + F * k] = g;
        }
        s += g;
    }
    s = warp_sum(s);
    for (int k = lane; k < K; k += 32) Wd[f + F * k] = (s > 0.0) ? Wd[f + F * k] / s : 1.0 / K;
}
__global__ void svwh_draw_H_kernel(const int32_t *__restrict__ cN, int64_t N, int K, double alpha, double beta,
                                   uint64_t seed, uint32_t j, double *__restrict__ Hd /* N x K column-major */)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * K) return;
    const int64_t n = i / K;
    const int k = (int)(i % K);
    PhiloxStream rs(seed, (uint32_t)n, (uint32_t)k | 0x80000000u | ((uint32_t)(n >> 32) << 16), j);
    const double x = rs.gamma(alpha + (double)cN[(2 * n + 1) * K + k]);
    const double y = rs.gamma(beta + (double)cN[(2 * n) * K + k]);
    Hd[n + N * k] = x / (x + y);
}
__global__ void svwh_draw_V_kernel(const double *__restrict__ Wd, const double *__restrict__ Hd, int64_t F, int64_t N, int K,
                                   uint64_t seed, uint32_t j, int32_t *__restrict__ Vd)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * N) return;
    const int64_t f = i % F, n = i / F;
    double p = 0.0;
    for (int k = 0; k < K; k++) p += Wd[f + F * k] * Hd[n + N * k];
    const philox4 r = philox4x32_10((uint32_t)f, (uint32_t)n, (uint32_t)(n >> 32), j ^ 0xE0000000u, (uint32_t)seed, (uint32_t)(seed >> 32));
    Vd[i] = u64_to_unif(r.x, r.y) < p ? 1 : 0;             // :611-616
}
__global__ void svwh_counts_kernel(const int32_t *__restrict__ Z, const int32_t *__restrict__ V, int64_t F, int64_t N, int K,
                                   int32_t *cF, int32_t *cN, unsigned long long *bad)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * N) return;
    const int32_t v = V[i];
    if (v == NBMF_NA_INTEGER) return;                       // compute_counts skips NA entries of V (:45)
    const int32_t k = Z[i];
    if (k < 0 || k >= K) { atomicAdd(bad, 1ull); return; }
    atomicAdd(&cF[(i % F) * K + k], 1);
    atomicAdd(&cN[(2 * (i / F) + (v ? 1 : 0)) * K + k], 1);
}

extern "C" int nbmf_samples_VWH_DirBer(nbmf_handle *h, const int32_t *V, const int32_t *Z_samples, int64_t F, int64_t N,
                                       int J, int K, double gamma, double alpha, double beta, uint64_t seed,
                                       double *W_samples, double *H_samples, int32_t *V_samples)
{
    if (!h) return NBMF_ERR_ARG;
    if (!V || !Z_samples || F <= 0 || N <= 0 || J <= 0 || K <= 0 || !(gamma > 0) || !(alpha > 0) || !(beta > 0))
        return fail(h, NBMF_ERR_ARG, "nbmf_samples_VWH_DirBer: bad arguments");
    CK(cudaSetDevice(h->device));
    DevBuf<int32_t> dz, dv, cF, cN, vs; DevBuf<double> Wd, Hd; DevBuf<unsigned long long> bad;
    const int64_t slice = F * N;
    CK(dz.alloc((size_t)slice)); CK(dv.alloc((size_t)slice)); CK(vs.alloc((size_t)slice));
    CK(cF.alloc((size_t)F * K)); CK(cN.alloc((size_t)2 * N * K)); CK(Wd.alloc((size_t)F * K)); CK(Hd.alloc((size_t)N * K)); CK(bad.alloc(1));
    CK(cudaMemsetAsync(bad, 0, 8, h->stream));
    CK(cudaMemcpyAsync(dv, V, (size_t)slice * 4, cudaMemcpyHostToDevice, h->stream));
    for (int j = 0; j < J; j++) {
        CK(cudaMemcpyAsync(dz, Z_samples + (size_t)j * slice, (size_t)slice * 4, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemsetAsync(cF, 0, (size_t)F * K * 4, h->stream));
        CK(cudaMemsetAsync(cN, 0, (size_t)2 * N * K * 4, h->stream));
        svwh_counts_kernel<<<nblk(slice, 256), 256, 0, h->stream>>>(dz, dv, F, N, K, cF, cN, bad);
        LAUNCH_CHECK();
        svwh_draw_W_kernel<<<nblk(F * 32, 256), 256, 0, h->stream>>>(cF, F, K, gamma, seed, (uint32_t)j, Wd);
        LAUNCH_CHECK();
        svwh_draw_H_kernel<<<nblk(N * (int64_t)K, 256), 256, 0, h->stream>>>(cN, N, K, alpha, beta, seed, (uint32_t)j, Hd);
        LAUNCH_CHECK();
        if (W_samples) CK(cudaMemcpyAsync(W_samples + (size_t)j * F * K, Wd, (size_t)F * K * 8, cudaMemcpyDeviceToHost, h->stream));
        if (H_samples) CK(cudaMemcpyAsync(H_samples + (size_t)j * N * K, Hd, (size_t)N * K * 8, cudaMemcpyDeviceToHost, h->stream));
        if (V_samples) {
            svwh_draw_V_kernel<<<nblk(slice, 256), 256, 0, h->stream>>>(Wd, Hd, F, N, K, seed, (uint32_t)j, vs);
            LAUNCH_CHECK();
            CK(cudaMemcpyAsync(V_samples + (size_t)j * slice, vs, (size_t)slice * 4, cudaMemcpyDeviceToHost, h->stream));
        }
        CK(cudaStreamSynchronize(h->stream));
    }
    unsigned long long nbad = 0;
    CK(cudaMemcpy(&nbad, bad, 8, cudaMemcpyDeviceToHost));
    if (nbad) return fail(h, NBMF_ERR_ARG, "Z_samples contains labels outside 0..K-1");
    return NBMF_OK;
}

// Gamma / Beta / Dirichlet draws of the device samplers, exposed for the distribution tests
// (tests/test_gpu_rng.py): which = 0 Gamma(a) fp64 stream, 1 Gamma(a) fp32 stream (gibbs_fast.cuh),
// 2 Beta(a, b) fp32 stream as the Gibbs draw kernel forms it, 3 the same draw as its logit log(H / (1 - H)) (Beta with
// shapes << 1 puts most of its mass closer to 0 and 1 than a double resolves).
__global__ void rng_probe_kernel(int which, double a, double b, uint64_t seed, int64_t n, double *out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (which == 0) {
        PhiloxStream rs(seed, (uint32_t)i, (uint32_t)(i >> 32), 0x77u);
        out[i] = rs.gamma(a);
    } else if (which == 1) {
        PhiloxF rs(seed, (uint32_t)i, (uint32_t)(i >> 32), 0x78u);
        float lb;
        const float x = rs.gamma_log((float)a, &lb);
        out[i] = (double)x * exp((double)lb);
    } else {
        PhiloxF rs(seed, (uint32_t)i, (uint32_t)(i >> 32), 0x79u);
        float lx, ly;
        const float x = rs.gamma_log((float)a, &lx), y = rs.gamma_log((float)b, &ly);
        const float dl = (logf(y) + ly) - (logf(x) + lx);
        out[i] = which == 3 ? -(double)dl : 1.0 / (1.0 + exp((double)dl));
    }
}
extern "C" int nbmf_debug_rng_probe(nbmf_handle *h, int which, double a, double b, uint64_t seed, int64_t n, double *out)
{
    if (!h || !out || n <= 0) return NBMF_ERR_ARG;
    CK(cudaSetDevice(h->device));
    DevBuf<double> d;
    CK(d.alloc((size_t)n));
    rng_probe_kernel<<<nblk(n, 256), 256, 0, h->stream>>>(which, a, b, seed, n, d);
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(out, d, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return NBMF_OK;
}

// FP32 FMA peak of the device (bench.py --workload c3 reports the Gibbs sweep against it)
extern "C" int nbmf_debug_ffma_peak(nbmf_handle *h, double *tflops)
{
    if (!h || !tflops) return NBMF_ERR_ARG;
    CK(cudaSetDevice(h->device));
    DevBuf<float> d;
    CK(d.alloc(4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 4096, blocks = h->sm_count * 8, threads = 512;
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0, h->stream);
        ffma_peak_kernel<<<blocks, threads, 0, h->stream>>>(d, iters);
        cudaEventRecord(e1, h->stream);
        cudaStreamSynchronize(h->stream);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * 64.0 * iters * (double)blocks * threads;
        if (ms > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *tflops = best;
    return cudaGetLastError() == cudaSuccess ? NBMF_OK : NBMF_ERR_CUDA;
}

// ---------------------------------------------------------------------------
// sample_gibbs_cpp (sample_gibbs.cpp:22-80): one column's sweep over a one-hot F x K matrix C given W
// (paper orientation), h collapsed: probs_k = (alpha + L_k) lik_fk with the row's own assignment
// removed.  as_coded != 0 reproduces the reference as it is written: the draw `ans` (:69) is never
// stored, so every row is zero after the first sweep and the result is C_init / (number of kept
// states) whatever the draws were.  as_coded == 0 stores the draw (what R's sample_gibbs_Z,
// R/MCEM - sample_gibbs.R:1-37, does) -- a sequential chain over f, one warp, lanes over k.
// ---------------------------------------------------------------------------
__global__ void sgc_kernel(const int32_t *__restrict__ v, const double *__restrict__ W, const int32_t *__restrict__ C0,
                           int64_t F, int K, double alpha, int iter, int nburnin, uint64_t seed, int as_coded,
                           int32_t *lab, double *__restrict__ Cmean, int *kept_out)
{
    extern __shared__ int sgc_L[];   // [K] column sums of C
    const int lane = threadIdx.x;
    for (int k = lane; k < K; k += 32) sgc_L[k] = 0;
    __syncwarp();
    // labels of the initial one-hot matrix (-1: an all-zero row)
    for (int64_t f = lane; f < F; f += 32) {
        int z = -1;
        for (int k = 0; k < K; k++) if (C0[f + F * k] != 0) { z = k; break; }
        lab[f] = z;
        if (z >= 0) atomicAdd(&sgc_L[z], 1);
    }
    __syncwarp();
    for (int64_t j = lane; j < F * K; j += 32) Cmean[j] = (double)C0[j];     // :47 the initial state is a sample
    int kept = 1;
    __syncwarp();
    for (int i = 1; i < iter; i++) {
        for (int64_t f = 0; f < F; f++) {
            const int z = lab[f];
            __syncwarp();
            if (lane == 0 && z >= 0) { sgc_L[z] -= 1; lab[f] = -1; }        // :60-61 remove the current value
            __syncwarp();
            if (!as_coded) {
                const int vf = v[f];
                // probs (:64-65) and an inverse-CDF draw shared by the warp
                double tot = 0.0;
                for (int k = lane; k < K; k += 32) { const double w = W[f + F * k]; tot += (alpha + sgc_L[k]) * (vf * w + (1 - vf) * (1.0 - w)); }
                tot = warp_sum(tot);
                const philox4 r = philox4x32_10((uint32_t)f, (uint32_t)(f >> 32), 0xD1000000u, (uint32_t)i, (uint32_t)seed, (uint32_t)(seed >> 32));
                const double target = u64_to_unif(r.x, r.y) * tot;
                double c = 0.0;
                int kz = K - 1;
                for (int k = 0; k < K; k++) {   // every lane walks the same scan
                    const double w = W[f + F * k];
                    c += (alpha + sgc_L[k]) * (vf * w + (1 - vf) * (1.0 - w));
                    if (target < c) { kz = k; break; }
                }
                __syncwarp();
                if (lane == 0) { lab[f] = kz; sgc_L[kz] += 1; }
                __syncwarp();
            }
        }
        if (i >= nburnin) {                                                  // :73-76
            __syncwarp();
            for (int64_t f = lane; f < F; f += 32) { const int z = lab[f]; if (z >= 0) Cmean[f + F * z] += 1.0; }
            kept++;
        }
    }
    __syncwarp();
    for (int64_t j = lane; j < F * K; j += 32) Cmean[j] /= (double)kept;     // :78
    if (lane == 0 && kept_out) *kept_out = kept;
}

extern "C" int nbmf_sample_gibbs_cpp(nbmf_handle *h, const int32_t *v_n, const double *W, const int32_t *C, int64_t F, int K,
                                     double alpha, int iter, double burnin, uint64_t seed, int as_coded, double *C_mean)
{
    if (!h) return NBMF_ERR_ARG;
    if (!v_n || !W || !C || !C_mean || F <= 0 || K <= 0 || K > 4096) return fail(h, NBMF_ERR_ARG, "nbmf_sample_gibbs_cpp: bad arguments");
    CK(cudaSetDevice(h->device));
    DevBuf<int32_t> dv, dc, lab; DevBuf<double> dW, dm;
    CK(dv.alloc((size_t)F)); CK(dc.alloc((size_t)F * K)); CK(lab.alloc((size_t)F)); CK(dW.alloc((size_t)F * K)); CK(dm.alloc((size_t)F * K));
    CK(cudaMemcpyAsync(dv, v_n, (size_t)F * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dc, C, (size_t)F * K * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dW, W, (size_t)F * K * 8, cudaMemcpyHostToDevice, h->stream));
    sgc_kernel<<<1, 32, (size_t)K * sizeof(int), h->stream>>>(dv, dW, dc, F, K, alpha, iter, (int)std::floor(iter * burnin), seed, as_coded, lab, dm, nullptr);
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(C_mean, dm, (size_t)F * K * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return NBMF_OK;
}
