// nbmf_cuda.cu -- C ABI of libnbmf_cuda (see include/nbmf_cuda.h) and the host
// orchestration of the VB / CVB0 / Gibbs hot path.  No CPU fallback: every
// compute entry point needs a CUDA device.
#include "kernels_basic.cuh"
#include "tensor_pass.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <type_traits>
#include <vector>

#define CK(call) NBMF_CUDA_CHECK(h, call)
#define LAUNCH_CHECK() CK(cudaGetLastError())

static inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

// device allocation that is released on every return path
template <class T> struct DevBuf {
    T *p = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t count) { return cudaMalloc((void **)&p, (count ? count : 1) * sizeof(T)); }
    operator T *() const { return p; }
};

static int fail(nbmf_handle *h, int code, const char *msg)
{
    if (h) h->err = msg;
    return code;
}

// ---------------------------------------------------------------------------
// lifecycle
// ---------------------------------------------------------------------------
extern "C" const char *nbmf_version(void) { return "nbmf_cuda 0.1 (sm_100a)"; }

extern "C" int nbmf_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int nbmf_create(nbmf_handle **out, const nbmf_opts *opts)
{
    if (!out) return NBMF_ERR_ARG;
    *out = nullptr;
    nbmf_handle *h = new (std::nothrow) nbmf_handle();
    if (!h) return NBMF_ERR_OOM;
    if (opts) h->opts = *opts;
    h->device = h->opts.device;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0 || h->device < 0 || h->device >= n) {
        cudaGetLastError();
        delete h;
        return NBMF_ERR_CUDA; // no device: fail loudly, there is no CPU path
    }
    cudaDeviceProp p;
    if (cudaSetDevice(h->device) != cudaSuccess || cudaGetDeviceProperties(&p, h->device) != cudaSuccess) {
        delete h;
        return NBMF_ERR_CUDA;
    }
    h->sm_count = p.multiProcessorCount;
    h->cc_major = p.major; h->cc_minor = p.minor;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete h;
        return NBMF_ERR_CUDA;
    }
    for (int i = 0; i < 8; i++) cudaEventCreate(&h->ev[i]);
    if (cudaMalloc(&h->scal, SC_COUNT * sizeof(double)) != cudaSuccess) {
        delete h;
        return NBMF_ERR_OOM;
    }
    cudaMemset(h->scal, 0, SC_COUNT * sizeof(double));
    *out = h;
    return NBMF_OK;
}

static void free_data(nbmf_handle *h)
{
    cudaFree(h->vc); cudaFree(h->mc); cudaFree(h->vr); cudaFree(h->mr);
    h->vc = h->mc = h->vr = h->mr = nullptr;
    cudaFree(h->Z); h->Z = nullptr;
}
static void free_model(nbmf_handle *h)
{
    cudaFree(h->statF); cudaFree(h->statN); cudaFree(h->A); cudaFree(h->BB);
    cudaFree(h->GF); cudaFree(h->GN); cudaFree(h->EW); cudaFree(h->EH);
    h->statF = h->statN = h->A = h->BB = h->GF = h->GN = h->EW = h->EH = nullptr;
    tp_free(h);
    h->Kc = h->Kp = 0; h->Kp_alloc_K = 0;
}

extern "C" void nbmf_destroy(nbmf_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    free_data(h); free_model(h);
    cudaFree(h->scal); cudaFree(h->lbs); cudaFree(h->stage);
    for (int i = 0; i < 8; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    for (int i = 0; i < 33; i++) if (h->pipe_ev[i]) cudaEventDestroy(h->pipe_ev[i]);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->xchg_stream) { cudaStreamDestroy(h->xchg_stream); cudaEventDestroy(h->xchg_ev[0]); cudaEventDestroy(h->xchg_ev[1]); }
    cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" const char *nbmf_last_error(const nbmf_handle *h) { return h ? h->err.c_str() : "null handle"; }
extern "C" void *nbmf_stream(nbmf_handle *h) { return h ? (void *)h->stream : nullptr; }
extern "C" int nbmf_synchronize(nbmf_handle *h)
{
    if (!h) return NBMF_ERR_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    return NBMF_OK;
}
extern "C" int nbmf_set_allreduce_hook(nbmf_handle *h, nbmf_allreduce_fn fn, void *ctx)
{
    if (!h) return NBMF_ERR_ARG;
    h->hook = fn; h->hook_ctx = ctx;
    return NBMF_OK;
}
extern "C" int nbmf_get_timing(const nbmf_handle *h, nbmf_timing *out)
{
    if (!h || !out) return NBMF_ERR_ARG;
    *out = h->timing;
    return NBMF_OK;
}
extern "C" int nbmf_get_dims(const nbmf_handle *h, int64_t *F, int64_t *N, int64_t *nobs)
{
    if (!h) return NBMF_ERR_ARG;
    if (F) *F = h->F;
    if (N) *N = h->N;
    if (nobs) *nobs = h->nobs;
    return NBMF_OK;
}

static int ensure_stage(nbmf_handle *h, size_t bytes)
{
    if (h->stage_bytes >= bytes) return NBMF_OK;
    cudaFree(h->stage); h->stage = nullptr; h->stage_bytes = 0;
    if (cudaMalloc(&h->stage, bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(h, NBMF_ERR_OOM, "device staging buffer allocation failed");
    }
    h->stage_bytes = bytes;
    return NBMF_OK;
}

static int exchange(nbmf_handle *h, double *buf, size_t count)
{
    if (!h->hook) return NBMF_OK;
    int rc = h->hook(h->hook_ctx, (void *)buf, count, (void *)h->stream);
    if (rc != 0) return fail(h, NBMF_ERR_STATE, "allreduce hook failed");
    return NBMF_OK;
}

// The same exchange started on a side stream behind the current point of the main stream, so
// that whatever the main stream launches next overlaps it; exchange_end makes the main stream
// wait for the result.  The side stream has the highest priority: the collective's few CTAs are
// placed before the persistent contraction kernel that follows, whose work queue absorbs the
// SMs that come late.
static int exchange_begin(nbmf_handle *h, double *buf, size_t count)
{
    if (!h->hook) return NBMF_OK;
    if (!h->xchg_stream) {
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&h->xchg_stream, cudaStreamNonBlocking, hi));
        CK(cudaEventCreateWithFlags(&h->xchg_ev[0], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->xchg_ev[1], cudaEventDisableTiming));
    }
    CK(cudaEventRecord(h->xchg_ev[0], h->stream));
    CK(cudaStreamWaitEvent(h->xchg_stream, h->xchg_ev[0], 0));
    int rc = h->hook(h->hook_ctx, buf, count, (void *)h->xchg_stream);
    if (rc != 0) return fail(h, NBMF_ERR_CUDA, "all-reduce hook failed");
    CK(cudaEventRecord(h->xchg_ev[1], h->xchg_stream));
    h->xchg_pending = true;
    // The collective's kernel and the cols pass become runnable at the same moment; if the
    // persistent cols kernel is placed first it holds every SM and the collective runs after it
    // instead of beside it.  20 us on the main stream decide that race (NBMF_XCHG_PAUSE_US).
    static const int pause_us = [] { const char *v = getenv("NBMF_XCHG_PAUSE_US"); return v ? atoi(v) : 20; }();
    if (pause_us > 0) {
        pause_kernel<<<1, 32, 0, h->stream>>>((unsigned)pause_us * 1000u);
        LAUNCH_CHECK();
    }
    h->timing.launches += 1;
    return NBMF_OK;
}
static int exchange_end(nbmf_handle *h)
{
    if (!h->xchg_pending) return NBMF_OK;
    h->xchg_pending = false;
    CK(cudaStreamWaitEvent(h->stream, h->xchg_ev[1], 0));
    return NBMF_OK;
}

// ---------------------------------------------------------------------------
// data
// ---------------------------------------------------------------------------
static int alloc_planes(nbmf_handle *h, int64_t F, int64_t N)
{
    CK(cudaSetDevice(h->device));
    if (h->vc && h->F == F && h->N == N) { // same shape (e.g. one C-ABI call per VB iteration): keep every buffer
        // every loader rewrites all data words with plain stores and never touches the padding
        // words, which are still zero from the allocation: nothing to clear
        cudaFree(h->Z); h->Z = nullptr;
        h->nobs = 0;
        tp_invalidate_data(h);
        return NBMF_OK;
    }
    free_data(h);
    free_model(h);
    h->F = F; h->N = N; h->nobs = 0;
    h->Fw = round_up64(F, 256) / 32;
    h->Nw = round_up64(N, 256) / 32;
    size_t cb = (size_t)N * h->Fw * 4, rb = (size_t)F * h->Nw * 4;
    if (cudaMalloc(&h->vc, cb) != cudaSuccess || cudaMalloc(&h->mc, cb) != cudaSuccess ||
        cudaMalloc(&h->vr, rb) != cudaSuccess || cudaMalloc(&h->mr, rb) != cudaSuccess) {
        cudaGetLastError();
        free_data(h);
        return fail(h, NBMF_ERR_OOM, "bit plane allocation failed");
    }
    CK(cudaMemsetAsync(h->vc, 0, cb, h->stream)); CK(cudaMemsetAsync(h->mc, 0, cb, h->stream));
    CK(cudaMemsetAsync(h->vr, 0, rb, h->stream)); CK(cudaMemsetAsync(h->mr, 0, rb, h->stream));
    return NBMF_OK;
}

static int build_row_planes(nbmf_handle *h)
{
    int64_t rg = (h->N + 31) / 32, cw = (h->F + 31) / 32;
    dim3 grid((unsigned)((cw + TB_WORDS - 1) / TB_WORDS), (unsigned)((rg + TB_GROUPS - 1) / TB_GROUPS));
    transpose_bits_kernel<<<grid, TB_WORDS * 32, 0, h->stream>>>(h->vc, h->N, h->Fw, h->F, h->vr, h->Nw);
    LAUNCH_CHECK();
    if (h->has_na)
        transpose_bits_kernel<<<grid, TB_WORDS * 32, 0, h->stream>>>(h->mc, h->N, h->Fw, h->F, h->mr, h->Nw);
    else   // every entry observed: the transposed mask is known without reading anything
        fill_mask_kernel<<<nblk(h->F * h->Nw, 256), 256, 0, h->stream>>>(h->mr, h->F, h->Nw, h->N);
    LAUNCH_CHECK();
    return NBMF_OK;
}

// the part of the row-packed plane that columns [c0, c1) contribute (c0 a multiple of 32); the
// data has no NA on this path, so the mask planes are filled up front
static int build_row_planes_block(nbmf_handle *h, int64_t c0, int64_t c1)
{
    int64_t rg = (c1 - c0 + 31) / 32, cw = (h->F + 31) / 32;
    dim3 grid((unsigned)((cw + TB_WORDS - 1) / TB_WORDS), (unsigned)((rg + TB_GROUPS - 1) / TB_GROUPS));
    transpose_bits_kernel<<<grid, TB_WORDS * 32, 0, h->stream>>>(h->vc + c0 * h->Fw, c1 - c0, h->Fw, h->F,
                                                               h->vr + c0 / 32, h->Nw);
    LAUNCH_CHECK();
    return NBMF_OK;
}

extern "C" int nbmf_set_data_i32(nbmf_handle *h, const int32_t *V, int64_t F, int64_t N)
{
    if (!h) return NBMF_ERR_ARG;
    if (!V || F <= 0 || N <= 0) return fail(h, NBMF_ERR_ARG, "nbmf_set_data_i32: bad arguments");
    int rc = alloc_planes(h, F, N);
    if (rc) return rc;
    // stream the int32 matrix through a bounded staging buffer, whole columns at a time
    int64_t cols_per = std::max<int64_t>(1, (int64_t)(256ll << 20) / (F * 4));
    cols_per = std::min(cols_per, N);
    rc = ensure_stage(h, (size_t)cols_per * F * 4 + 64);
    if (rc) return rc;
    unsigned long long *cnt = (unsigned long long *)h->scal + SC_TMP; // [0]=nobs [1]=bad
    CK(cudaMemsetAsync(cnt, 0, 16, h->stream));
    for (int64_t n0 = 0; n0 < N; n0 += cols_per) {
        int64_t nc = std::min(cols_per, N - n0);
        CK(cudaMemcpyAsync(h->stage, V + n0 * F, (size_t)nc * F * 4, cudaMemcpyHostToDevice, h->stream));
        pack_cols_kernel<<<nblk(nc * h->Fw * 32, 256), 256, 0, h->stream>>>(
            (const int32_t *)h->stage, F, nc, n0, h->Fw, h->vc, h->mc, cnt, cnt + 1);
        LAUNCH_CHECK();
        CK(cudaStreamSynchronize(h->stream)); // the host buffer may be pageable; keep it simple
    }
    unsigned long long res[2];
    CK(cudaMemcpyAsync(res, cnt, 16, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (res[1]) return fail(h, NBMF_ERR_ARG, "V contains values other than 0, 1 and NA");
    h->nobs = (int64_t)res[0];
    h->has_na = (h->nobs != F * N);
    rc = build_row_planes(h);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return NBMF_OK;
}

extern "C" int nbmf_set_data_bits(nbmf_handle *h, const uint32_t *vbits, const uint32_t *mbits,
                                  int64_t F, int64_t N)
{
    if (!h) return NBMF_ERR_ARG;
    if (!vbits || F <= 0 || N <= 0) return fail(h, NBMF_ERR_ARG, "nbmf_set_data_bits: bad arguments");
    int rc = alloc_planes(h, F, N);
    if (rc) return rc;
    int64_t fw_in = (F + 31) / 32;
    CK(cudaMemcpy2DAsync(h->vc, (size_t)h->Fw * 4, vbits, (size_t)fw_in * 4, (size_t)fw_in * 4,
                         (size_t)N, cudaMemcpyHostToDevice, h->stream));
    if (mbits) {
        CK(cudaMemcpy2DAsync(h->mc, (size_t)h->Fw * 4, mbits, (size_t)fw_in * 4, (size_t)fw_in * 4,
                             (size_t)N, cudaMemcpyHostToDevice, h->stream));
        unsigned long long *cnt = (unsigned long long *)h->scal + SC_TMP;
        CK(cudaMemsetAsync(cnt, 0, 8, h->stream));
        popcount_kernel<<<nblk(N * h->Fw, 256), 256, 0, h->stream>>>(h->mc, N * h->Fw, cnt);
        LAUNCH_CHECK();
        unsigned long long c = 0;
        CK(cudaMemcpyAsync(&c, cnt, 8, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        h->nobs = (int64_t)c;
        h->has_na = (h->nobs != F * N);
    } else {
        fill_mask_kernel<<<nblk(N * h->Fw, 256), 256, 0, h->stream>>>(h->mc, N, h->Fw, F);
        LAUNCH_CHECK();
        h->has_na = false;
        h->nobs = F * N;
    }
    rc = build_row_planes(h);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return NBMF_OK;
}

extern "C" int nbmf_get_data_bits(nbmf_handle *h, uint32_t *vbits, uint32_t *mbits)
{
    if (!h || !h->vc) return h ? fail(h, NBMF_ERR_STATE, "no data") : NBMF_ERR_ARG;
    CK(cudaSetDevice(h->device));
    int64_t fw = (h->F + 31) / 32;
    if (vbits)
        CK(cudaMemcpy2DAsync(vbits, (size_t)fw * 4, h->vc, (size_t)h->Fw * 4, (size_t)fw * 4,
                             (size_t)h->N, cudaMemcpyDeviceToHost, h->stream));
    if (mbits)
        CK(cudaMemcpy2DAsync(mbits, (size_t)fw * 4, h->mc, (size_t)h->Fw * 4, (size_t)fw * 4,
                             (size_t)h->N, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return NBMF_OK;
}

extern "C" int nbmf_generate_synthetic(nbmf_handle *h, int64_t F, int64_t N, int Ktrue, double a,
                                       double b, uint64_t seed)
{
    if (!h) return NBMF_ERR_ARG;
    if (F <= 0 || N <= 0 || Ktrue <= 0 || Ktrue > 16 || !(a > 0) || !(b > 0))
        return fail(h, NBMF_ERR_ARG, "nbmf_generate_synthetic: bad arguments (Ktrue <= 16)");
    int rc = alloc_planes(h, F, N);
    if (rc) return rc;
    DevBuf<float> Wt, Ht;
    CK(Wt.alloc((size_t)F * Ktrue));
    CK(Ht.alloc((size_t)N * Ktrue));
    unsigned long long *cnt = (unsigned long long *)h->scal + SC_TMP;
    CK(cudaMemsetAsync(cnt, 0, 16, h->stream));
    synth_W_kernel<<<nblk(F, 256), 256, 0, h->stream>>>(F, Ktrue, a, seed, Wt);
    LAUNCH_CHECK();
    synth_H_kernel<<<nblk(N * Ktrue, 256), 256, 0, h->stream>>>(N, h->opts.n_offset, Ktrue, a, b, seed, Ht);
    LAUNCH_CHECK();
    synth_V_kernel<<<nblk(N * h->Fw, 256), 256, 0, h->stream>>>(F, N, h->Fw, h->opts.n_offset, Ktrue,
                                                               Wt, Ht, seed, h->vc, cnt);
    LAUNCH_CHECK();
    fill_mask_kernel<<<nblk(N * h->Fw, 256), 256, 0, h->stream>>>(h->mc, N, h->Fw, F);
    LAUNCH_CHECK();
    h->has_na = false;
    h->nobs = F * N;
    rc = build_row_planes(h);
    CK(cudaStreamSynchronize(h->stream));
    return rc;
}

// ---------------------------------------------------------------------------
// model buffers / statistics
// ---------------------------------------------------------------------------
static void choose_kchunks(int K, int *KC, int *NKC)
{
    if (K <= 4) { *KC = 4; *NKC = 1; }
    else if (K <= 8) { *KC = 8; *NKC = 1; }
    else {
        *KC = 16;
        int n = 1;
        while (16 * n < K) n <<= 1;
        *NKC = n;
    }
}

static int ensure_model(nbmf_handle *h, int Kc)
{
    if (!h->vc) return fail(h, NBMF_ERR_STATE, "no data: call nbmf_set_data_* first");
    if (Kc <= 0 || Kc > 512) return fail(h, NBMF_ERR_ARG, "K must be in 1..512");
    if (h->Kc == Kc && h->statF) return NBMF_OK;
    CK(cudaSetDevice(h->device));
    if (h->statF && h->Kp_alloc_K == Kc) { // buffers of the right size exist: only reset the statistics
        h->Kc = Kc;
        CK(cudaMemsetAsync(h->statF, 0, (size_t)h->F * h->Kp * 8, h->stream));
        CK(cudaMemsetAsync(h->statN, 0, (size_t)2 * h->N * h->Kp * 8, h->stream));
        return NBMF_OK;
    }
    free_model(h);
    choose_kchunks(Kc, &h->KC, &h->NKC);
    h->Kc = Kc; h->Kp = h->KC * h->NKC; h->Kp_alloc_K = Kc;
    size_t fb = (size_t)h->F * h->Kp * 8, nb = (size_t)2 * h->N * h->Kp * 8;
    if (cudaMalloc(&h->statF, fb) != cudaSuccess || cudaMalloc(&h->statN, nb) != cudaSuccess ||
        cudaMalloc(&h->A, fb) != cudaSuccess || cudaMalloc(&h->BB, nb) != cudaSuccess ||
        cudaMalloc(&h->GF, fb) != cudaSuccess || cudaMalloc(&h->GN, nb) != cudaSuccess ||
        cudaMalloc(&h->EW, fb) != cudaSuccess || cudaMalloc(&h->EH, nb / 2) != cudaSuccess) {
        cudaGetLastError();
        free_model(h);
        return fail(h, NBMF_ERR_OOM, "model buffer allocation failed");
    }
    CK(cudaMemsetAsync(h->statF, 0, fb, h->stream));
    CK(cudaMemsetAsync(h->statN, 0, nb, h->stream));
    return NBMF_OK;
}

extern "C" int nbmf_init_from_labels(nbmf_handle *h, const int32_t *Z, int Kc)
{
    if (!h) return NBMF_ERR_ARG;
    if (!Z) return fail(h, NBMF_ERR_ARG, "Z is NULL");
    h->Kc = 0; // force re-allocation / zeroing
    int rc = ensure_model(h, Kc);
    if (rc) return rc;
    const int64_t F = h->F, N = h->N;
    // keep the labels on the device when they fit comfortably (needed by the exact
    // CVB0 wavefront and by the Gibbs entry point)
    cudaFree(h->Z); h->Z = nullptr;
    bool keep = (double)F * N * 4 <= 8e9;
    if (keep) {
        if (cudaMalloc(&h->Z, (size_t)F * N * 4) != cudaSuccess) { cudaGetLastError(); h->Z = nullptr; keep = false; }
    }
    int64_t cols_per = std::min<int64_t>(N, std::max<int64_t>(1, (int64_t)(256ll << 20) / (F * 4)));
    rc = ensure_stage(h, (size_t)cols_per * F * 4 + 64);
    if (rc) return rc;
    unsigned long long *cnt = (unsigned long long *)h->scal + SC_TMP;
    CK(cudaMemsetAsync(cnt, 0, 16, h->stream));
    for (int64_t n0 = 0; n0 < N; n0 += cols_per) {
        int64_t nc = std::min(cols_per, N - n0);
        int32_t *dst = keep ? h->Z + n0 * F : (int32_t *)h->stage;
        CK(cudaMemcpyAsync(dst, Z + n0 * F, (size_t)nc * F * 4, cudaMemcpyHostToDevice, h->stream));
        hist_labels_kernel<<<nblk(nc * F, 256), 256, 0, h->stream>>>(
            dst, h->vc, h->mc, F, nc, n0, h->Fw, Kc, h->Kp, h->statF, h->statN, cnt + 1);
        LAUNCH_CHECK();
        CK(cudaStreamSynchronize(h->stream));
    }
    unsigned long long bad = 0;
    CK(cudaMemcpy(&bad, cnt + 1, 8, cudaMemcpyDeviceToHost));
    if (bad) return fail(h, NBMF_ERR_ARG, "Z contains labels outside 0..Kc-1 (the reference would write out of bounds)");
    rc = exchange(h, h->statF, (size_t)F * h->Kp);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return NBMF_OK;
}

extern "C" int nbmf_init_random_labels(nbmf_handle *h, int Kc, uint64_t seed)
{
    if (!h) return NBMF_ERR_ARG;
    h->Kc = 0;
    int rc = ensure_model(h, Kc);
    if (rc) return rc;
    const int64_t F = h->F, N = h->N;
    cudaFree(h->Z); h->Z = nullptr;
    if ((double)F * N * 4 <= 8e9) {
        if (cudaMalloc(&h->Z, (size_t)F * N * 4) != cudaSuccess) { cudaGetLastError(); h->Z = nullptr; }
    }
    hist_random_cols_kernel<<<(unsigned)N, 256, 2 * Kc * sizeof(int), h->stream>>>(
        h->vc, h->mc, F, h->Fw, h->opts.n_offset, Kc, h->Kp, seed, h->statN, h->Z);
    LAUNCH_CHECK();
    hist_random_rows_kernel<<<(unsigned)F, 256, Kc * sizeof(int), h->stream>>>(
        h->mr, N, h->Nw, h->opts.n_offset, Kc, h->Kp, seed, h->statF);
    LAUNCH_CHECK();
    rc = exchange(h, h->statF, (size_t)F * h->Kp);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return NBMF_OK;
}

static int upload_rows(nbmf_handle *h, const double *src, int64_t rows, int K, double *dst, int mul,
                       int add)
{
    int rc = ensure_stage(h, (size_t)rows * K * 8);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h->stage, src, (size_t)rows * K * 8, cudaMemcpyHostToDevice, h->stream));
    colmajor_to_rows_kernel<<<nblk(rows * K, 256), 256, 0, h->stream>>>((const double *)h->stage, rows,
                                                                        K, dst, h->Kp, mul, add);
    LAUNCH_CHECK();
    CK(cudaStreamSynchronize(h->stream));
    return NBMF_OK;
}
static int download_rows(nbmf_handle *h, const double *src, int64_t rows, int K, double *dst, int mul,
                         int add)
{
    int rc = ensure_stage(h, (size_t)rows * K * 8);
    if (rc) return rc;
    rows_to_colmajor_kernel<<<nblk(rows * K, 256), 256, 0, h->stream>>>(src, rows, K, (double *)h->stage,
                                                                        h->Kp, mul, add);
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(dst, h->stage, (size_t)rows * K * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return NBMF_OK;
}

extern "C" int nbmf_set_stats(nbmf_handle *h, int Kc, const double *n_total, const double *f_ones,
                              const double *f_zeros)
{
    if (!h) return NBMF_ERR_ARG;
    if (!n_total || !f_ones || !f_zeros) return fail(h, NBMF_ERR_ARG, "NULL statistics");
    h->Kc = 0;
    int rc = ensure_model(h, Kc);
    if (rc) return rc;
    if ((rc = upload_rows(h, n_total, h->F, Kc, h->statF, 1, 0))) return rc;
    if ((rc = upload_rows(h, f_ones, h->N, Kc, h->statN, 2, 1))) return rc;
    if ((rc = upload_rows(h, f_zeros, h->N, Kc, h->statN, 2, 0))) return rc;
    return NBMF_OK;
}

extern "C" int nbmf_get_stats(nbmf_handle *h, int Kc, double *n_total, double *f_ones, double *f_zeros)
{
    if (!h) return NBMF_ERR_ARG;
    if (!h->statF || Kc != h->Kc) return fail(h, NBMF_ERR_STATE, "no statistics with this Kc");
    CK(cudaSetDevice(h->device));
    int rc;
    if (n_total && (rc = download_rows(h, h->statF, h->F, Kc, n_total, 1, 0))) return rc;
    if (f_ones && (rc = download_rows(h, h->statN, h->N, Kc, f_ones, 2, 1))) return rc;
    if (f_zeros && (rc = download_rows(h, h->statN, h->N, Kc, f_zeros, 2, 0))) return rc;
    return NBMF_OK;
}

// ---------------------------------------------------------------------------
// contraction passes
// ---------------------------------------------------------------------------
static bool use_tensor_path(const nbmf_handle *h)
{
    int p = h->opts.precision;
    if (p == NBMF_PREC_FP64) return false;
    if (p == NBMF_PREC_TENSOR || p == NBMF_PREC_TENSOR_FAST) return true;
    return tp_supported(h); // AUTO
}

template <int KC>
static cudaError_t launch_pass(const PassArgs &a, dim3 grid, size_t smem, cudaStream_t s)
{
    cudaError_t e = cudaFuncSetAttribute(pass_fp64_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    pass_fp64_kernel<KC><<<grid, 256, smem, s>>>(a);
    return cudaGetLastError();
}

// owner_rows != 0: owner = rows f, output GF; else owner = (n,b) lanes, output GN
static int run_pass_fp64(nbmf_handle *h, int owner_rows, double *sumlogD)
{
    PassArgs a{};
    a.Kp = h->Kp;
    a.sumlogD = sumlogD;
    if (owner_rows) {
        a.U = h->A; a.W = h->BB; a.G = h->GF;
        a.vbits = h->vr; a.mbits = h->has_na ? h->mr : nullptr;
        a.words = h->Nw; a.L = h->F; a.S = h->N; a.ob = 1; a.sb = 2;
    } else {
        a.U = h->BB; a.W = h->A; a.G = h->GN;
        a.vbits = h->vc; a.mbits = h->has_na ? h->mc : nullptr;
        a.words = h->Fw; a.L = 2 * h->N; a.S = h->F; a.ob = 2; a.sb = 1;
    }
    int TR = 32;
    while ((size_t)TR * a.sb * a.Kp * 8 > 65536 && TR > 1) TR >>= 1;
    a.TR = TR;
    size_t smem = (size_t)TR * a.sb * a.Kp * 8;
    const int OL = 256 / h->NKC;
    int64_t bx = (a.L + OL - 1) / OL;
    int64_t want = 4ll * h->sm_count;
    int64_t splits = std::max<int64_t>(1, std::min<int64_t>((want + bx - 1) / bx, (a.S + 255) / 256));
    a.s_per_split = round_up64((a.S + splits - 1) / splits, 32);
    splits = (a.S + a.s_per_split - 1) / a.s_per_split;
    if (splits > 1) CK(cudaMemsetAsync(a.G, 0, (size_t)a.L * a.Kp * 8, h->stream));
    dim3 grid((unsigned)bx, (unsigned)splits);
    cudaError_t e;
    if (h->KC == 4) e = launch_pass<4>(a, grid, smem, h->stream);
    else if (h->KC == 8) e = launch_pass<8>(a, grid, smem, h->stream);
    else e = launch_pass<16>(a, grid, smem, h->stream);
    CK(e);
    h->timing.launches += 1 + (splits > 1);
    return NBMF_OK;
}

static int run_pass(nbmf_handle *h, int owner_rows, double *sumlogD)
{
    if (use_tensor_path(h)) return tp_pass(h, owner_rows, sumlogD);
    return run_pass_fp64(h, owner_rows, sumlogD);
}

static int run_params(nbmf_handle *h, int K, double alpha, double beta, double gamma, int mean_mode,
                      int want_lb)
{
    params_rows_kernel<<<nblk(h->F * 32, 256), 256, 0, h->stream>>>(h->statF, h->F, K, h->Kp, gamma,
                                                                    mean_mode, want_lb, h->A, h->EW, h->scal);
    LAUNCH_CHECK();
    params_cols_kernel<<<nblk(h->N * 32, 256), 256, 0, h->stream>>>(h->statN, h->N, K, h->Kp, alpha, beta,
                                                                    mean_mode, want_lb, h->BB, h->EH, h->scal);
    LAUNCH_CHECK();
    h->timing.launches += 2;
    if (use_tensor_path(h)) return tp_convert_operands(h);
    return NBMF_OK;
}

// ---------------------------------------------------------------------------
// VB (begin / step / end)
// ---------------------------------------------------------------------------
extern "C" int nbmf_vb_begin(nbmf_handle *h, int K, double alpha, double beta, double gamma,
                             int mean_params)
{
    if (!h) return NBMF_ERR_ARG;
    if (!h->vc) return fail(h, NBMF_ERR_STATE, "no data");
    if (!h->statF || h->Kc != K)
        return fail(h, NBMF_ERR_STATE, "statistics missing or initialised with a different K");
    if (!(alpha > 0) || !(beta > 0) || !(gamma > 0)) return fail(h, NBMF_ERR_ARG, "priors must be > 0");
    CK(cudaSetDevice(h->device));
    if (use_tensor_path(h)) {
        int rc = tp_prepare(h);
        if (rc) return rc;
    }
    h->vb_active = true; h->vb_K = K; h->vb_iter = 0; h->vb_mean_params = mean_params;
    h->vb_alpha = alpha; h->vb_beta = beta; h->vb_gamma = gamma;
    h->timing = nbmf_timing{};
    CK(cudaMemsetAsync(h->scal, 0, SC_COUNT * sizeof(double), h->stream));
    CK(cudaEventRecord(h->ev[0], h->stream));
    return NBMF_OK;
}

extern "C" int nbmf_vb_step(nbmf_handle *h, int checklb)
{
    if (!h) return NBMF_ERR_ARG;
    if (!h->vb_active) return fail(h, NBMF_ERR_STATE, "nbmf_vb_begin not called");
    CK(cudaSetDevice(h->device));
    const int i = h->vb_iter;
    if (i >= h->lbs_cap) { // trace: 3 doubles per iteration (sum log D, row prior term, column prior term)
        int cap = std::max(256, h->lbs_cap * 2);
        double *nl = nullptr;
        CK(cudaMalloc(&nl, (size_t)cap * 24));
        CK(cudaMemsetAsync(nl, 0, (size_t)cap * 24, h->stream));
        if (h->lbs) CK(cudaMemcpyAsync(nl, h->lbs, (size_t)h->lbs_cap * 24, cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        cudaFree(h->lbs);
        h->lbs = nl; h->lbs_cap = cap;
    }
    int rc;
    if (checklb) CK(cudaMemsetAsync(h->scal, 0, 3 * sizeof(double), h->stream)); // SUMLOGD, PRIOR_ROWS, PRIOR_COLS
    CK(cudaEventRecord(h->ev[1], h->stream));
    if ((rc = run_params(h, h->vb_K, h->vb_alpha, h->vb_beta, h->vb_gamma, h->vb_mean_params, checklb))) return rc;
    CK(cudaEventRecord(h->ev[2], h->stream));
    if (h->pipe_blocks > 0) {
        // first step of nbmf_vb_aspect_bits: both passes, column block by column block, as the
        // upload of V arrives; reductions, exchange and corrections follow once
        const int nb = h->pipe_blocks;
        h->pipe_blocks = 0;
        if ((rc = tp_passes_pipelined(h, nb, h->pipe_n0, h->pipe_ev, build_row_planes_block,
                                      checklb ? h->scal + SC_SUMLOGD : nullptr))) return rc;
        if (checklb && (rc = tp_log_correction(h, 1, h->scal + SC_SUMLOGD))) return rc;
        CK(cudaEventRecord(h->ev[3], h->stream));
        if ((rc = exchange(h, h->GF, (size_t)h->F * h->Kp))) return rc;
        CK(cudaEventRecord(h->ev[4], h->stream));
        if (checklb && (rc = tp_log_correction(h, 0, h->scal + SC_SUMLOGD))) return rc;
    } else {
    // Dirichlet-side statistic: owner = rows, reduce over the local columns, then
    // the one exchange step, then n_total = A (.) G_F
    if ((rc = run_pass(h, 1, checklb ? h->scal + SC_SUMLOGD : nullptr))) return rc;
    if (checklb && use_tensor_path(h) && (rc = tp_log_correction(h, 1, h->scal + SC_SUMLOGD))) return rc;
    CK(cudaEventRecord(h->ev[3], h->stream));
    // the cols pass needs nothing from the exchange (both statistics follow from the parameters),
    // so on the tensor path the exchange runs beside it and is joined before the finalisation
    if (use_tensor_path(h)) rc = exchange_begin(h, h->GF, (size_t)h->F * h->Kp);
    else rc = exchange(h, h->GF, (size_t)h->F * h->Kp);
    if (rc) return rc;
    CK(cudaEventRecord(h->ev[4], h->stream));
    // Beta-side statistics: owner = (column, branch) lanes, all local
    if ((rc = run_pass(h, 0, nullptr))) return rc;
    if (checklb && use_tensor_path(h) && (rc = tp_log_correction(h, 0, h->scal + SC_SUMLOGD))) return rc;
    if ((rc = exchange_end(h))) return rc;
    }
    if (use_tensor_path(h)) {
        if ((rc = tp_finalize(h))) return rc;
    } else {
        finalize_stats_kernel<<<nblk(h->F * h->Kp, 256), 256, 0, h->stream>>>(h->A, h->GF, h->F * h->Kp, h->statF);
        LAUNCH_CHECK();
        finalize_stats_kernel<<<nblk(2 * h->N * h->Kp, 256), 256, 0, h->stream>>>(h->BB, h->GN, 2 * h->N * h->Kp, h->statN);
        LAUNCH_CHECK();
        h->timing.launches += 2;
    }
    // lb_i = sum log D + (pW - qW) + (pH - qH).  The first and third are local to the
    // rank's columns, the second is replicated (nbmf_vb_get_lb_terms for sharded runs).
    CK(cudaMemcpyAsync(h->lbs + 3 * (size_t)i, h->scal + SC_SUMLOGD, 24, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaEventRecord(h->ev[5], h->stream));
    h->vb_iter++;
    h->timing.iters = h->vb_iter;
    return NBMF_OK;
}

extern "C" int nbmf_vb_end(nbmf_handle *h, double *E_W, double *E_H, double *lowerbounds, int n_lb,
                           double *E_Z)
{
    if (!h) return NBMF_ERR_ARG;
    if (!h->vb_active) return fail(h, NBMF_ERR_STATE, "nbmf_vb_begin not called");
    CK(cudaSetDevice(h->device));
    CK(cudaEventRecord(h->ev[6], h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->vb_active = false;
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[6]); h->timing.total_ms = ms;
    if (h->vb_iter > 0) {
        cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]); h->timing.params_ms = ms;
        cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]); h->timing.pass_rows_ms = ms;
        cudaEventElapsedTime(&ms, h->ev[3], h->ev[4]); h->timing.exchange_ms = ms;
        cudaEventElapsedTime(&ms, h->ev[4], h->ev[5]); h->timing.pass_cols_ms = ms;
    }
    int rc;
    if (use_tensor_path(h) && (rc = tp_check_error(h))) return rc;
    double neg = 0;
    CK(cudaMemcpy(&neg, h->scal + SC_NEG, 8, cudaMemcpyDeviceToHost));
    if (neg != 0.0) return fail(h, NBMF_ERR_NEGATIVE, "Negative values in alpha_vb / beta_vb / gamma_vb");
    const int K = h->vb_K;
    if (h->vb_iter == 0) { // E_W / E_H of the current statistics
        if ((rc = run_params(h, K, h->vb_alpha, h->vb_beta, h->vb_gamma, h->vb_mean_params, 0))) return rc;
    }
    if (E_W && (rc = download_rows(h, h->EW, h->F, K, E_W, 1, 0))) return rc;
    if (E_H && (rc = download_rows(h, h->EH, h->N, K, E_H, 1, 0))) return rc;
    if (lowerbounds && n_lb > 0) {
        int n = std::min(n_lb, h->vb_iter);
        std::vector<double> a((size_t)std::max(1, n) * 3, 0.0);
        if (n > 0) CK(cudaMemcpy(a.data(), h->lbs, (size_t)n * 24, cudaMemcpyDeviceToHost));
        for (int i = 0; i < n_lb; i++) lowerbounds[i] = (i < n) ? a[3 * i] + a[3 * i + 1] + a[3 * i + 2] : 0.0;
    }
    if (E_Z) {
        const int64_t cnt = h->F * h->N * (int64_t)K;
        if ((double)cnt * 8 > 16e9) return fail(h, NBMF_ERR_UNSUPPORTED, "E_Z too large to return; pass NULL");
        double *dz = nullptr;
        if (cudaMalloc(&dz, (size_t)cnt * 8) != cudaSuccess) { cudaGetLastError(); return fail(h, NBMF_ERR_OOM, "E_Z allocation failed"); }
        ez_aspect_kernel<<<nblk(h->F * h->N, 256), 256, 0, h->stream>>>(h->A, h->BB, h->vc, h->has_na ? h->mc : nullptr,
                                                                       h->F, h->N, h->Fw, K, h->Kp, dz);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(E_Z, dz, (size_t)cnt * 8, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        cudaFree(dz);
        CK(e);
    }
    return NBMF_OK;
}

extern "C" int nbmf_vb_get_lb_terms(nbmf_handle *h, double *terms3, int n_iter)
{
    if (!h || !terms3) return NBMF_ERR_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    int n = std::min(n_iter, std::min(h->vb_iter, h->lbs_cap));
    for (int i = 0; i < 3 * n_iter; i++) terms3[i] = 0.0;
    if (n > 0) CK(cudaMemcpy(terms3, h->lbs, (size_t)n * 24, cudaMemcpyDeviceToHost));
    return NBMF_OK;
}

extern "C" int nbmf_vb_aspect(nbmf_handle *h, int K, double alpha, double beta, double gamma, int iter,
                              int checklb, double *E_W, double *E_H, double *lowerbounds, double *E_Z)
{
    if (!h) return NBMF_ERR_ARG;
    if (iter < 0) return fail(h, NBMF_ERR_ARG, "iter < 0");
    int rc = nbmf_vb_begin(h, K, alpha, beta, gamma, 0);
    if (rc) return rc;
    for (int i = 0; i < iter; i++)
        if ((rc = nbmf_vb_step(h, checklb))) { h->vb_active = false; return rc; }
    return nbmf_vb_end(h, E_W, E_H, lowerbounds, iter, E_Z);
}

// VB_Aspect_Rcpp with V given as host bit planes and the initial statistics as host arrays: one
// call = upload + iter iterations + results.  The upload is cut into column blocks on a copy
// stream and the first iteration consumes the blocks as they land (both passes of a block need
// nothing but that block's columns), so the PCIe transfer overlaps the contraction.
extern "C" int nbmf_vb_aspect_bits(nbmf_handle *h, const uint32_t *vbits, int64_t F, int64_t N, int K,
                                   double alpha, double beta, double gamma, int iter, int checklb,
                                   const double *n_total, const double *f_ones, const double *f_zeros,
                                   double *E_W, double *E_H, double *lowerbounds)
{
    if (!h) return NBMF_ERR_ARG;
    if (!vbits || F <= 0 || N <= 0 || iter < 0) return fail(h, NBMF_ERR_ARG, "nbmf_vb_aspect_bits: bad arguments");
    CK(cudaSetDevice(h->device));
    // block size: about N/8 columns (a multiple of 64 = one owner tile of the cols pass = two
    // steps of the rows pass); at most 32 blocks
    int64_t blk = h->opts.pipe_block_cols > 0 ? (int64_t)h->opts.pipe_block_cols
                                              : std::max<int64_t>(2048, ((N + 7) / 8 + 2047) / 2048 * 2048);
    blk = std::max<int64_t>(64, (blk + 63) / 64 * 64);
    while ((N + blk - 1) / blk > 32) blk *= 2;
    const int nb = (int)((N + blk - 1) / blk);
    int rc = alloc_planes(h, F, N);
    if (rc) return rc;
    h->has_na = false; h->nobs = F * N;
    const bool pipelined = iter >= 1 && nb >= 2 && use_tensor_path(h);
    if (!pipelined) {
        if ((rc = nbmf_set_data_bits(h, vbits, nullptr, F, N))) return rc;
        if ((rc = nbmf_set_stats(h, K, n_total, f_ones, f_zeros))) return rc;
        return nbmf_vb_aspect(h, K, alpha, beta, gamma, iter, checklb, E_W, E_H, lowerbounds, nullptr);
    }
    tp_invalidate_data(h);
    if (!h->copy_stream) CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int b = 0; b <= nb; b++) h->pipe_n0[b] = std::min<int64_t>(N, (int64_t)b * blk);
    // the statistics first (the parameters of the first iteration need all of them), then V
    if ((rc = nbmf_set_stats(h, K, n_total, f_ones, f_zeros))) return rc;   // synchronises h->stream
    const int64_t fw_in = (F + 31) / 32;
    for (int b = 0; b < nb; b++) {
        const int64_t c0 = h->pipe_n0[b], c1 = h->pipe_n0[b + 1];
        if (!h->pipe_ev[b]) CK(cudaEventCreateWithFlags(&h->pipe_ev[b], cudaEventDisableTiming));
        CK(cudaMemcpy2DAsync(h->vc + c0 * h->Fw, (size_t)h->Fw * 4, vbits + c0 * fw_in, (size_t)fw_in * 4,
                             (size_t)fw_in * 4, (size_t)(c1 - c0), cudaMemcpyHostToDevice, h->copy_stream));
        CK(cudaEventRecord(h->pipe_ev[b], h->copy_stream));
    }
    fill_mask_kernel<<<nblk(N * h->Fw, 256), 256, 0, h->stream>>>(h->mc, N, h->Fw, F);
    LAUNCH_CHECK();
    fill_mask_kernel<<<nblk(F * h->Nw, 256), 256, 0, h->stream>>>(h->mr, F, h->Nw, N);
    LAUNCH_CHECK();
    if ((rc = nbmf_vb_begin(h, K, alpha, beta, gamma, 0))) { cudaStreamSynchronize(h->copy_stream); return rc; }
    h->pipe_blocks = nb;
    for (int i = 0; i < iter; i++)
        if ((rc = nbmf_vb_step(h, checklb))) {
            h->vb_active = false; h->pipe_blocks = 0;
            cudaStreamSynchronize(h->copy_stream);   // the host buffer must not be in use after the return
            return rc;
        }
    rc = nbmf_vb_end(h, E_W, E_H, lowerbounds, iter, nullptr);
    cudaStreamSynchronize(h->copy_stream);
    return rc;
}

extern "C" int nbmf_vb_argmax_labels(nbmf_handle *h, int32_t *Z_out)
{
    if (!h) return NBMF_ERR_ARG;
    if (!h->A || !h->BB || h->Kc <= 0) return fail(h, NBMF_ERR_STATE, "no VB operands: run nbmf_vb_aspect / nbmf_vb_step first");
    if (h->hook) return fail(h, NBMF_ERR_UNSUPPORTED, "argmax labels under column sharding: call per rank on its own block");
    CK(cudaSetDevice(h->device));
    const int64_t F = h->F, N = h->N;
    if ((double)F * N * 4 > 64e9) return fail(h, NBMF_ERR_UNSUPPORTED, "label matrix too large");
    cudaFree(h->Z); h->Z = nullptr;
    if (cudaMalloc(&h->Z, (size_t)F * N * 4) != cudaSuccess) { cudaGetLastError(); return fail(h, NBMF_ERR_OOM, "label allocation failed"); }
    argmax_labels_kernel<<<nblk(F * N, 256), 256, 0, h->stream>>>(h->A, h->BB, h->vc, h->mc, F, N, h->Fw, h->Kc, h->Kp, h->Z);
    LAUNCH_CHECK();
    if (Z_out) CK(cudaMemcpyAsync(Z_out, h->Z, (size_t)F * N * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return NBMF_OK;
}

extern "C" int nbmf_lower_bound(nbmf_handle *h, int K, double alpha, double beta, double gamma,
                                double *lb, double *terms7)
{
    if (!h) return NBMF_ERR_ARG;
    if (!h->statF || h->Kc != K) return fail(h, NBMF_ERR_STATE, "statistics missing");
    CK(cudaSetDevice(h->device));
    if (use_tensor_path(h)) { int rc0 = tp_prepare(h); if (rc0) return rc0; }
    CK(cudaMemsetAsync(h->scal, 0, 3 * sizeof(double), h->stream));
    int rc;
    if ((rc = run_params(h, K, alpha, beta, gamma, 0, 1))) return rc;
    if ((rc = run_pass(h, 1, h->scal + SC_SUMLOGD))) return rc;
    if (use_tensor_path(h)) {
        if ((rc = tp_log_correction(h, 1, h->scal + SC_SUMLOGD))) return rc;
        if ((rc = run_pass(h, 0, nullptr))) return rc;
        if ((rc = tp_log_correction(h, 0, h->scal + SC_SUMLOGD))) return rc;
    }
    double s[3];
    CK(cudaMemcpyAsync(s, h->scal, 24, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (use_tensor_path(h) && (rc = tp_check_error(h))) return rc;
    if (lb) *lb = s[0] + s[1] + s[2];
    if (terms7) { // pW-qW, pZ+pV-qZ (= sum log D), pH-qH, rest 0
        for (int i = 0; i < 7; i++) terms7[i] = 0;
        terms7[0] = s[1]; terms7[1] = s[0]; terms7[2] = s[2];
    }
    return NBMF_OK;
}

// ---------------------------------------------------------------------------
// VB_DirBer_Rcpp
// ---------------------------------------------------------------------------
// Launch a wavefront kernel with G lanes per entry (G = smallest power of two >= Kc in [4, 32], halved
// while that lets a diagonal fit one 1024-thread pass): one block
// while a diagonal needs at most four passes of 1024 threads (block barrier), else a cooperative
// grid (grid barrier between diagonals).
template <template <bool, int> class KERN, typename ARGS>
static cudaError_t launch_wavefront(nbmf_handle *h, ARGS &a, int Kc, int64_t diag)
{
    int G = 4;
    while (G < Kc && G < 32) G <<= 1;
    // a short diagonal is better served by one pass of one block than by more lanes per entry
    while (G > 4 && diag * G > 1024 && diag * (G / 2) <= 1024) G >>= 1;
    const int64_t need = diag * G;
    const bool single = need <= 4096;
    const int threads = single ? (int)std::min<int64_t>(1024, std::max<int64_t>(32, round_up64(need, 32))) : 256;
    void *args[] = {&a};
    auto go = [&](auto gtag) -> cudaError_t {
        constexpr int GG = decltype(gtag)::value;
        if (single) {
            KERN<true, GG>::launch(a, 1, threads, h->stream);
            return cudaGetLastError();
        }
        int per_sm = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, KERN<false, GG>::ptr(), threads, 0);
        if (e != cudaSuccess) return e;
        const int blocks = (int)std::min<int64_t>((need + threads - 1) / threads, (int64_t)per_sm * h->sm_count);
        return cudaLaunchCooperativeKernel(KERN<false, GG>::ptr(), dim3(blocks), dim3(threads), args, 0, h->stream);
    };
    switch (G) {
    case 4: return go(std::integral_constant<int, 4>{});
    case 8: return go(std::integral_constant<int, 8>{});
    case 16: return go(std::integral_constant<int, 16>{});
    default: return go(std::integral_constant<int, 32>{});
    }
}
template <bool SB, int G> struct Cvb0Kern {
    static void *ptr() { return (void *)cvb0_wavefront_kernel<SB, G>; }
    static void launch(WaveArgs &a, int blocks, int threads, cudaStream_t st) { cvb0_wavefront_kernel<SB, G><<<blocks, threads, 0, st>>>(a); }
};
template <int VARIANT> struct GibbsWave {
    template <bool SB, int G> struct Kern {
        static void *ptr() { return (void *)gibbs_collapsed_wavefront_kernel<SB, G, VARIANT>; }
        static void launch(GibbsWaveArgs &a, int blocks, int threads, cudaStream_t st) { gibbs_collapsed_wavefront_kernel<SB, G, VARIANT><<<blocks, threads, 0, st>>>(a); }
    };
};

static int dirber_wavefront(nbmf_handle *h, int K, double alpha, double beta, double gamma, int iter,
                            double *E_W, double *E_H, double *E_Z)
{
    const int Kc = K + 1;
    const int64_t F = h->F, N = h->N;
    if (!h->Z) return fail(h, NBMF_ERR_STATE, "exact wavefront mode needs labels: call nbmf_init_from_labels(Z, K+1)");
    if (h->Kc != Kc) return fail(h, NBMF_ERR_STATE, "statistics must be initialised with Kc = K+1 (collapsed_gibbs_z_VB.cpp:101)");
    if (h->hook) return fail(h, NBMF_ERR_UNSUPPORTED, "exact wavefront mode is replicas-only (no column sharding)");
    if ((double)F * N * Kc * 8 > 64e9) return fail(h, NBMF_ERR_UNSUPPORTED, "E_Z does not fit: use NBMF_DIRBER_CONTRACTION");
    double *EZ = nullptr, *buf = nullptr;
    size_t fk = (size_t)F * Kc, nk = (size_t)N * Kc;
    if (cudaMalloc(&EZ, (size_t)F * N * Kc * 8) != cudaSuccess || cudaMalloc(&buf, (2 * fk + 4 * nk) * 8) != cudaSuccess) {
        cudaGetLastError(); cudaFree(EZ);
        return fail(h, NBMF_ERR_OOM, "wavefront buffers allocation failed");
    }
    WaveArgs a{};
    a.vc = h->vc; a.mc = h->has_na ? h->mc : nullptr; a.F = F; a.N = N; a.Fw = h->Fw; a.Kc = Kc;
    a.sweeps = iter > 1 ? iter - 1 : 0;
    a.alpha = alpha; a.beta = beta; a.gamma = gamma;
    a.EZ = EZ; a.cF = buf; a.cN1 = buf + fk; a.cN0 = a.cN1 + nk; a.gvb = a.cN0 + nk; a.avb = a.gvb + fk; a.bvb = a.avb + nk;
    int rc = NBMF_OK;
    cudaError_t e = cudaMemsetAsync(buf, 0, (2 * fk + 4 * nk) * 8, h->stream);
    // counters from the statistics (one-hot E_Z + compute_vbcounts)
    if (e == cudaSuccess) e = cudaMemcpy2DAsync(a.cF, Kc * 8, h->statF, h->Kp * 8, Kc * 8, F, cudaMemcpyDeviceToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpy2DAsync(a.cN1, Kc * 8, h->statN + h->Kp, 2 * h->Kp * 8, Kc * 8, N, cudaMemcpyDeviceToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpy2DAsync(a.cN0, Kc * 8, h->statN, 2 * h->Kp * 8, Kc * 8, N, cudaMemcpyDeviceToDevice, h->stream);
    if (e == cudaSuccess) {
        onehot_ez_kernel<<<nblk(F * N, 256), 256, 0, h->stream>>>(h->Z, a.mc, F, N, h->Fw, Kc, EZ);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && a.sweeps > 0) {
        e = launch_wavefront<Cvb0Kern>(h, a, Kc, std::min(F, N));
        h->timing.launches += 1;
    }
    double *outW = nullptr;
    if (e == cudaSuccess) e = cudaMalloc(&outW, (fk + nk) * 8);
    if (e == cudaSuccess) {
        cudaMemsetAsync(h->scal + SC_NEG, 0, 8, h->stream);
        dirber_outputs_kernel<<<nblk(std::max(F, N), 256), 256, 0, h->stream>>>(a.gvb, a.avb, a.bvb, F, N, Kc, outW, outW + fk, h->scal);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && E_W) e = cudaMemcpyAsync(E_W, outW, fk * 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess && E_H) e = cudaMemcpyAsync(E_H, outW + fk, nk * 8, cudaMemcpyDeviceToHost, h->stream);
    double neg = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&neg, h->scal + SC_NEG, 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess && E_Z) {
        double *cube = nullptr;
        e = cudaMalloc(&cube, (size_t)F * N * Kc * 8);
        if (e == cudaSuccess) {
            ez_to_cube_kernel<<<nblk(F * N * Kc, 256), 256, 0, h->stream>>>(EZ, F, N, Kc, cube);
            e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaMemcpyAsync(E_Z, cube, (size_t)F * N * Kc * 8, cudaMemcpyDeviceToHost, h->stream);
            cudaStreamSynchronize(h->stream);
            cudaFree(cube);
        }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(EZ); cudaFree(buf); cudaFree(outW);
    if (e != cudaSuccess) { h->err = std::string("wavefront: ") + cudaGetErrorString(e); cudaGetLastError(); return NBMF_ERR_CUDA; }
    if (neg != 0.0) return fail(h, NBMF_ERR_NEGATIVE, "Negative values in alpha_vb / beta_vb");
    return rc;
}

extern "C" int nbmf_vb_dirber(nbmf_handle *h, int K, double alpha, double beta, double gamma, int iter,
                              int mode, double *E_W, double *E_H, double *E_Z)
{
    if (!h) return NBMF_ERR_ARG;
    if (!h->vc) return fail(h, NBMF_ERR_STATE, "no data");
    if (!(alpha > 0) || !(beta > 0) || !(gamma > 0)) return fail(h, NBMF_ERR_ARG, "priors must be > 0");
    CK(cudaSetDevice(h->device));
    if (mode == NBMF_DIRBER_EXACT_WAVEFRONT) {
        cudaEventRecord(h->ev[0], h->stream);
        h->timing = nbmf_timing{};
        int rc = dirber_wavefront(h, K, alpha, beta, gamma, iter, E_W, E_H, E_Z);
        cudaEventRecord(h->ev[6], h->stream);
        cudaStreamSynchronize(h->stream);
        float ms = 0; cudaEventElapsedTime(&ms, h->ev[0], h->ev[6]);
        h->timing.total_ms = ms; h->timing.iters = iter > 1 ? iter - 1 : 0;
        return rc;
    }
    // contraction form: same Kc = K+1 columns, iter-1 batch sweeps, mean parameters
    const int Kc = K + 1;
    if (h->Kc != Kc) return fail(h, NBMF_ERR_STATE, "statistics must be initialised with Kc = K+1");
    int rc = nbmf_vb_begin(h, Kc, alpha, beta, gamma, 1);
    if (rc) return rc;
    for (int i = 1; i < iter; i++)
        if ((rc = nbmf_vb_step(h, 0))) { h->vb_active = false; return rc; }
    if (E_Z) return fail(h, NBMF_ERR_UNSUPPORTED, "E_Z is only returned by the exact wavefront mode");
    return nbmf_vb_end(h, E_W, E_H, nullptr, 0, nullptr);
}

// ---------------------------------------------------------------------------
// Gibbs
// ---------------------------------------------------------------------------
extern "C" int nbmf_gibbs_dirber(nbmf_handle *h, int K, double alpha, double beta, double gamma, int iter,
                                 double burnin, uint64_t seed, double *E_W, double *E_H,
                                 const int64_t *test_idx, int64_t n_test, double *E_V_test, int32_t *Z_last)
{
    if (!h) return NBMF_ERR_ARG;
    if (!h->vc) return fail(h, NBMF_ERR_STATE, "no data");
    if (K <= 0 || K > 256) return fail(h, NBMF_ERR_ARG, "Gibbs: K must be in 1..256");
    if (!(alpha > 0) || !(beta > 0) || !(gamma > 0)) return fail(h, NBMF_ERR_ARG, "priors must be > 0");
    CK(cudaSetDevice(h->device));
    const int64_t F = h->F, N = h->N;
    const size_t fk = (size_t)F * K, nk = (size_t)N * K;
    int32_t *cF = nullptr, *cN = nullptr, *Zd = nullptr;
    float *W = nullptr, *H = nullptr, *wc = nullptr, *hc = nullptr;
    double *accW = nullptr, *accH = nullptr, *accV = nullptr, *outb = nullptr;
    int64_t *didx = nullptr;
    cudaError_t e = cudaSuccess;
    auto A1 = [&](void **p, size_t b) { if (e == cudaSuccess) e = cudaMalloc(p, b ? b : 8); };
    A1((void **)&cF, fk * 4); A1((void **)&cN, 2 * nk * 4);
    A1((void **)&W, fk * 4); A1((void **)&H, nk * 4); A1((void **)&wc, fk * 4); A1((void **)&hc, nk * 4);
    A1((void **)&accW, fk * 8); A1((void **)&accH, nk * 8); A1((void **)&outb, std::max(fk, nk) * 8);
    // column sharding: the row counts n_total (F x K) are the one thing exchanged per sweep; the
    // hook sums fp64, so the int32 counts take a round trip through `outb`
    auto exchange_counts = [&]() -> cudaError_t {
        if (!h->hook) return cudaSuccess;
        i32_to_f64_kernel<<<nblk((int64_t)fk, 256), 256, 0, h->stream>>>(cF, (int64_t)fk, outb);
        if (h->hook(h->hook_ctx, (void *)outb, fk, (void *)h->stream) != 0) return cudaErrorUnknown;
        f64_to_i32_kernel<<<nblk((int64_t)fk, 256), 256, 0, h->stream>>>(outb, (int64_t)fk, cF);
        return cudaGetLastError();
    };
    if (n_test > 0) { A1((void **)&accV, (size_t)n_test * 8); A1((void **)&didx, (size_t)n_test * 8); }
    if (Z_last) A1((void **)&Zd, (size_t)F * N * 4);
    int rc = NBMF_OK;
    int kept = 0;
    const int nburnin = (int)std::floor(iter * burnin);
    unsigned long long *cnt = (unsigned long long *)h->scal + SC_TMP;
    h->timing = nbmf_timing{};
    if (e == cudaSuccess) {
        cudaMemsetAsync(cF, 0, fk * 4, h->stream); cudaMemsetAsync(cN, 0, 2 * nk * 4, h->stream);
        cudaMemsetAsync(accW, 0, fk * 8, h->stream); cudaMemsetAsync(accH, 0, nk * 8, h->stream);
        cudaMemsetAsync(cnt, 0, 16, h->stream);
        if (n_test > 0) {
            cudaMemsetAsync(accV, 0, (size_t)n_test * 8, h->stream);
            e = cudaMemcpyAsync(didx, test_idx, (size_t)n_test * 8, cudaMemcpyHostToDevice, h->stream);
        }
    }
    // initial counts: from stored labels (compute_counts) when available, else from the
    // fp64 statistics (which hold integer counts after a label init)
    if (e == cudaSuccess) {
        if (h->Z) {
            counts_from_labels_kernel<<<nblk(F * N, 256), 256, 0, h->stream>>>(h->Z, h->vc, h->has_na ? h->mc : nullptr, F, N, h->Fw, K, cF, cN, cnt + 1);
            e = cudaGetLastError();
            unsigned long long bad = 0;
            if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, cnt + 1, 8, cudaMemcpyDeviceToHost, h->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
            if (e == cudaSuccess && bad) { rc = fail(h, NBMF_ERR_ARG, "Z contains labels outside 0..K-1 (BernoulliNMF.R:18-20 forces K_init=10)"); }
            if (e == cudaSuccess && rc == NBMF_OK) e = exchange_counts();
        } else if (h->statF && h->Kc == K) {
            // no label matrix on the device (too large): the statistics of a label initialisation
            // are the integer counts themselves
            f64_stats_to_counts_kernel<<<nblk((int64_t)std::max(fk, 2 * nk), 256), 256, 0, h->stream>>>(
                h->statF, h->statN, F, N, K, h->Kp, cF, cN);
            e = cudaGetLastError();
        } else {
            rc = fail(h, NBMF_ERR_STATE, "Gibbs needs labels: call nbmf_init_from_labels or nbmf_init_random_labels first");
        }
    }
    const size_t smem_sweep = (size_t)4 * (160 * (size_t)K + 32);
    if (e == cudaSuccess && rc == NBMF_OK)
        e = cudaFuncSetAttribute(gibbs_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sweep);
    if (e == cudaSuccess && rc == NBMF_OK) {
        cudaEventRecord(h->ev[0], h->stream);
        dim3 sgrid((unsigned)((F + 31) / 32), (unsigned)((N + 31) / 32));
        for (int i = 1; i < iter && e == cudaSuccess; i++) {
            gibbs_draw_W_kernel<<<nblk(F * 32, 256), 256, 0, h->stream>>>(cF, F, K, gamma, seed, (uint32_t)i, W);
            gibbs_draw_H_kernel<<<nblk(N * K, 256), 256, 0, h->stream>>>(cN, N, h->opts.n_offset, K, alpha, beta, seed, (uint32_t)i, H);
            cudaMemsetAsync(cF, 0, fk * 4, h->stream);
            cudaMemsetAsync(cN, 0, 2 * nk * 4, h->stream);
            gibbs_sweep_kernel<<<sgrid, 256, smem_sweep, h->stream>>>(h->vc, h->has_na ? h->mc : nullptr, F, N, h->Fw, h->opts.n_offset, K, W, H, seed, (uint32_t)i, cF, cN, (i == iter - 1) ? Zd : nullptr);
            h->timing.launches += 5;
            if (e == cudaSuccess) e = exchange_counts();
            if (i >= nburnin) {
                gibbs_accumulate_kernel<<<nblk(std::max(F, N), 256), 256, 0, h->stream>>>(cF, cN, F, N, K, alpha, beta, gamma, accW, accH, wc, hc);
                if (n_test > 0)
                    gibbs_ev_test_kernel<<<nblk(n_test, 256), 256, 0, h->stream>>>(didx, n_test, F, K, wc, hc, accV);
                h->timing.launches += 1 + (n_test > 0);
                kept++;
            }
            e = cudaGetLastError();
        }
        cudaEventRecord(h->ev[6], h->stream);
    }
    if (e == cudaSuccess && rc == NBMF_OK) {
        double sc = kept > 0 ? 1.0 / kept : 0.0;
        if (E_W) {
            scale_rows_to_colmajor_kernel<<<nblk(F * K, 256), 256, 0, h->stream>>>(accW, F, K, sc, outb);
            e = cudaMemcpyAsync(E_W, outb, fk * 8, cudaMemcpyDeviceToHost, h->stream);
            cudaStreamSynchronize(h->stream);
        }
        if (E_H && e == cudaSuccess) {
            scale_rows_to_colmajor_kernel<<<nblk(N * K, 256), 256, 0, h->stream>>>(accH, N, K, sc, outb);
            e = cudaMemcpyAsync(E_H, outb, nk * 8, cudaMemcpyDeviceToHost, h->stream);
            cudaStreamSynchronize(h->stream);
        }
        if (E_V_test && n_test > 0 && e == cudaSuccess) {
            e = cudaMemcpyAsync(E_V_test, accV, (size_t)n_test * 8, cudaMemcpyDeviceToHost, h->stream);
            cudaStreamSynchronize(h->stream);
            for (int64_t t = 0; t < n_test; t++) E_V_test[t] *= sc;
        }
        if (Z_last && e == cudaSuccess) {
            if (iter <= 1 && h->Z) e = cudaMemcpyAsync(Z_last, h->Z, (size_t)F * N * 4, cudaMemcpyDeviceToHost, h->stream);
            else e = cudaMemcpyAsync(Z_last, Zd, (size_t)F * N * 4, cudaMemcpyDeviceToHost, h->stream);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        float ms = 0; cudaEventElapsedTime(&ms, h->ev[0], h->ev[6]);
        h->timing.total_ms = ms; h->timing.iters = iter > 1 ? iter - 1 : 0;
    }
    cudaFree(cF); cudaFree(cN); cudaFree(W); cudaFree(H); cudaFree(wc); cudaFree(hc);
    cudaFree(accW); cudaFree(accH); cudaFree(accV); cudaFree(outb); cudaFree(didx); cudaFree(Zd);
    if (e != cudaSuccess) { h->err = std::string("gibbs: ") + cudaGetErrorString(e); cudaGetLastError(); return NBMF_ERR_CUDA; }
    return rc;
}

// The reference's collapsed samplers with their own schedule: the sequential sweep in anti-diagonal
// wavefront order, uniforms keyed (seed; f, n, sweep).  variant 0 = Gibbs_DirBer_finite_Rcpp
// (collapsed_gibbs_z.cpp:363-437), 1 = Gibbs_DirBer_DP_Rcpp (:275-359), 2 = Gibbs_DirDir_finite_Rcpp
// (:441-557).  Sized for the configurations the reference itself can run (labels and the sample
// cubes live on the device); replicas only.
static int collapsed_wavefront(nbmf_handle *h, int variant, int K, double alpha, double beta, double gamma,
                               int iter, double burnin, uint64_t seed, int32_t *Z_samples, int32_t *C_samples,
                               int *n_samples_written, int32_t *Z_last, int32_t *C_last)
{
    if (!h->vc) return fail(h, NBMF_ERR_STATE, "no data");
    if (K <= 0 || iter < 0 || !(alpha > 0) || !(beta > 0) || !(gamma > 0) || !(burnin >= 0) || burnin > 1)
        return fail(h, NBMF_ERR_ARG, "collapsed sampler: bad arguments");
    if (!h->Z) return fail(h, NBMF_ERR_STATE, "the collapsed sweep needs labels: call nbmf_init_from_labels(Z, K)");
    if (h->hook) return fail(h, NBMF_ERR_UNSUPPORTED, "the collapsed wavefront sweep is replicas-only (no column sharding)");
    if (variant == 1 && h->has_na)
        return fail(h, NBMF_ERR_UNSUPPORTED, "Gibbs_DirBer_DP_Rcpp has no NA test (collapsed_gibbs_z.cpp:308-310): V must be fully observed");
    if (variant == 2 && K < 2) return fail(h, NBMF_ERR_ARG, "Gibbs_DirDir_finite_Rcpp needs K >= 2 (a zero entry forbids one component)");
    CK(cudaSetDevice(h->device));
    const int64_t F = h->F, N = h->N;
    const int nburnin = (int)std::floor(iter * burnin);
    const int nsamples = iter - (int)std::floor(iter * burnin);
    const size_t slice = (size_t)F * N * 4;
    const bool two = variant == 2;
    const bool keep = (Z_samples != nullptr || (two && C_samples != nullptr)) && nsamples > 0;
    if (keep && (double)slice * nsamples * (two ? 2 : 1) > 32e9)
        return fail(h, NBMF_ERR_UNSUPPORTED, "the sample cubes do not fit on the device: pass NULL (nbmf_gibbs_dirber returns the posterior means without them)");
    int32_t *cF = nullptr, *cN = nullptr, *cF2 = nullptr, *cN2 = nullptr, *Zs = nullptr, *Cs = nullptr, *Cd = nullptr;
    int *nw = nullptr;
    unsigned long long *cnt = (unsigned long long *)h->scal + SC_TMP;
    const size_t fk = (size_t)F * K, nk = (size_t)2 * N * K;
    cudaError_t e = cudaMalloc(&cF, fk * 4);
    if (e == cudaSuccess) e = cudaMalloc(&cN, nk * 4);
    if (e == cudaSuccess) e = cudaMalloc(&nw, 4);
    if (e == cudaSuccess && keep) e = cudaMalloc(&Zs, slice * nsamples);
    if (e == cudaSuccess && two) e = cudaMalloc(&cF2, fk * 4);
    if (e == cudaSuccess && two) e = cudaMalloc(&cN2, nk * 4);
    if (e == cudaSuccess && two) e = cudaMalloc(&Cd, slice);
    if (e == cudaSuccess && two && keep) e = cudaMalloc(&Cs, slice * nsamples);
    int rc = NBMF_OK;
    if (e == cudaSuccess) {
        cudaMemsetAsync(cF, 0, fk * 4, h->stream); cudaMemsetAsync(cN, 0, nk * 4, h->stream);
        cudaMemsetAsync(nw, 0, 4, h->stream); cudaMemsetAsync(cnt, 0, 16, h->stream);
        if (keep) cudaMemsetAsync(Zs, 0, slice * nsamples, h->stream);   // :380-381
        if (two && keep) cudaMemsetAsync(Cs, 0, slice * nsamples, h->stream);
        counts_from_labels_kernel<<<nblk(F * N, 256), 256, 0, h->stream>>>(h->Z, h->vc, h->has_na ? h->mc : nullptr, F, N, h->Fw, K, cF, cN, cnt + 1);
        e = cudaGetLastError();
        if (e == cudaSuccess && two) {   // C = Z, cnt_c = cnt_z (:465-473)
            e = cudaMemcpyAsync(Cd, h->Z, slice, cudaMemcpyDeviceToDevice, h->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(cF2, cF, fk * 4, cudaMemcpyDeviceToDevice, h->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(cN2, cN, nk * 4, cudaMemcpyDeviceToDevice, h->stream);
        }
        unsigned long long bad = 0;
        if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, cnt + 1, 8, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e == cudaSuccess && bad) rc = fail(h, NBMF_ERR_ARG, "Z contains labels outside 0..K-1");
    }
    cudaEventRecord(h->ev[0], h->stream);
    h->timing = nbmf_timing{};
    if (e == cudaSuccess && rc == NBMF_OK && iter > 1) {
        GibbsWaveArgs a{};
        a.vc = h->vc; a.mc = h->has_na ? h->mc : nullptr; a.F = F; a.N = N; a.Fw = h->Fw;
        a.K = K; a.iter = iter; a.nburnin = nburnin; a.nsamples = nsamples;
        a.alpha = alpha; a.beta = beta; a.gamma = gamma; a.seed = seed;
        a.Z = h->Z; a.cF = cF; a.cN = cN; a.Z_samples = keep ? Zs : nullptr; a.n_written = nw;
        a.C = Cd; a.cF2 = cF2; a.cN2 = cN2; a.C_samples = keep ? Cs : nullptr;
        const int64_t diag = std::min(F, N);
        if (variant == 0) e = launch_wavefront<GibbsWave<0>::Kern>(h, a, K, diag);
        else if (variant == 1) e = launch_wavefront<GibbsWave<1>::Kern>(h, a, K, diag);
        else e = launch_wavefront<GibbsWave<2>::Kern>(h, a, K, diag);
        h->timing.launches += 1;
    }
    cudaEventRecord(h->ev[6], h->stream);
    int written = 0;
    if (e == cudaSuccess && rc == NBMF_OK) {
        e = cudaMemcpyAsync(&written, nw, 4, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess && keep && Z_samples) e = cudaMemcpyAsync(Z_samples, Zs, slice * nsamples, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess && keep && two && C_samples) e = cudaMemcpyAsync(C_samples, Cs, slice * nsamples, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess && Z_last) e = cudaMemcpyAsync(Z_last, h->Z, slice, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess && two && C_last) e = cudaMemcpyAsync(C_last, Cd, slice, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        float ms = 0; cudaEventElapsedTime(&ms, h->ev[0], h->ev[6]);
        h->timing.total_ms = ms; h->timing.iters = iter > 1 ? iter - 1 : 0;
    }
    cudaFree(cF); cudaFree(cN); cudaFree(nw); cudaFree(Zs); cudaFree(cF2); cudaFree(cN2); cudaFree(Cd); cudaFree(Cs);
    if (e != cudaSuccess) { h->err = std::string("collapsed Gibbs: ") + cudaGetErrorString(e); cudaGetLastError(); return NBMF_ERR_CUDA; }
    if (rc) return rc;
    if (n_samples_written) *n_samples_written = written;
    return NBMF_OK;
}

extern "C" int nbmf_gibbs_dirber_collapsed(nbmf_handle *h, int K, double alpha, double beta, double gamma,
                                           int iter, double burnin, uint64_t seed, int32_t *Z_samples,
                                           int *n_samples_written, int32_t *Z_last)
{
    if (!h) return NBMF_ERR_ARG;
    return collapsed_wavefront(h, 0, K, alpha, beta, gamma, iter, burnin, seed, Z_samples, nullptr, n_samples_written, Z_last, nullptr);
}

// Kmax = 200 components (collapsed_gibbs_z.cpp:282); alpha and beta are accepted and ignored, as in the
// reference (its Beta prior is the constant epsilon = 1e-8, :281,317-318).  The labels must have been
// installed with nbmf_init_from_labels(Z, 200).
extern "C" int nbmf_gibbs_dirber_dp_collapsed(nbmf_handle *h, double alpha, double beta, double gamma, int iter,
                                              double burnin, uint64_t seed, int32_t *Z_samples,
                                              int *n_samples_written, int32_t *Z_last)
{
    if (!h) return NBMF_ERR_ARG;
    (void)alpha; (void)beta;
    if (h->Kc != 200) return fail(h, NBMF_ERR_STATE, "Gibbs_DirBer_DP_Rcpp works on Kmax = 200 components: call nbmf_init_from_labels(Z, 200)");
    return collapsed_wavefront(h, 1, 200, 1.0, 1.0, gamma, iter, burnin, seed, Z_samples, nullptr, n_samples_written, Z_last, nullptr);
}

extern "C" int nbmf_gibbs_dirdir_collapsed(nbmf_handle *h, int K, double alpha, double gamma, int iter, double burnin,
                                           uint64_t seed, int32_t *Z_samples, int32_t *C_samples,
                                           int *n_samples_written, int32_t *Z_last, int32_t *C_last)
{
    if (!h) return NBMF_ERR_ARG;
    return collapsed_wavefront(h, 2, K, alpha, 1.0, gamma, iter, burnin, seed, Z_samples, C_samples, n_samples_written, Z_last, C_last);
}

extern "C" int nbmf_gibbs_sweep_given_W(nbmf_handle *h, const int32_t *v_n, const double *W, int64_t F,
                                        int K, double alpha, int iter, double burnin, uint64_t seed,
                                        double *C_mean)
{
    if (!h) return NBMF_ERR_ARG;
    if (!v_n || !W || !C_mean || F <= 0 || K <= 0 || K > 1024) return fail(h, NBMF_ERR_ARG, "bad arguments");
    CK(cudaSetDevice(h->device));
    DevBuf<int32_t> dv, lab;
    DevBuf<double> dW, dC;
    CK(dv.alloc((size_t)F)); CK(lab.alloc((size_t)F));
    CK(dW.alloc((size_t)F * K)); CK(dC.alloc((size_t)F * K));
    CK(cudaMemcpyAsync(dv, v_n, (size_t)F * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dW, W, (size_t)F * K * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(dC, 0, (size_t)F * K * 8, h->stream));
    gibbs_given_W_kernel<<<1, 1024, (size_t)K * 12 + 16, h->stream>>>(dv, dW, F, K, alpha, iter, (int)std::floor(iter * burnin), seed, lab, dC);
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(C_mean, dC, (size_t)F * K * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return NBMF_OK;
}

// ---------------------------------------------------------------------------
// predictive log-likelihood
// ---------------------------------------------------------------------------
extern "C" int nbmf_predictive_ll(nbmf_handle *h, const double *E_W, const double *E_H, int K,
                                  const int32_t *Vt, int64_t F, int64_t N, double *ll, int64_t *n_test)
{
    if (!h) return NBMF_ERR_ARG;
    if (!E_W || !E_H || !Vt || F <= 0 || N <= 0 || K <= 0) return fail(h, NBMF_ERR_ARG, "bad arguments");
    CK(cudaSetDevice(h->device));
    DevBuf<double> dW, dH, stg, dll;
    DevBuf<int32_t> dV;
    int64_t cols_per = std::min<int64_t>(N, std::max<int64_t>(1, (int64_t)(256ll << 20) / (F * 4)));
    CK(dW.alloc((size_t)F * K)); CK(dH.alloc((size_t)N * K));
    CK(stg.alloc((size_t)std::max(F, N) * K)); CK(dll.alloc(2));
    CK(dV.alloc((size_t)cols_per * F));
    CK(cudaMemsetAsync(dll, 0, 16, h->stream));
    CK(cudaMemcpyAsync(stg, E_W, (size_t)F * K * 8, cudaMemcpyHostToDevice, h->stream));
    colmajor_to_rows_kernel<<<nblk(F * K, 256), 256, 0, h->stream>>>(stg, F, K, dW, K, 1, 0);
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpyAsync(stg, E_H, (size_t)N * K * 8, cudaMemcpyHostToDevice, h->stream));
    colmajor_to_rows_kernel<<<nblk(N * K, 256), 256, 0, h->stream>>>(stg, N, K, dH, K, 1, 0);
    CK(cudaStreamSynchronize(h->stream));
    for (int64_t n0 = 0; n0 < N; n0 += cols_per) {
        int64_t nc = std::min(cols_per, N - n0);
        CK(cudaMemcpyAsync(dV, Vt + n0 * F, (size_t)nc * F * 4, cudaMemcpyHostToDevice, h->stream));
        predictive_ll_kernel<<<nblk(nc * F, 256), 256, 0, h->stream>>>(dV, F, nc, n0, K, K, dW, dH, dll, (unsigned long long *)(dll.p + 1));
        LAUNCH_CHECK();
        CK(cudaStreamSynchronize(h->stream));
    }
    double out[2];
    CK(cudaMemcpy(out, dll, 16, cudaMemcpyDeviceToHost));
    if (ll) *ll = out[0];
    if (n_test) { unsigned long long c; memcpy(&c, &out[1], 8); *n_test = (int64_t)c; }
    return NBMF_OK;
}
