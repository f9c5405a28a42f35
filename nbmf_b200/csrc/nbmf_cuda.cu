// nbmf_cuda.cu -- C ABI of libnbmf_cuda (see include/nbmf_cuda.h) and the host
// orchestration of the VB / CVB0 / Gibbs hot path.  No CPU fallback: every
// compute entry point needs a CUDA device.
#include "kernels_basic.cuh"
#include "pass_mma64.cuh"
#include "gibbs_fast.cuh"
#include "tensor_pass.h"
#include "host_common.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <type_traits>
#include <vector>


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
    cudaFree(h->scal); cudaFree(h->lbs); cudaFree(h->stage); cudaFree(h->io);
    if (h->out_stream) { cudaStreamDestroy(h->out_stream); cudaEventDestroy(h->out_ev); }
    for (int i = 0; i < 8; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    for (int i = 0; i < 33; i++) if (h->pipe_ev[i]) cudaEventDestroy(h->pipe_ev[i]);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->xchg_stream) {
        cudaStreamDestroy(h->xchg_stream); cudaEventDestroy(h->xchg_ev[0]); cudaEventDestroy(h->xchg_ev[1]);
        cudaEventDestroy(h->xchg_t[0]); cudaEventDestroy(h->xchg_t[1]);
    }
    nbmf_comm_release(h);
    cudaFree(h->hook_scratch);
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
    tp_invalidate_data(h);   // the exact row counts of the tensor path were summed with the previous set of ranks
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
    if (!nbmf_sharded(h)) return NBMF_OK;
    return nbmf_sum_ranks(h, buf, count, 0, h->stream);
}

// The same exchange started on a side stream behind the current point of the main stream, so
// that whatever the main stream launches next overlaps it; exchange_end makes the main stream
// wait for the result.  The side stream has the highest priority: the collective's few CTAs are
// placed before the persistent contraction kernel that follows, whose work queue absorbs the
// SMs that come late.  The collective itself is bracketed by events on the side stream
// (timing.exchange_ms is its device time, not the time the main stream spent waiting).
static int exchange_begin(nbmf_handle *h, double *buf, size_t count)
{
    if (!nbmf_sharded(h)) return NBMF_OK;
    if (!h->xchg_stream) {
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&h->xchg_stream, cudaStreamNonBlocking, hi));
        CK(cudaEventCreateWithFlags(&h->xchg_ev[0], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->xchg_ev[1], cudaEventDisableTiming));
        CK(cudaEventCreate(&h->xchg_t[0]));
        CK(cudaEventCreate(&h->xchg_t[1]));
    }
    CK(cudaEventRecord(h->xchg_ev[0], h->stream));
    CK(cudaStreamWaitEvent(h->xchg_stream, h->xchg_ev[0], 0));
    CK(cudaEventRecord(h->xchg_t[0], h->xchg_stream));
    int rc = nbmf_sum_ranks(h, buf, count, 0, h->xchg_stream);
    if (rc) return rc;
    CK(cudaEventRecord(h->xchg_t[1], h->xchg_stream));
    CK(cudaEventRecord(h->xchg_ev[1], h->xchg_stream));
    h->xchg_pending = true;
    h->xchg_timed = true;
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
    if (res[1]) {   // part of the planes has been overwritten: leave the handle without data rather than with a mixture
        free_data(h);
        return fail(h, NBMF_ERR_ARG, "V contains values other than 0, 1 and NA");
    }
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
    if (K <= 8) { *KC = 8; *NKC = 1; }   // Kp >= 8: one k tile of the DMMA kernel
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

static int ensure_io(nbmf_handle *h, size_t bytes)
{
    if (h->io_bytes >= bytes) return NBMF_OK;
    CK(cudaStreamSynchronize(h->stream));
    if (h->out_stream) CK(cudaStreamSynchronize(h->out_stream));
    cudaFree(h->io); h->io = nullptr; h->io_bytes = 0;
    if (cudaMalloc(&h->io, bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(h, NBMF_ERR_OOM, "device staging buffer allocation failed");
    }
    h->io_bytes = bytes;
    return NBMF_OK;
}

// host column-major [ld x K] (rows `rows` of it) -> device row-major [rows][Kp], asynchronously on `st`,
// through the staging area at byte offset `off` of h->io (the caller sized it with ensure_io)
static int upload_rows_async(nbmf_handle *h, const double *src, int64_t ld, int64_t rows, int K, double *dst,
                             int mul, int add, size_t off, cudaStream_t st)
{
    double *stg = (double *)((char *)h->io + off);
    if (ld == rows) CK(cudaMemcpyAsync(stg, src, (size_t)rows * K * 8, cudaMemcpyHostToDevice, st));
    else CK(cudaMemcpy2DAsync(stg, (size_t)rows * 8, src, (size_t)ld * 8, (size_t)rows * 8, (size_t)K, cudaMemcpyHostToDevice, st));
    colmajor_to_rows_kernel<<<nblk(rows * K, 256), 256, 0, st>>>(stg, rows, K, dst, h->Kp, mul, add);
    LAUNCH_CHECK();
    return NBMF_OK;
}
static int download_rows_async(nbmf_handle *h, const double *src, int64_t rows, int K, double *dst, int64_t ld,
                               int mul, int add, size_t off, cudaStream_t st)
{
    double *stg = (double *)((char *)h->io + off);
    rows_to_colmajor_kernel<<<nblk(rows * K, 256), 256, 0, st>>>(src, rows, K, stg, h->Kp, mul, add);
    LAUNCH_CHECK();
    if (ld == rows) CK(cudaMemcpyAsync(dst, stg, (size_t)rows * K * 8, cudaMemcpyDeviceToHost, st));
    else CK(cudaMemcpy2DAsync(dst, (size_t)ld * 8, stg, (size_t)rows * 8, (size_t)rows * 8, (size_t)K, cudaMemcpyDeviceToHost, st));
    return NBMF_OK;
}
static int upload_rows(nbmf_handle *h, const double *src, int64_t rows, int K, double *dst, int mul, int add)
{
    int rc = ensure_io(h, (size_t)rows * K * 8);
    if (rc) return rc;
    if ((rc = upload_rows_async(h, src, rows, rows, K, dst, mul, add, 0, h->stream))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return NBMF_OK;
}
static int download_rows(nbmf_handle *h, const double *src, int64_t rows, int K, double *dst, int mul, int add)
{
    int rc = ensure_io(h, (size_t)rows * K * 8);
    if (rc) return rc;
    if ((rc = download_rows_async(h, src, rows, K, dst, rows, mul, add, 0, h->stream))) return rc;
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
    CK(cudaStreamSynchronize(h->stream));
    if ((rc = tp_check_error(h))) return rc;   // an aborted tensor pass leaves garbage behind
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
    return h->path_mode == NBMF_PATH_TENSOR || h->path_mode == NBMF_PATH_TENSOR_FAST || h->path_mode == NBMF_PATH_TENSOR_LO8;
}

// Fixes the contraction variant of a session from the precision option, K, the sizes and -- under
// AUTO -- the number of updates the caller plans (the f16 variants' error grows with the iteration
// count): must run before anything asks use_tensor_path().
static void resolve_path(nbmf_handle *h, int K, int horizon)
{
    const int p = h->opts.precision;
    if (horizon <= 0) horizon = h->opts.planned_iters > 0 ? h->opts.planned_iters : 100;   // VB_Aspect_Rcpp's default iter
    h->vb_horizon = horizon;
    if (p == NBMF_PREC_FP64) h->path_mode = NBMF_PATH_FP64;
    else if (p == NBMF_PREC_FP32) h->path_mode = NBMF_PATH_FP32;
    else if (p == NBMF_PREC_TENSOR) h->path_mode = NBMF_PATH_TENSOR;
    else if (p == NBMF_PREC_TENSOR_FAST) h->path_mode = NBMF_PATH_TENSOR_FAST;
    else if (p == NBMF_PREC_TENSOR_LO8) h->path_mode = NBMF_PATH_TENSOR_LO8;
    else h->path_mode = tp_auto_mode(h, K, horizon);
}

template <typename T, int KC>
static cudaError_t launch_simt(const PassArgs &a, dim3 grid, size_t smem, cudaStream_t s)
{
    cudaError_t e = cudaFuncSetAttribute(pass_simt_kernel<T, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    pass_simt_kernel<T, KC><<<grid, 256, smem, s>>>(a);
    return cudaGetLastError();
}

template <int KP>
static cudaError_t launch_mma64(const Mma64Args &a, dim3 grid, bool rows, bool want_log, cudaStream_t s)
{
    const size_t smem = Mma64Cfg<KP>::SMEM;
    cudaError_t e;
    if (rows && want_log) {
        e = cudaFuncSetAttribute(pass_mma64_kernel<KP, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) pass_mma64_kernel<KP, true, true><<<grid, 256, smem, s>>>(a);
    } else if (rows) {
        e = cudaFuncSetAttribute(pass_mma64_kernel<KP, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) pass_mma64_kernel<KP, true, false><<<grid, 256, smem, s>>>(a);
    } else {
        e = cudaFuncSetAttribute(pass_mma64_kernel<KP, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) pass_mma64_kernel<KP, false, false><<<grid, 256, smem, s>>>(a);
    }
    return e != cudaSuccess ? e : cudaGetLastError();
}

// FP64 pass on the DMMA pipe (pass_mma64.cuh), Kp in {8, 16, 32, 64, 128}
static int run_pass_mma64(nbmf_handle *h, int owner_rows, double *sumlogD)
{
    Mma64Args a{};
    a.sumlogD = sumlogD;
    if (owner_rows) {
        a.U = h->A; a.W = h->BB; a.G = h->GF;
        a.vbits = h->vr; a.mbits = h->has_na ? h->mr : nullptr;
        a.words = h->Nw; a.L = h->F; a.S_rows = 2 * h->N; a.ob = 1; a.sb = 2;
    } else {
        a.U = h->BB; a.W = h->A; a.G = h->GN;
        a.vbits = h->vc; a.mbits = h->has_na ? h->mc : nullptr;
        a.words = h->Fw; a.L = 2 * h->N; a.S_rows = h->F; a.ob = 2; a.sb = 1;
    }
    const int64_t bx = (a.L + 63) / 64;
    const int64_t steps = (a.S_rows + 63) / 64;
    // enough blocks for a few waves; a split streams at least 8 steps so that the owner tile load amortises
    const int64_t want = 4ll * h->sm_count;
    int64_t splits = std::max<int64_t>(1, std::min<int64_t>((want + bx - 1) / bx, (steps + 7) / 8));
    a.rows_per_split = ((steps + splits - 1) / splits) * 64;
    splits = (a.S_rows + a.rows_per_split - 1) / a.rows_per_split;
    if (splits > 1) CK(cudaMemsetAsync(a.G, 0, (size_t)a.L * h->Kp * 8, h->stream));
    dim3 grid((unsigned)bx, (unsigned)splits);
    cudaError_t e;
    const bool wl = sumlogD != nullptr;
    switch (h->Kp) {
    case 8: e = launch_mma64<8>(a, grid, owner_rows, wl, h->stream); break;
    case 16: e = launch_mma64<16>(a, grid, owner_rows, wl, h->stream); break;
    case 32: e = launch_mma64<32>(a, grid, owner_rows, wl, h->stream); break;
    case 64: e = launch_mma64<64>(a, grid, owner_rows, wl, h->stream); break;
    case 128: e = launch_mma64<128>(a, grid, owner_rows, wl, h->stream); break;
    default: return fail(h, NBMF_ERR_STATE, "internal: DMMA pass with Kp > 128");
    }
    CK(e);
    h->timing.launches += 1 + (splits > 1);
    return NBMF_OK;
}

// owner_rows != 0: owner = rows f, output GF; else owner = (n,b) lanes, output GN.
// fp32 = the FFMA variant (NBMF_PREC_FP32); fp64 SIMT serves Kp > 128 (and NBMF_FP64_KERNEL=simt).
static int run_pass_simt(nbmf_handle *h, int owner_rows, double *sumlogD, bool fp32)
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
    const size_t esz = fp32 ? 4 : 8;
    int TR = 32;
    while ((size_t)TR * a.sb * a.Kp * esz > 65536 && TR > 1) TR >>= 1;
    a.TR = TR;
    size_t smem = (size_t)TR * a.sb * a.Kp * esz;
    const int OL = 256 / h->NKC;
    int64_t bx = (a.L + OL - 1) / OL;
    int64_t want = 4ll * h->sm_count;
    int64_t splits = std::max<int64_t>(1, std::min<int64_t>((want + bx - 1) / bx, (a.S + 255) / 256));
    a.s_per_split = round_up64((a.S + splits - 1) / splits, 32);
    splits = (a.S + a.s_per_split - 1) / a.s_per_split;
    if (splits > 1) CK(cudaMemsetAsync(a.G, 0, (size_t)a.L * a.Kp * 8, h->stream));
    dim3 grid((unsigned)bx, (unsigned)splits);
    cudaError_t e;
    if (fp32) {
        if (h->KC == 8) e = launch_simt<float, 8>(a, grid, smem, h->stream);
        else e = launch_simt<float, 16>(a, grid, smem, h->stream);
    } else {
        if (h->KC == 8) e = launch_simt<double, 8>(a, grid, smem, h->stream);
        else e = launch_simt<double, 16>(a, grid, smem, h->stream);
    }
    CK(e);
    h->timing.launches += 1 + (splits > 1);
    return NBMF_OK;
}

static int run_pass(nbmf_handle *h, int owner_rows, double *sumlogD)
{
    if (use_tensor_path(h)) return tp_pass(h, owner_rows, sumlogD);
    if (h->path_mode == NBMF_PATH_FP32) return run_pass_simt(h, owner_rows, sumlogD, true);
    static const bool force_simt = [] { const char *v = getenv("NBMF_FP64_KERNEL"); return v && strcmp(v, "simt") == 0; }();
    if (h->Kp <= 128 && !force_simt) return run_pass_mma64(h, owner_rows, sumlogD);
    return run_pass_simt(h, owner_rows, sumlogD, false);
}

// slice: 0 = every row; 1 = every row, but only the rank's own rows enter the ELBO's prior term (first iteration of a
// row-sliced session: the statistics are still whole everywhere); 2 = the rank's own rows only
static int run_params(nbmf_handle *h, int K, double alpha, double beta, double gamma, int mean_mode,
                      int want_lb, int slice = 0)
{
    const int64_t f0 = slice == 2 ? h->row_lo : 0, f1 = slice == 2 ? h->row_hi : h->F;
    const int64_t lb0 = slice ? h->row_lo : 0, lb1 = slice ? h->row_hi : h->F;
    if (f1 > f0) {
        params_rows_kernel<<<nblk((f1 - f0) * 32, 256), 256, 0, h->stream>>>(h->statF, f0, f1, K, h->Kp, gamma,
                                                                             mean_mode, want_lb, lb0, lb1, h->A, h->EW, h->scal);
        LAUNCH_CHECK();
    }
    params_cols_kernel<<<nblk(h->N * 32, 256), 256, 0, h->stream>>>(h->statN, h->N, K, h->Kp, alpha, beta,
                                                                    mean_mode, want_lb, h->BB, h->EH, h->scal);
    LAUNCH_CHECK();
    h->timing.launches += 2;
    if (use_tensor_path(h)) return tp_convert_operands(h, slice == 2);
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
    resolve_path(h, K, h->horizon_next);
    h->horizon_next = 0;
    if (use_tensor_path(h)) {
        int rc = tp_prepare(h);
        if (rc) return rc;
    }
    h->xchg_timed = false;
    // row slicing: library communicator + tensor path (NBMF_ROW_SLICING=0 keeps the row side replicated)
    static const bool slicing_on = [] { const char *v = getenv("NBMF_ROW_SLICING"); return !(v && atoi(v) == 0); }();
    h->vb_sliced = slicing_on && h->comm != nullptr && h->comm_world > 1 && use_tensor_path(h);
    h->row_lo = 0; h->row_hi = h->F;
    if (h->vb_sliced) nbmf_row_block(h->F, h->comm_world, h->comm_rank, &h->row_lo, &h->row_hi);
    h->vb_active = true; h->vb_K = K; h->vb_iter = 0; h->vb_mean_params = mean_params;
    h->vb_alpha = alpha; h->vb_beta = beta; h->vb_gamma = gamma;
    h->timing = nbmf_timing{};
    h->timing.path = h->path_mode;
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
    if ((rc = run_params(h, h->vb_K, h->vb_alpha, h->vb_beta, h->vb_gamma, h->vb_mean_params, checklb,
                         h->vb_sliced ? (i == 0 ? 1 : 2) : 0))) return rc;
    CK(cudaEventRecord(h->ev[2], h->stream));
    // row-sliced session: E_W exists in row blocks spread over the ranks from the second iteration on (a collective:
    // the condition must not depend on which buffers this rank was given)
    if (h->early_iter == i && h->vb_sliced && i > 0 && (rc = nbmf_allgather_rows(h, h->EW, (size_t)h->Kp * 8, h->F, h->stream))) return rc;
    if (h->early_iter == i && (h->early_EW || h->early_EH)) {
        // E_W / E_H of the call are the parameters at the top of its LAST iteration
        // (collapsed_gibbs_z_VB.cpp:507-514): final as soon as that iteration's parameter kernels
        // have run, so their download overlaps the contraction (and, full duplex, the upload of V)
        const size_t off = (size_t)(h->F + 2 * h->N) * h->vb_K * 8;   // behind the statistics' staging area
        CK(cudaEventRecord(h->out_ev, h->stream));
        CK(cudaStreamWaitEvent(h->out_stream, h->out_ev, 0));
        if (h->early_EW && (rc = download_rows_async(h, h->EW, h->F, h->vb_K, h->early_EW, h->F, 1, 0, off, h->out_stream))) return rc;
        if (h->early_EH && (rc = download_rows_async(h, h->EH, h->N, h->vb_K, h->early_EH, h->N, 1, 0,
                                                     off + (size_t)h->F * h->vb_K * 8, h->out_stream))) return rc;
        h->early_done = true;
    }
    if (h->pipe_blocks > 0 && !use_tensor_path(h)) { h->pipe_blocks = 0; return fail(h, NBMF_ERR_STATE, "internal: pipelined upload without the tensor path"); }
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
        if (checklb && (rc = tp_log_correction_rows_slice(h, h->scal + SC_SUMLOGD))) return rc;
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
    if (checklb && use_tensor_path(h) && (rc = tp_log_correction_rows_slice(h, h->scal + SC_SUMLOGD))) return rc;
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
    if (h->vb_sliced && h->vb_iter > 0) {
        // row-sliced session: make the replicated fp64 arrays whole again (every rank holds its own rows of the
        // statistics, and -- from the second iteration on -- of the parameters), and sum the ELBO's row term,
        // which every rank accumulated over its own rows only, so that it is the replicated term callers expect
        int rc0;
        if ((rc0 = nbmf_allgather_rows(h, h->statF, (size_t)h->Kp * 8, h->F, h->stream))) return rc0;
        if (h->vb_iter > 1) {
            if ((rc0 = nbmf_allgather_rows(h, h->EW, (size_t)h->Kp * 8, h->F, h->stream))) return rc0;
            if ((rc0 = nbmf_allgather_rows(h, h->A, (size_t)h->Kp * 8, h->F, h->stream))) return rc0;
        }
        const int n = std::min(h->vb_iter, h->lbs_cap);
        if ((rc0 = ensure_stage(h, (size_t)n * 8))) return rc0;
        strided_copy_kernel<<<nblk(n, 256), 256, 0, h->stream>>>(h->lbs + 1, 3, (double *)h->stage, 1, n);
        LAUNCH_CHECK();
        if ((rc0 = nbmf_sum_ranks(h, h->stage, (size_t)n, 0, h->stream))) return rc0;
        strided_copy_kernel<<<nblk(n, 256), 256, 0, h->stream>>>((const double *)h->stage, 1, h->lbs + 1, 3, n);
        LAUNCH_CHECK();
    }
    h->vb_sliced = false;
    CK(cudaEventRecord(h->ev[6], h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->vb_active = false;
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[6]); h->timing.total_ms = ms;
    if (h->vb_iter > 0) {
        cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]); h->timing.params_ms = ms;
        cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]); h->timing.pass_rows_ms = ms;
        cudaEventElapsedTime(&ms, h->ev[3], h->ev[4]); h->timing.exchange_ms = ms;
        if (h->xchg_timed && cudaEventElapsedTime(&ms, h->xchg_t[0], h->xchg_t[1]) == cudaSuccess)
            h->timing.exchange_ms = ms;   // overlapped exchange: the collective's own device time on the side stream
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
    if (h->early_done && h->early_EW == E_W && h->early_EH == E_H) {
        CK(cudaStreamSynchronize(h->out_stream));   // already on their way since the last iteration's parameter kernels
    } else {
        if (h->early_done) CK(cudaStreamSynchronize(h->out_stream));
        if (E_W && (rc = download_rows(h, h->EW, h->F, K, E_W, 1, 0))) return rc;
        if (E_H && (rc = download_rows(h, h->EH, h->N, K, E_H, 1, 0))) return rc;
    }
    h->early_done = false; h->early_iter = -1; h->early_EW = h->early_EH = nullptr;
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
    { int rc0 = tp_check_error(h); if (rc0) return rc0; }
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
    h->horizon_next = std::max(1, iter);   // AUTO precision: this call's own iteration count
    int rc = nbmf_vb_begin(h, K, alpha, beta, gamma, 0);
    if (rc) return rc;
    for (int i = 0; i < iter; i++)
        if ((rc = nbmf_vb_step(h, checklb))) { h->vb_active = false; return rc; }
    return nbmf_vb_end(h, E_W, E_H, lowerbounds, iter, E_Z);
}

// VB_Aspect_Rcpp with V given as host bit planes and the initial statistics as host arrays: one
// call = upload + iter iterations + results.  The upload is cut into column blocks on a copy
// stream and the first iteration consumes the blocks as they land (both passes of a block need
// nothing but that block's columns), so the PCIe transfer overlaps the contraction; the blocks
// shrink towards the end so that little compute is left when the last one has landed.  E_W / E_H
// leave beside the last iteration (see nbmf_vb_step).  With a communicator the row-side arrays are
// the root's: n_total is read on rank 0 and broadcast over NVLink, E_W may be NULL elsewhere.
extern "C" int nbmf_vb_aspect_bits(nbmf_handle *h, const uint32_t *vbits, int64_t F, int64_t N, int K,
                                   double alpha, double beta, double gamma, int iter, int checklb,
                                   const double *n_total, const double *f_ones, const double *f_zeros,
                                   double *E_W, double *E_H, double *lowerbounds)
{
    if (!h) return NBMF_ERR_ARG;
    if (!vbits || F <= 0 || N <= 0 || iter < 0 || K <= 0) return fail(h, NBMF_ERR_ARG, "nbmf_vb_aspect_bits: bad arguments");
    const bool root = h->comm == nullptr || h->comm_rank == 0;
    if (!f_ones || !f_zeros || (root && !n_total)) return fail(h, NBMF_ERR_ARG, "nbmf_vb_aspect_bits: NULL statistics");
    CK(cudaSetDevice(h->device));
    int rc = alloc_planes(h, F, N);
    if (rc) return rc;
    h->has_na = false; h->nobs = F * N;
    // the contraction variant follows from THIS call's K and iteration count (not from whatever
    // model the handle held before)
    resolve_path(h, K, std::max(1, iter));
    // column blocks: multiples of 64 columns (one owner tile of the cols pass = two steps of the rows
    // pass); seven of about N/8, then a tail of N/16, N/32, N/32
    int64_t unit = h->opts.pipe_block_cols > 0 ? (int64_t)h->opts.pipe_block_cols
                                               : std::max<int64_t>(2048, ((N + 7) / 8 + 2047) / 2048 * 2048);
    unit = std::max<int64_t>(64, (unit + 63) / 64 * 64);
    while ((N + unit - 1) / unit > 24) unit *= 2;
    int nb = 0;
    h->pipe_n0[0] = 0;
    for (int64_t c = 0; c < N && nb < 32;) {
        int64_t len = unit;
        const int64_t left = N - c;
        if (h->opts.pipe_block_cols <= 0 && left <= unit && unit >= 4 * 2048) {   // the tail: halves, quarters
            len = left > unit / 2 ? unit / 2 : (left > unit / 4 ? unit / 4 : left);
            len = std::max<int64_t>(64, (len + 63) / 64 * 64);
        }
        c = std::min(N, c + len);
        h->pipe_n0[++nb] = c;
    }
    if (h->pipe_n0[nb] < N) h->pipe_n0[nb] = N;
    const bool pipelined = iter >= 1 && nb >= 2 && use_tensor_path(h);
    if (!pipelined) {
        if ((rc = nbmf_set_data_bits(h, vbits, nullptr, F, N))) return rc;
        if (h->comm && h->comm_world > 1) {
            // install the local statistics, then the root's row-side statistic over NVLink
            std::vector<double> zero;
            const double *nt = n_total;
            if (!nt) { zero.assign((size_t)F * K, 0.0); nt = zero.data(); }
            if ((rc = nbmf_set_stats(h, K, nt, f_ones, f_zeros))) return rc;
            if ((rc = nbmf_bcast_root(h, h->statF, (size_t)F * h->Kp * 8, h->stream))) return rc;
        } else if ((rc = nbmf_set_stats(h, K, n_total, f_ones, f_zeros))) return rc;
        return nbmf_vb_aspect(h, K, alpha, beta, gamma, iter, checklb, E_W, E_H, lowerbounds, nullptr);
    }
    tp_invalidate_data(h);
    if (!h->copy_stream) CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    if (!h->out_stream) {
        CK(cudaStreamCreateWithFlags(&h->out_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&h->out_ev, cudaEventDisableTiming));
    }
    // the statistics first (the parameters of the first iteration need all of them), asynchronously,
    // each array through its own part of the staging area; then V, block by block, on the copy stream
    h->Kc = 0;
    if ((rc = ensure_model(h, K))) return rc;
    const size_t fk = (size_t)F * K * 8, nk = (size_t)N * K * 8;
    if ((rc = ensure_io(h, 2 * (fk + 2 * nk)))) return rc;   // uploads in [0, fk + 2 nk), early outputs behind them
    if (root && (rc = upload_rows_async(h, n_total, F, F, K, h->statF, 1, 0, 0, h->stream))) return rc;
    if ((rc = upload_rows_async(h, f_ones, N, N, K, h->statN, 2, 1, fk, h->stream))) return rc;
    if ((rc = upload_rows_async(h, f_zeros, N, N, K, h->statN, 2, 0, fk + nk, h->stream))) return rc;
    if (h->comm && h->comm_world > 1 && (rc = nbmf_bcast_root(h, h->statF, (size_t)F * h->Kp * 8, h->stream))) return rc;
    // the first block's copy is queued behind the statistics (one PCIe direction serves both)
    CK(cudaEventRecord(h->out_ev, h->stream));
    CK(cudaStreamWaitEvent(h->copy_stream, h->out_ev, 0));
    const int64_t fw_in = (F + 31) / 32;
    for (int b = 0; b < nb; b++) {
        const int64_t c0 = h->pipe_n0[b], c1 = h->pipe_n0[b + 1];
        if (!h->pipe_ev[b]) CK(cudaEventCreateWithFlags(&h->pipe_ev[b], cudaEventDisableTiming));
        CK(cudaMemcpy2DAsync(h->vc + c0 * h->Fw, (size_t)h->Fw * 4, vbits + c0 * fw_in, (size_t)fw_in * 4,
                             (size_t)fw_in * 4, (size_t)(c1 - c0), cudaMemcpyHostToDevice, h->copy_stream));
        CK(cudaEventRecord(h->pipe_ev[b], h->copy_stream));
    }
    // from here on the host buffer is in use by the copy stream: every exit synchronises it, and a
    // failure leaves the handle without data rather than with half-built planes
    auto bail = [&](int code) {
        cudaStreamSynchronize(h->copy_stream);
        cudaStreamSynchronize(h->stream);
        if (h->out_stream) cudaStreamSynchronize(h->out_stream);
        h->vb_active = false; h->pipe_blocks = 0;
        h->early_done = false; h->early_iter = -1; h->early_EW = h->early_EH = nullptr;
        free_data(h);
        return code;
    };
    fill_mask_kernel<<<nblk(N * h->Fw, 256), 256, 0, h->stream>>>(h->mc, N, h->Fw, F);
    if (cudaGetLastError() != cudaSuccess) return bail(fail(h, NBMF_ERR_CUDA, "fill_mask launch failed"));
    fill_mask_kernel<<<nblk(F * h->Nw, 256), 256, 0, h->stream>>>(h->mr, F, h->Nw, N);
    if (cudaGetLastError() != cudaSuccess) return bail(fail(h, NBMF_ERR_CUDA, "fill_mask launch failed"));
    h->horizon_next = std::max(1, iter);
    if ((rc = nbmf_vb_begin(h, K, alpha, beta, gamma, 0))) return bail(rc);
    if (!use_tensor_path(h)) return bail(fail(h, NBMF_ERR_STATE, "internal: pipelined upload without the tensor path"));
    h->pipe_blocks = nb;
    h->early_EW = E_W; h->early_EH = E_H; h->early_iter = iter - 1; h->early_done = false;
    for (int i = 0; i < iter; i++)
        if ((rc = nbmf_vb_step(h, checklb))) return bail(rc);
    rc = nbmf_vb_end(h, E_W, E_H, lowerbounds, iter, nullptr);
    if (rc) return bail(rc);
    cudaStreamSynchronize(h->copy_stream);
    return rc;
}

extern "C" int nbmf_vb_argmax_labels(nbmf_handle *h, int32_t *Z_out)
{
    if (!h) return NBMF_ERR_ARG;
    if (!h->A || !h->BB || h->Kc <= 0) return fail(h, NBMF_ERR_STATE, "no VB operands: run nbmf_vb_aspect / nbmf_vb_step first");
    CK(cudaSetDevice(h->device));   // under column sharding the labels are those of the handle's own block
    CK(cudaStreamSynchronize(h->stream));
    { int rc0 = tp_check_error(h); if (rc0) return rc0; }
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
    resolve_path(h, K, 1);
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
    if (nbmf_sharded(h) && (h->hook || h->comm_world > 1)) return fail(h, NBMF_ERR_UNSUPPORTED, "exact wavefront mode is replicas-only (no column sharding)");
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
    h->horizon_next = std::max(1, iter - 1);
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
template <int KT>
static cudaError_t launch_gibbs_sweep(const GibbsSweepArgs &a, dim3 grid, int warps, cudaStream_t st)
{
    cudaError_t e;
    if (a.mc) {
        e = cudaFuncSetAttribute(gibbs_sweep_kt_kernel<KT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GibbsSmem<KT>::BYTES);
        if (e != cudaSuccess) return e;
        gibbs_sweep_kt_kernel<KT, true><<<grid, 32 * warps, GibbsSmem<KT>::bytes(warps), st>>>(a);
    } else {
        e = cudaFuncSetAttribute(gibbs_sweep_kt_kernel<KT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GibbsSmem<KT>::BYTES);
        if (e != cudaSuccess) return e;
        gibbs_sweep_kt_kernel<KT, false><<<grid, 32 * warps, GibbsSmem<KT>::bytes(warps), st>>>(a);
    }
    return cudaGetLastError();
}

template <int KT>
static int gibbs_sweep_occupancy(int warps)   // resident blocks per SM
{
    int occ = 0;
    // the whole shared-memory carve-out: the resident block count is set by shared memory and registers, not by L1
    cudaFuncSetAttribute(gibbs_sweep_kt_kernel<KT, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(gibbs_sweep_kt_kernel<KT, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(gibbs_sweep_kt_kernel<KT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GibbsSmem<KT>::BYTES);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gibbs_sweep_kt_kernel<KT, false>, 32 * warps, GibbsSmem<KT>::bytes(warps));
    cudaGetLastError();
    return occ;
}

// Uncollapsed blocked chain, K <= 32: two launches per sweep (gibbs_fast.cuh).  Sweep i draws W, H from
// the counts of sweep i-1, then samples every entry; a kept sweep (i >= floor(iter * burnin), the
// reference's bookkeeping, collapsed_gibbs_z.cpp:423-426) adds the Rao-Blackwell means of ITS counts,
// which the next draw launch (or a final accumulation-only one) does.
static int gibbs_dirber_fast(nbmf_handle *h, int K, double alpha, double beta, double gamma, int iter,
                             double burnin, uint64_t seed, double *E_W, double *E_H,
                             const int64_t *test_idx, int64_t n_test, double *E_V_test, int32_t *Z_last)
{
    const int64_t F = h->F, N = h->N;
    const int KT = K <= 4 ? 4 : K <= 8 ? 8 : K <= 16 ? 16 : 32;
    const size_t fk = (size_t)F * K, nk = (size_t)N * K, fkt = (size_t)F * KT, nkt = (size_t)N * KT;
    // rows per tile = 32 x warps per block: 4..8 warps, whichever leaves the fewest idle warps in the last tile
    // (MNIST: 784 rows = 25 bit words -> 5 warps x 5 tiles)
    const int64_t f_words = (F + 31) / 32;
    int warps = GIBBS_WMAX;
    {
        int64_t best = -1;
        for (int w = GIBBS_WMAX; w >= 4; w--) {
            const int64_t waste = (f_words + w - 1) / w * w - f_words;
            if (best < 0 || waste < best) { best = waste; warps = w; }
        }
    }
    const int f_tiles = (int)((f_words + warps - 1) / warps);
    // column-side counts: per-row-tile partial buffers (plain stores) while they stay small, integer atomics beyond
    const bool cn_atomic = f_tiles > 16;
    DevBuf<int32_t> cF, cN, cNp, Zd;
    DevBuf<float> W, HH, wc, hc;
    DevBuf<double> accW, accH, accV, outb;
    DevBuf<int64_t> didx;
    CK(cF.alloc(fkt)); CK(cN.alloc(2 * nkt));
    if (!cn_atomic) CK(cNp.alloc((size_t)f_tiles * 2 * nkt));
    CK(W.alloc(fkt)); CK(HH.alloc(2 * nkt)); CK(wc.alloc(fk)); CK(hc.alloc(nk));
    CK(accW.alloc(fk)); CK(accH.alloc(nk)); CK(outb.alloc(std::max(fk, nk)));
    if (n_test > 0) { CK(accV.alloc((size_t)n_test)); CK(didx.alloc((size_t)n_test)); }
    if (Z_last) CK(Zd.alloc((size_t)F * N));
    unsigned long long *cnt = (unsigned long long *)h->scal + SC_TMP;
    const int nburnin = (int)std::floor(iter * burnin);
    h->timing = nbmf_timing{};
    CK(cudaMemsetAsync(cF, 0, fkt * 4, h->stream)); CK(cudaMemsetAsync(cN, 0, 2 * nkt * 4, h->stream));
    CK(cudaMemsetAsync(accW, 0, fk * 8, h->stream)); CK(cudaMemsetAsync(accH, 0, nk * 8, h->stream));
    CK(cudaMemsetAsync(cnt, 0, 16, h->stream));
    if (n_test > 0) {
        CK(cudaMemsetAsync(accV, 0, (size_t)n_test * 8, h->stream));
        CK(cudaMemcpyAsync(didx, test_idx, (size_t)n_test * 8, cudaMemcpyHostToDevice, h->stream));
    }
    // initial counts: from stored labels (compute_counts) when available, else from the fp64
    // statistics (which hold integer counts after a label initialisation)
    if (h->Z) {
        gibbs_counts_from_labels_kernel<<<nblk(F * N, 256), 256, 0, h->stream>>>(h->Z, h->vc, h->has_na ? h->mc : nullptr, F, N, h->Fw, K, KT, cF, cN, cnt + 1);
        LAUNCH_CHECK();
        unsigned long long bad = 0;
        CK(cudaMemcpyAsync(&bad, cnt + 1, 8, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (bad) return fail(h, NBMF_ERR_ARG, "Z contains labels outside 0..K-1 (BernoulliNMF.R:18-20 forces K_init=10)");
        int rc = nbmf_sum_ranks(h, cF, fkt, 1, h->stream);   // row counts are over all ranks' columns
        if (rc) return rc;
    } else if (h->statF && h->Kc == K) {
        gibbs_counts_from_stats_kernel<<<nblk((int64_t)std::max(fkt, 2 * nkt), 256), 256, 0, h->stream>>>(
            h->statF, h->statN, F, N, K, h->Kp, KT, cF, cN);
        LAUNCH_CHECK();
    } else return fail(h, NBMF_ERR_STATE, "Gibbs needs labels: call nbmf_init_from_labels or nbmf_init_random_labels first");

    GibbsDrawArgs d{};
    d.F = F; d.N = N; d.n_offset = h->opts.n_offset; d.K = K; d.KT = KT;
    d.alpha = alpha; d.beta = beta; d.gamma = gamma; d.seed = seed;
    d.cF = cF; d.cN = cN; d.cNp = cNp; d.W = W; d.HH = HH; d.accW = accW; d.accH = accH; d.w_cur = wc; d.h_cur = hc;
    GibbsSweepArgs sw{};
    sw.vc = h->vc; sw.mc = h->has_na ? h->mc : nullptr; sw.F = F; sw.N = N; sw.Fw = h->Fw; sw.n_offset = h->opts.n_offset;
    sw.W = W; sw.HH = HH; sw.seed = seed; sw.K = K; sw.cF = cF;
    sw.cn_atomic = cn_atomic ? 1 : 0; sw.cNp = cn_atomic ? (int32_t *)cN : (int32_t *)cNp;
    // enough blocks for a few waves of 148 SMs, each amortising its row histograms over ct sub-tiles
    const int NBc = GIBBS_NB(KT);
    const int64_t n_tiles = (N + NBc - 1) / NBc;
    // sub-tiles per block: the grid should fill whole waves of resident blocks (a block costs its ct sub-tiles plus
    // about a third of one for its row histograms and weights)
    int ct = 1;
    {
        int occ = 0;
        switch (KT) {
        case 4: occ = gibbs_sweep_occupancy<4>(warps); break;
        case 8: occ = gibbs_sweep_occupancy<8>(warps); break;
        case 16: occ = gibbs_sweep_occupancy<16>(warps); break;
        default: occ = gibbs_sweep_occupancy<32>(warps); break;
        }
        cudaGetLastError();
        const int64_t cap = (int64_t)std::max(1, occ) * h->sm_count;
        double best = 0;
        for (int c = 1; c <= GIBBS_CT_MAX; c++) {
            const int64_t blocks = ((n_tiles + c - 1) / c) * f_tiles;
            const double cost = (double)((blocks + cap - 1) / cap) * (c + 0.35);
            if (c == 1 || cost < best) { best = cost; ct = c; }
        }
    }
    sw.ct = ct;
    const dim3 sgrid((unsigned)((n_tiles + ct - 1) / ct), (unsigned)f_tiles);
    const unsigned dgrid = nblk((F + (N + 32 / KT - 1) / (32 / KT)) * 32, 256);   // a warp per row, a warp per 32 / KT columns
    int kept = 0;
    int rc = NBMF_OK;
    CK(cudaEventRecord(h->ev[0], h->stream));
    bool parts_valid = false;      // cNp holds the last sweep's column-side counts
    for (int i = 1; i < iter; i++) {
        // counts of sweep i-1 (or the initial ones) -> means of a kept sweep, W and H of sweep i
        d.sweep = (uint32_t)i; d.draw = 1;
        d.accumulate = (i - 1 >= 1 && i - 1 >= nburnin) ? 1 : 0;
        d.n_parts = (parts_valid && !cn_atomic) ? f_tiles : 0;
        d.zero_cn = cn_atomic ? 1 : 0;
        gibbs_draw_kernel<<<dgrid, 256, 0, h->stream>>>(d);
        LAUNCH_CHECK();
        if (d.accumulate) {
            kept++;
            if (n_test > 0) {
                gibbs_ev_test_kernel<<<nblk(n_test, 256), 256, 0, h->stream>>>(didx, n_test, F, K, wc, hc, accV);
                LAUNCH_CHECK();
                h->timing.launches += 1;
            }
        }
        sw.sweep = (uint32_t)i; sw.Zout = (i == iter - 1) ? (int32_t *)Zd : nullptr;
        cudaError_t e;
        switch (KT) {
        case 4: e = launch_gibbs_sweep<4>(sw, sgrid, warps, h->stream); break;
        case 8: e = launch_gibbs_sweep<8>(sw, sgrid, warps, h->stream); break;
        case 16: e = launch_gibbs_sweep<16>(sw, sgrid, warps, h->stream); break;
        default: e = launch_gibbs_sweep<32>(sw, sgrid, warps, h->stream); break;
        }
        CK(e);
        parts_valid = true;
        h->timing.launches += 2;
        // column sharding: the row counts are the one thing exchanged per sweep
        if ((rc = nbmf_sum_ranks(h, cF, fkt, 1, h->stream))) return rc;
    }
    // the last sweep's counts
    if (iter > 1 && iter - 1 >= nburnin) {
        d.sweep = (uint32_t)iter; d.draw = 0; d.accumulate = 1;
        d.n_parts = (parts_valid && !cn_atomic) ? f_tiles : 0; d.zero_cn = 0;
        gibbs_draw_kernel<<<dgrid, 256, 0, h->stream>>>(d);
        LAUNCH_CHECK();
        kept++;
        if (n_test > 0) {
            gibbs_ev_test_kernel<<<nblk(n_test, 256), 256, 0, h->stream>>>(didx, n_test, F, K, wc, hc, accV);
            LAUNCH_CHECK();
        }
        h->timing.launches += 1 + (n_test > 0);
    }
    CK(cudaEventRecord(h->ev[6], h->stream));
    if (kept == 0) {
        // iter <= 1 or a burn-in that keeps nothing: the reference divides by zero kept samples; say so
        CK(cudaStreamSynchronize(h->stream));
        return fail(h, NBMF_ERR_ARG, "Gibbs: no sweep is kept (iter must be >= 2 and floor(iter*burnin) <= iter-1)");
    }
    const double sc = 1.0 / kept;
    if (E_W) {
        scale_rows_to_colmajor_kernel<<<nblk(F * K, 256), 256, 0, h->stream>>>(accW, F, K, sc, outb);
        LAUNCH_CHECK();
        CK(cudaMemcpyAsync(E_W, outb, fk * 8, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    if (E_H) {
        scale_rows_to_colmajor_kernel<<<nblk(N * K, 256), 256, 0, h->stream>>>(accH, N, K, sc, outb);
        LAUNCH_CHECK();
        CK(cudaMemcpyAsync(E_H, outb, nk * 8, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    if (E_V_test && n_test > 0) {
        CK(cudaMemcpyAsync(E_V_test, accV, (size_t)n_test * 8, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        for (int64_t t = 0; t < n_test; t++) E_V_test[t] *= sc;
    }
    if (Z_last) CK(cudaMemcpyAsync(Z_last, Zd, (size_t)F * N * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    float ms = 0; cudaEventElapsedTime(&ms, h->ev[0], h->ev[6]);
    h->timing.total_ms = ms; h->timing.iters = iter - 1;
    return NBMF_OK;
}

// K > 32: one 32 x 32 entry tile per block, two passes over K, shared-memory histograms
static int gibbs_dirber_wide(nbmf_handle *h, int K, double alpha, double beta, double gamma, int iter,
                             double burnin, uint64_t seed, double *E_W, double *E_H,
                             const int64_t *test_idx, int64_t n_test, double *E_V_test, int32_t *Z_last)
{
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
    // column sharding: the row counts n_total (F x K) are the one thing exchanged per sweep.  A
    // failed exchange ends the chain: carrying on with rank-local counts would return wrong means.
    int xrc = NBMF_OK;
    auto exchange_counts = [&]() -> cudaError_t {
        xrc = nbmf_sum_ranks(h, cF, fk, 1, h->stream);
        return xrc == NBMF_OK ? cudaGetLastError() : cudaErrorUnknown;
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
            e = cudaGetLastError();
            if (e == cudaSuccess) e = exchange_counts();
            if (e != cudaSuccess) break;
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
    if (xrc != NBMF_OK) return xrc;   // h->err holds the collective's message
    if (e != cudaSuccess) { h->err = std::string("gibbs: ") + cudaGetErrorString(e); cudaGetLastError(); return NBMF_ERR_CUDA; }
    return rc;
}

extern "C" int nbmf_gibbs_dirber(nbmf_handle *h, int K, double alpha, double beta, double gamma, int iter,
                                 double burnin, uint64_t seed, double *E_W, double *E_H,
                                 const int64_t *test_idx, int64_t n_test, double *E_V_test, int32_t *Z_last)
{
    if (!h) return NBMF_ERR_ARG;
    if (!h->vc) return fail(h, NBMF_ERR_STATE, "no data");
    if (K <= 0 || K > 256) return fail(h, NBMF_ERR_ARG, "Gibbs: K must be in 1..256");
    if (!(alpha > 0) || !(beta > 0) || !(gamma > 0)) return fail(h, NBMF_ERR_ARG, "priors must be > 0");
    if (iter < 2 || !(burnin >= 0) || std::floor(iter * burnin) > iter - 1)
        return fail(h, NBMF_ERR_ARG, "Gibbs: no sweep is kept (iter must be >= 2 and floor(iter*burnin) <= iter-1)");
    CK(cudaSetDevice(h->device));
    static const bool force_wide = [] { const char *v = getenv("NBMF_GIBBS_KERNEL"); return v && strcmp(v, "wide") == 0; }();
    if (K <= 32 && !force_wide)
        return gibbs_dirber_fast(h, K, alpha, beta, gamma, iter, burnin, seed, E_W, E_H, test_idx, n_test, E_V_test, Z_last);
    return gibbs_dirber_wide(h, K, alpha, beta, gamma, iter, burnin, seed, E_W, E_H, test_idx, n_test, E_V_test, Z_last);
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
    if (nbmf_sharded(h) && (h->hook || h->comm_world > 1)) return fail(h, NBMF_ERR_UNSUPPORTED, "the collapsed wavefront sweep is replicas-only (no column sharding)");
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

#include "nbmf_extra.cuh"
