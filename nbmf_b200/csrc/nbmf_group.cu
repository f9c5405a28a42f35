// nbmf_group.cu -- the collectives of the column-sharded path, owned by the library.
//
// "Multi-GPU is internal to the handle".  Two ways to get there:
//   * one process per GPU (torchrun, MPI, ...): every rank creates its handle, rank 0 asks for an id
//     (nbmf_comm_unique_id), ships its 128 bytes to the others by whatever means it has, and every
//     rank calls nbmf_comm_init_rank -- from then on the handle's iterations exchange the F x K
//     Dirichlet-side statistic themselves (ncclAllReduce over NVLink / NVSwitch on the handle's
//     priority side stream, overlapped with the cols pass);
//   * one process, several GPUs -- what the reference's single R process is
//     (src/RcppExports.cpp:259-274 -> VB_Aspect_Rcpp): nbmf_group_create builds one handle per
//     device, an NCCL communicator over them (ncclCommInitAll) and one host thread per device; the
//     nbmf_group_* calls shard the columns of V over the devices, run the per-handle entry points
//     on the workers and hand back one result, so the Rcpp body stays a single call.
#include "host_common.h"
#include "tensor_pass.h"
#include <nccl.h>
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define NCCLCK(h, call)                                                                   \
    do {                                                                                   \
        ncclResult_t r__ = (call);                                                         \
        if (r__ != ncclSuccess) {                                                          \
            (h)->err = std::string(#call) + ": " + ncclGetErrorString(r__);                \
            return NBMF_ERR_CUDA;                                                          \
        }                                                                                  \
    } while (0)

// ---------------------------------------------------------------------------
// per-handle communicator
// ---------------------------------------------------------------------------
extern "C" int nbmf_comm_unique_id(void *id128)
{
    if (!id128) return NBMF_ERR_ARG;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return NBMF_ERR_CUDA;
    memcpy(id128, &id, sizeof id);
    return NBMF_OK;
}

extern "C" int nbmf_comm_init_rank(nbmf_handle *h, const void *id128, int world, int rank)
{
    if (!h) return NBMF_ERR_ARG;
    if (!id128 || world < 1 || rank < 0 || rank >= world) return fail(h, NBMF_ERR_ARG, "nbmf_comm_init_rank: bad arguments");
    if (h->comm) return fail(h, NBMF_ERR_STATE, "the handle already has a communicator");
    CK(cudaSetDevice(h->device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclComm_t c = nullptr;
    NCCLCK(h, ncclCommInitRank(&c, world, id, rank));
    h->comm = (void *)c; h->comm_rank = rank; h->comm_world = world; h->comm_owned = true;
    tp_invalidate_data(h);   // exact row counts of the tensor path: to be summed over the new set of ranks
    return NBMF_OK;
}

void nbmf_comm_release(nbmf_handle *h)
{
    if (h->comm && h->comm_owned) ncclCommDestroy((ncclComm_t)h->comm);
    h->comm = nullptr; h->comm_owned = false; h->comm_world = 1; h->comm_rank = 0;
}

extern "C" int nbmf_comm_destroy(nbmf_handle *h)
{
    if (!h) return NBMF_ERR_ARG;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    nbmf_comm_release(h);
    tp_invalidate_data(h);
    return NBMF_OK;
}

extern "C" int nbmf_comm_info(const nbmf_handle *h, int *rank, int *world)
{
    if (!h) return NBMF_ERR_ARG;
    if (rank) *rank = h->comm ? h->comm_rank : 0;
    if (world) *world = h->comm ? h->comm_world : 1;
    return NBMF_OK;
}

__global__ void grp_i32_to_f64_kernel(const int32_t *__restrict__ a, int64_t n, double *__restrict__ b)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = (double)a[i];
}
__global__ void grp_f64_to_i32_kernel(const double *__restrict__ a, int64_t n, int32_t *__restrict__ b)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = (int32_t)llrint(a[i]);
}
__global__ void grp_f32_to_f64_kernel(const float *__restrict__ a, int64_t n, double *__restrict__ b)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = (double)a[i];
}
__global__ void grp_f64_to_f32_kernel(const double *__restrict__ a, int64_t n, float *__restrict__ b)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = (float)a[i];
}

int nbmf_sum_ranks(nbmf_handle *h, void *buf, size_t count, int kind, cudaStream_t st)
{
    if (h->comm) {
        if (h->comm_world <= 1) return NBMF_OK;
        const ncclDataType_t dt = kind == 0 ? ncclDouble : kind == 1 ? ncclInt32 : ncclFloat;
        NCCLCK(h, ncclAllReduce(buf, buf, count, dt, ncclSum, (ncclComm_t)h->comm, st));
        return NBMF_OK;
    }
    if (!h->hook) return NBMF_OK;
    if (kind == 0) {
        if (h->hook(h->hook_ctx, buf, count, (void *)st) != 0) return fail(h, NBMF_ERR_STATE, "allreduce hook failed");
        return NBMF_OK;
    }
    // the hook sums fp64: other types take a round trip through a scratch buffer
    if (h->hook_scratch_bytes < count * 8) {
        cudaStreamSynchronize(st);
        cudaFree(h->hook_scratch); h->hook_scratch = nullptr; h->hook_scratch_bytes = 0;
        if (cudaMalloc(&h->hook_scratch, count * 8) != cudaSuccess) { cudaGetLastError(); return fail(h, NBMF_ERR_OOM, "exchange scratch allocation failed"); }
        h->hook_scratch_bytes = count * 8;
    }
    double *tmp = (double *)h->hook_scratch;
    const unsigned nb = (unsigned)((count + 255) / 256);
    if (kind == 1) grp_i32_to_f64_kernel<<<nb, 256, 0, st>>>((const int32_t *)buf, (int64_t)count, tmp);
    else grp_f32_to_f64_kernel<<<nb, 256, 0, st>>>((const float *)buf, (int64_t)count, tmp);
    if (h->hook(h->hook_ctx, (void *)tmp, count, (void *)st) != 0) return fail(h, NBMF_ERR_STATE, "allreduce hook failed");
    if (kind == 1) grp_f64_to_i32_kernel<<<nb, 256, 0, st>>>(tmp, (int64_t)count, (int32_t *)buf);
    else grp_f64_to_f32_kernel<<<nb, 256, 0, st>>>(tmp, (int64_t)count, (float *)buf);
    if (cudaGetLastError() != cudaSuccess) return fail(h, NBMF_ERR_CUDA, "exchange conversion kernels failed");
    return NBMF_OK;
}

int nbmf_max_ranks_u64(nbmf_handle *h, unsigned long long *buf, size_t count, cudaStream_t st)
{
    if (!h->comm || h->comm_world <= 1) return NBMF_OK;
    NCCLCK(h, ncclAllReduce(buf, buf, count, ncclUint64, ncclMax, (ncclComm_t)h->comm, st));
    return NBMF_OK;
}

int nbmf_bcast_root(nbmf_handle *h, void *buf, size_t bytes, cudaStream_t st)
{
    if (!h->comm || h->comm_world <= 1) return NBMF_OK;
    NCCLCK(h, ncclBroadcast(buf, buf, bytes, ncclChar, 0, (ncclComm_t)h->comm, st));
    return NBMF_OK;
}

// several arrays with the same row partition in ONE grouped NCCL launch (bufs[i] has rows of row_bytes[i] bytes)
int nbmf_allgather_rows_multi(nbmf_handle *h, void *const *bufs, const size_t *row_bytes, int n_bufs, int64_t rows_total, cudaStream_t st)
{
    if (!h->comm || h->comm_world <= 1) return NBMF_OK;
    const int G = h->comm_world;
    const int64_t per = (rows_total + G - 1) / G;
    const bool even = per * G == rows_total;             // equal blocks: one all-gather per array; else one broadcast per owner
    NCCLCK(h, ncclGroupStart());
    for (int i = 0; i < n_bufs; i++) {
        char *base = (char *)bufs[i];
        if (even) {
            ncclResult_t e = ncclAllGather(base + (size_t)h->comm_rank * per * row_bytes[i], base, (size_t)per * row_bytes[i], ncclChar,
                                           (ncclComm_t)h->comm, st);
            if (e != ncclSuccess) { ncclGroupEnd(); h->err = std::string("ncclAllGather: ") + ncclGetErrorString(e); return NBMF_ERR_CUDA; }
            continue;
        }
        for (int r = 0; r < G; r++) {
            int64_t lo, hi;
            nbmf_row_block(rows_total, G, r, &lo, &hi);
            if (hi <= lo) continue;
            char *p = base + (size_t)lo * row_bytes[i];
            ncclResult_t e = ncclBroadcast(p, p, (size_t)(hi - lo) * row_bytes[i], ncclChar, r, (ncclComm_t)h->comm, st);
            if (e != ncclSuccess) { ncclGroupEnd(); h->err = std::string("ncclBroadcast: ") + ncclGetErrorString(e); return NBMF_ERR_CUDA; }
        }
    }
    NCCLCK(h, ncclGroupEnd());
    return NBMF_OK;
}

int nbmf_allgather_rows(nbmf_handle *h, void *buf, size_t row_bytes, int64_t rows_total, cudaStream_t st)
{
    void *bufs[1] = {buf};
    return nbmf_allgather_rows_multi(h, bufs, &row_bytes, 1, rows_total, st);
}

// ---------------------------------------------------------------------------
// single-process group: one handle, one worker thread per device
// ---------------------------------------------------------------------------
struct nbmf_group {
    std::vector<nbmf_handle *> hs;
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    std::function<int(int)> job;
    uint64_t gen = 0;
    int pending = 0;
    bool stop = false;
    std::vector<int> rcs;
    std::string err;
    int64_t F = 0, N = 0;                 // global shape of the data the group holds
    std::vector<int64_t> n_lo;            // column block [n_lo[r], n_lo[r+1]) of rank r

    // run `fn(rank)` on every worker thread, return the first non-zero status
    int run(std::function<int(int)> fn)
    {
        std::unique_lock<std::mutex> lk(mu);
        job = std::move(fn);
        pending = (int)hs.size();
        gen++;
        cv_go.notify_all();
        cv_done.wait(lk, [&] { return pending == 0; });
        for (size_t r = 0; r < hs.size(); r++)
            if (rcs[r] != 0) { err = "device " + std::to_string(hs[r]->device) + ": " + hs[r]->err; return rcs[r]; }
        return NBMF_OK;
    }
    void worker(int r)
    {
        cudaSetDevice(hs[r]->device);
        uint64_t seen = 0;
        for (;;) {
            std::function<int(int)> fn;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_go.wait(lk, [&] { return stop || gen != seen; });
                if (stop) return;
                seen = gen;
                fn = job;
            }
            const int rc = fn(r);
            {
                std::unique_lock<std::mutex> lk(mu);
                rcs[r] = rc;
                if (--pending == 0) cv_done.notify_all();
            }
        }
    }
};

static void group_shard(nbmf_group *g, int64_t N)
{
    const int G = (int)g->hs.size();
    g->n_lo.assign(G + 1, 0);
    // contiguous blocks whose boundaries are multiples of 64 columns (an owner tile of the cols
    // pass, two steps of the rows pass, sixteen Philox groups of the Gibbs sweep)
    for (int r = 0; r <= G; r++) g->n_lo[r] = std::min<int64_t>(N, ((N * r / G) + 63) / 64 * 64);
    g->n_lo[G] = N;
    for (int r = 0; r < G; r++) {
        g->hs[r]->opts.n_global = G > 1 ? N : 0;
        g->hs[r]->opts.n_offset = g->n_lo[r];
    }
}

extern "C" int nbmf_group_create(nbmf_group **out, const nbmf_opts *opts, const int *devices, int n_devices)
{
    if (!out) return NBMF_ERR_ARG;
    *out = nullptr;
    if (!devices || n_devices < 1 || n_devices > 64) return NBMF_ERR_ARG;
    nbmf_group *g = new (std::nothrow) nbmf_group();
    if (!g) return NBMF_ERR_OOM;
    for (int i = 0; i < n_devices; i++) {
        nbmf_opts o{};
        if (opts) o = *opts;
        o.device = devices[i];
        nbmf_handle *h = nullptr;
        const int rc = nbmf_create(&h, &o);
        if (rc) {
            for (auto *x : g->hs) nbmf_destroy(x);
            delete g;
            return rc;
        }
        g->hs.push_back(h);
    }
    if (n_devices > 1) {
        std::vector<ncclComm_t> comms(n_devices);
        if (ncclCommInitAll(comms.data(), n_devices, devices) != ncclSuccess) {
            for (auto *x : g->hs) nbmf_destroy(x);
            delete g;
            return NBMF_ERR_CUDA;
        }
        for (int i = 0; i < n_devices; i++) {
            g->hs[i]->comm = (void *)comms[i]; g->hs[i]->comm_rank = i; g->hs[i]->comm_world = n_devices;
            g->hs[i]->comm_owned = true;
        }
    }
    g->rcs.assign(n_devices, 0);
    for (int i = 0; i < n_devices; i++) g->workers.emplace_back(&nbmf_group::worker, g, i);
    *out = g;
    return NBMF_OK;
}

extern "C" void nbmf_group_destroy(nbmf_group *g)
{
    if (!g) return;
    {
        std::unique_lock<std::mutex> lk(g->mu);
        g->stop = true;
        g->cv_go.notify_all();
    }
    for (auto &t : g->workers) t.join();
    for (auto *h : g->hs) nbmf_destroy(h);
    delete g;
}

extern "C" int nbmf_group_size(const nbmf_group *g) { return g ? (int)g->hs.size() : 0; }
extern "C" nbmf_handle *nbmf_group_handle(nbmf_group *g, int rank)
{
    return (g && rank >= 0 && rank < (int)g->hs.size()) ? g->hs[rank] : nullptr;
}
extern "C" const char *nbmf_group_last_error(const nbmf_group *g) { return g ? g->err.c_str() : "null group"; }
extern "C" int nbmf_group_columns(const nbmf_group *g, int rank, int64_t *n_lo, int64_t *n_hi)
{
    if (!g || rank < 0 || rank >= (int)g->hs.size() || g->n_lo.empty()) return NBMF_ERR_ARG;
    if (n_lo) *n_lo = g->n_lo[rank];
    if (n_hi) *n_hi = g->n_lo[rank + 1];
    return NBMF_OK;
}

extern "C" int nbmf_group_generate_synthetic(nbmf_group *g, int64_t F, int64_t N, int Ktrue, double a, double b, uint64_t seed)
{
    if (!g) return NBMF_ERR_ARG;
    g->F = F; g->N = N;
    group_shard(g, N);
    return g->run([&](int r) {
        return nbmf_generate_synthetic(g->hs[r], F, g->n_lo[r + 1] - g->n_lo[r], Ktrue, a, b, seed);
    });
}

extern "C" int nbmf_group_set_data_i32(nbmf_group *g, const int32_t *V, int64_t F, int64_t N)
{
    if (!g || !V) return NBMF_ERR_ARG;
    g->F = F; g->N = N;
    group_shard(g, N);
    return g->run([&](int r) {   // a column block of a column-major matrix is contiguous
        return nbmf_set_data_i32(g->hs[r], V + g->n_lo[r] * F, F, g->n_lo[r + 1] - g->n_lo[r]);
    });
}

extern "C" int nbmf_group_init_from_labels(nbmf_group *g, const int32_t *Z, int Kc)
{
    if (!g || !Z || g->n_lo.empty()) return NBMF_ERR_ARG;
    return g->run([&](int r) { return nbmf_init_from_labels(g->hs[r], Z + g->n_lo[r] * g->F, Kc); });
}

extern "C" int nbmf_group_init_random_labels(nbmf_group *g, int Kc, uint64_t seed)
{
    if (!g) return NBMF_ERR_ARG;
    return g->run([&](int r) { return nbmf_init_random_labels(g->hs[r], Kc, seed); });
}

extern "C" int nbmf_group_vb_begin(nbmf_group *g, int K, double alpha, double beta, double gamma, int mean_params)
{
    if (!g) return NBMF_ERR_ARG;
    return g->run([&](int r) { return nbmf_vb_begin(g->hs[r], K, alpha, beta, gamma, mean_params); });
}
extern "C" int nbmf_group_vb_step(nbmf_group *g, int checklb)
{
    if (!g) return NBMF_ERR_ARG;
    return g->run([&](int r) { return nbmf_vb_step(g->hs[r], checklb); });
}
extern "C" int nbmf_group_synchronize(nbmf_group *g)
{
    if (!g) return NBMF_ERR_ARG;
    return g->run([&](int r) { return nbmf_synchronize(g->hs[r]); });
}

// E_W F x K (from rank 0: the row side is replicated), E_H N x K (every rank its rows), the ELBO
// trace summed over the ranks' column terms plus the replicated row term
static int group_vb_end(nbmf_group *g, int K, double *E_W, double *E_H, double *lowerbounds, int n_lb)
{
    const int G = (int)g->hs.size();
    const int64_t N = g->N;
    std::vector<std::vector<double>> eh(G), terms(G);
    int rc = g->run([&](int r) {
        const int64_t nl = g->n_lo[r + 1] - g->n_lo[r];
        if (E_H) eh[r].resize((size_t)nl * K);
        nbmf_handle *h = g->hs[r];
        const int iters = h->vb_iter;
        int rc2 = nbmf_vb_end(h, r == 0 ? E_W : nullptr, E_H ? eh[r].data() : nullptr, nullptr, 0, nullptr);
        if (rc2) return rc2;
        if (lowerbounds && n_lb > 0) {
            terms[r].assign((size_t)3 * n_lb, 0.0);
            rc2 = nbmf_vb_get_lb_terms(h, terms[r].data(), std::min(n_lb, iters));
        }
        return rc2;
    });
    if (rc) return rc;
    if (E_H)
        for (int r = 0; r < G; r++) {
            const int64_t nl = g->n_lo[r + 1] - g->n_lo[r];
            for (int k = 0; k < K; k++)
                memcpy(E_H + (size_t)k * N + g->n_lo[r], eh[r].data() + (size_t)k * nl, (size_t)nl * 8);
        }
    if (lowerbounds)
        for (int i = 0; i < n_lb; i++) {
            double s = terms[0].empty() ? 0.0 : terms[0][3 * i + 1];          // pW - qW: replicated
            for (int r = 0; r < G; r++)
                if (!terms[r].empty()) s += terms[r][3 * i] + terms[r][3 * i + 2];   // sum log D and pH - qH of the block
            lowerbounds[i] = s;
        }
    return NBMF_OK;
}

extern "C" int nbmf_group_vb_end(nbmf_group *g, double *E_W, double *E_H, double *lowerbounds, int n_lb)
{
    if (!g || g->n_lo.empty()) return NBMF_ERR_ARG;
    return group_vb_end(g, g->hs[0]->vb_K, E_W, E_H, lowerbounds, n_lb);
}

// VB_Aspect_Rcpp(V, Z, K, alpha, beta, gamma, iter, checklb) (collapsed_gibbs_z_VB.cpp:373-526) over
// all devices of the group: what the Rcpp body of a multi-GPU build calls once.
extern "C" int nbmf_group_vb_aspect(nbmf_group *g, const int32_t *V, const int32_t *Z, int64_t F, int64_t N, int K,
                                    double alpha, double beta, double gamma, int iter, int checklb,
                                    double *E_W, double *E_H, double *lowerbounds)
{
    if (!g) return NBMF_ERR_ARG;
    if (!V || !Z || F <= 0 || N <= 0 || iter < 0) { g->err = "nbmf_group_vb_aspect: bad arguments"; return NBMF_ERR_ARG; }
    int rc = nbmf_group_set_data_i32(g, V, F, N);
    if (rc) return rc;
    if ((rc = nbmf_group_init_from_labels(g, Z, K))) return rc;
    if ((rc = g->run([&](int r) {
            nbmf_handle *h = g->hs[r];
            h->horizon_next = std::max(1, iter);
            int rc2 = nbmf_vb_begin(h, K, alpha, beta, gamma, 0);
            for (int i = 0; i < iter && rc2 == 0; i++) rc2 = nbmf_vb_step(h, checklb);
            if (rc2) h->vb_active = false;
            return rc2;
        })))
        return rc;
    if (lowerbounds && !checklb) { for (int i = 0; i < iter; i++) lowerbounds[i] = 0.0; }
    return group_vb_end(g, K, E_W, E_H, checklb ? lowerbounds : nullptr, iter);
}

// the same with V as host bit planes and the vbcounts as host arrays (nbmf_vb_aspect_bits): every
// device uploads its own column block, pipelined with its first iteration; n_total goes up once, on
// device 0, and reaches the others over NVLink; E_W comes back from device 0 only
extern "C" int nbmf_group_vb_aspect_bits(nbmf_group *g, const uint32_t *vbits, int64_t F, int64_t N, int K,
                                         double alpha, double beta, double gamma, int iter, int checklb,
                                         const double *n_total, const double *f_ones, const double *f_zeros,
                                         double *E_W, double *E_H, double *lowerbounds)
{
    if (!g) return NBMF_ERR_ARG;
    if (!vbits || !n_total || !f_ones || !f_zeros || F <= 0 || N <= 0 || iter < 0) { g->err = "nbmf_group_vb_aspect_bits: bad arguments"; return NBMF_ERR_ARG; }
    g->F = F; g->N = N;
    group_shard(g, N);
    const int G = (int)g->hs.size();
    const int64_t fw = (F + 31) / 32;
    // the column-side arrays are N x K column-major on the host: a rank's rows are strided there, so
    // each worker packs its block (host memcpy of nl*K doubles, tiny beside the bit plane)
    std::vector<std::vector<double>> fo(G), fz(G), eh(G), lb(G);
    int rc = g->run([&](int r) {
        nbmf_handle *h = g->hs[r];
        const int64_t c0 = g->n_lo[r], nl = g->n_lo[r + 1] - c0;
        fo[r].resize((size_t)nl * K); fz[r].resize((size_t)nl * K);
        if (E_H) eh[r].resize((size_t)nl * K);
        for (int k = 0; k < K; k++) {
            memcpy(fo[r].data() + (size_t)k * nl, f_ones + (size_t)k * N + c0, (size_t)nl * 8);
            memcpy(fz[r].data() + (size_t)k * nl, f_zeros + (size_t)k * N + c0, (size_t)nl * 8);
        }
        lb[r].assign((size_t)std::max(1, iter), 0.0);
        int rc2 = nbmf_vb_aspect_bits(h, vbits + c0 * fw, F, nl, K, alpha, beta, gamma, iter, checklb,
                                      r == 0 ? n_total : nullptr, fo[r].data(), fz[r].data(),
                                      r == 0 ? E_W : nullptr, E_H ? eh[r].data() : nullptr, nullptr);
        if (rc2) return rc2;
        if (lowerbounds && checklb) {
            std::vector<double> t3((size_t)3 * std::max(1, iter), 0.0);
            rc2 = nbmf_vb_get_lb_terms(h, t3.data(), iter);
            for (int i = 0; i < iter; i++) lb[r][i] = t3[3 * i] + t3[3 * i + 2] + (r == 0 ? t3[3 * i + 1] : 0.0);
        }
        return rc2;
    });
    if (rc) return rc;
    if (E_H)
        for (int r = 0; r < G; r++) {
            const int64_t c0 = g->n_lo[r], nl = g->n_lo[r + 1] - c0;
            for (int k = 0; k < K; k++) memcpy(E_H + (size_t)k * N + c0, eh[r].data() + (size_t)k * nl, (size_t)nl * 8);
        }
    if (lowerbounds)
        for (int i = 0; i < iter; i++) {
            double s = 0.0;
            if (checklb) for (int r = 0; r < G; r++) s += lb[r][i];
            lowerbounds[i] = s;
        }
    return NBMF_OK;
}

// Gibbs_DirBer over the group: the uncollapsed chain with the int32 row counts exchanged per sweep;
// draws are keyed by the global column, so the chain does not depend on the number of devices
extern "C" int nbmf_group_gibbs_dirber(nbmf_group *g, int K, double alpha, double beta, double gamma, int iter,
                                       double burnin, uint64_t seed, double *E_W, double *E_H)
{
    if (!g || g->n_lo.empty()) return NBMF_ERR_ARG;
    const int G = (int)g->hs.size();
    const int64_t N = g->N;
    std::vector<std::vector<double>> eh(G);
    int rc = g->run([&](int r) {
        const int64_t nl = g->n_lo[r + 1] - g->n_lo[r];
        if (E_H) eh[r].resize((size_t)nl * K);
        return nbmf_gibbs_dirber(g->hs[r], K, alpha, beta, gamma, iter, burnin, seed, r == 0 ? E_W : nullptr,
                                 E_H ? eh[r].data() : nullptr, nullptr, 0, nullptr, nullptr);
    });
    if (rc) return rc;
    if (E_H)
        for (int r = 0; r < G; r++) {
            const int64_t c0 = g->n_lo[r], nl = g->n_lo[r + 1] - c0;
            for (int k = 0; k < K; k++) memcpy(E_H + (size_t)k * N + c0, eh[r].data() + (size_t)k * nl, (size_t)nl * 8);
        }
    return NBMF_OK;
}
