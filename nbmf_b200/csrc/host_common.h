// host_common.h -- small host-side helpers shared by the translation units of libnbmf_cuda.
#pragma once
#include "common.cuh"
#include <algorithm>

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

static inline int fail(nbmf_handle *h, int code, const char *msg)
{
    if (h) h->err = msg;
    return code;
}

// true when this handle holds one column block of a larger V (collectives are in play)
static inline bool nbmf_sharded(const nbmf_handle *h) { return h->comm != nullptr || h->hook != nullptr; }

// sum `count` elements over the ranks, in place, ordered on `st`.  kind: 0 fp64, 1 int32, 2 fp32.
// The hook of nbmf_set_allreduce_hook only knows fp64.  (nbmf_group.cu)
int nbmf_sum_ranks(nbmf_handle *h, void *buf, size_t count, int kind, cudaStream_t st);
// maximum over the ranks of `count` uint64 values (operand exponents of the tensor path)
int nbmf_max_ranks_u64(nbmf_handle *h, unsigned long long *buf, size_t count, cudaStream_t st);
int nbmf_bcast_root(nbmf_handle *h, void *buf, size_t bytes, cudaStream_t st);
// every rank owns rows [r * ceil(rows/G), ...) of `buf` (rows of row_bytes bytes): make the array whole on all ranks
int nbmf_allgather_rows(nbmf_handle *h, void *buf, size_t row_bytes, int64_t rows_total, cudaStream_t st);
int nbmf_allgather_rows_multi(nbmf_handle *h, void *const *bufs, const size_t *row_bytes, int n_bufs, int64_t rows_total, cudaStream_t st);
// the row block of rank `rank` under that partition
static inline void nbmf_row_block(int64_t rows_total, int world, int rank, int64_t *lo, int64_t *hi)
{
    const int64_t per = (rows_total + world - 1) / world;
    *lo = std::min<int64_t>(rows_total, (int64_t)rank * per);
    *hi = std::min<int64_t>(rows_total, *lo + per);
}
void nbmf_comm_release(nbmf_handle *h);   // destroys a communicator created by nbmf_comm_init_rank
