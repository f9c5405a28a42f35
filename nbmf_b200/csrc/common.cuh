// common.cuh -- shared device helpers and the handle layout of libnbmf_cuda.
//
// Device data layout (HBM), code orientation of the reference (V is F x N):
//   vc / mc : value / mask bit planes packed along f, one row of Fw 32-bit words
//             per column n   (owner = column n, streamed index = f)
//   vr / mr : the same bits packed along n, Nw words per row f
//             (owner = row f, streamed index = n)
//   statF   : [F][Kp]   fp64  n_total                (vbcounts::n_total,
//             collapsed_gibbs_z_VB.cpp:14-35)
//   statN   : [2N][Kp]  fp64  row 2n+1 = f_ones(n,:), row 2n = f_zeros(n,:)
//   A       : [F][Kp]   fp64  Dirichlet-side operand  exp(E log W)  (:428-434)
//   BB      : [2N][Kp]  fp64  row 2n+1 = exp(E log H_n), row 2n = exp(E log(1-H_n))
//             (:437-444)
// Kp is K padded to the k-chunking of the contraction kernel; padded entries
// are 0 in operands and statistics.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include "../../include/nbmf_cuda.h"

#define NBMF_WARP 32

struct nbmf_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    nbmf_opts opts{};
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;

    int64_t F = 0, N = 0, nobs = 0;
    int64_t Fw = 0, Nw = 0;  // words per column / per row (padded to 8 words = 256 bits)
    uint32_t *vc = nullptr, *mc = nullptr, *vr = nullptr, *mr = nullptr;
    bool has_na = false;

    int Kc = 0, Kp = 0, KC = 0, NKC = 0;
    int Kp_alloc_K = 0;                          // K the model buffers were sized for (Kc = 0 marks 'statistics stale')
    double *statF = nullptr, *statN = nullptr;   // current statistics
    double *A = nullptr, *BB = nullptr;          // operands of the current iteration
    double *GF = nullptr, *GN = nullptr;         // raw contraction outputs
    double *EW = nullptr, *EH = nullptr;         // [F][Kp], [N][Kp] expectations of the last params
    double *scal = nullptr;                      // device scalars, see SC_* below
    double *lbs = nullptr;                       // device ELBO trace
    int lbs_cap = 0;
    int32_t *Z = nullptr;                        // labels, column-major f + F*n (optional)
    void *stage = nullptr; size_t stage_bytes = 0;
    void *io = nullptr; size_t io_bytes = 0;     // staging of host <-> device matrix transfers (layout change on the device)
    cudaStream_t out_stream = nullptr;           // early download of E_W / E_H beside the last iteration
    cudaEvent_t out_ev = nullptr;
    double *early_EW = nullptr, *early_EH = nullptr;   // host destinations, armed for step `early_iter`
    int early_iter = -1; bool early_done = false;

    // VB session (begin/step/end)
    bool vb_active = false;
    int vb_K = 0, vb_iter = 0, vb_mean_params = 0;
    double vb_alpha = 1, vb_beta = 1, vb_gamma = 1;

    // Row slicing of a sharded tensor-path session (library communicator only): the Dirichlet side is
    // replicated data, but from the second iteration on each rank computes parameters, operand split and
    // statistics for rows [row_lo, row_hi) only and the ranks exchange the f16 operands (all-gather); the
    // fp64 arrays are made whole again by nbmf_vb_end
    bool vb_sliced = false;
    int64_t row_lo = 0, row_hi = 0;

    // contraction variant of the current session (nbmf_path), fixed by resolve_path()
    int path_mode = 0;
    int vb_horizon = 0;                          // VB updates planned from this start (AUTO precision)
    int horizon_next = 0;                        // set by the single-call entry points for their nbmf_vb_begin

    // tensor path state (tensor_pass.cu)
    void *tp = nullptr;

    nbmf_allreduce_fn hook = nullptr; void *hook_ctx = nullptr;
    // collectives owned by the library (nbmf_comm_init_rank / nbmf_group_create): an NCCL communicator
    // over the ranks that share the columns of V; takes precedence over the hook
    void *comm = nullptr;                        // ncclComm_t
    int comm_rank = 0, comm_world = 1;
    bool comm_owned = false;
    void *hook_scratch = nullptr; size_t hook_scratch_bytes = 0;   // fp64 staging of non-fp64 exchanges through the hook
    nbmf_timing timing{};
    cudaEvent_t ev[8] = {};
    // exchange overlapped with the cols pass: side stream + begin / end events
    cudaStream_t xchg_stream = nullptr;
    cudaEvent_t xchg_ev[2] = {};
    cudaEvent_t xchg_t[2] = {};                  // timing events around the collective on the side stream
    bool xchg_pending = false, xchg_timed = false;
    // pipelined upload (nbmf_vb_aspect_bits): copy stream, one event per column block
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t pipe_ev[33] = {};
    int64_t pipe_n0[34] = {};
    int pipe_blocks = 0;                   // > 0 while the first step of a pipelined call is pending
};

enum { SC_SUMLOGD = 0, SC_PRIOR_ROWS = 1, SC_PRIOR_COLS = 2, SC_NEG = 3, SC_TMP = 4, SC_COUNT = 8 };

#define NBMF_CUDA_CHECK(h, call)                                                          \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            (h)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                \
            return NBMF_ERR_CUDA;                                                          \
        }                                                                                  \
    } while (0)

static inline int64_t round_up64(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// ---------------------------------------------------------------------------
// device math
// ---------------------------------------------------------------------------
__device__ __forceinline__ double dev_digamma(double x)
{
    // recurrence to x >= 10 then the asymptotic series; same scheme as the
    // oracle so that small-K parity is ~1e-15 (boost::math::digamma stand-in,
    // collapsed_gibbs_z_VB.cpp:430-442)
    double r = 0.0;
    if (!(x > 0.0)) return (x == 0.0) ? -INFINITY : NAN;
    while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
    double inv = 1.0 / x, inv2 = inv * inv;
    double s = inv2 * (1.0 / 12.0 - inv2 * (1.0 / 120.0 - inv2 * (1.0 / 252.0 -
               inv2 * (1.0 / 240.0 - inv2 * (1.0 / 132.0 - inv2 * (691.0 / 32760.0 -
               inv2 * (1.0 / 12.0)))))));
    return r + log(x) - 0.5 * inv - s;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double block_sum(double v, double *sh /* >= 32 doubles */)
{
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
    if (w == 0) v = warp_sum(v);
    return v; // valid in warp 0
}

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter-based
// ---------------------------------------------------------------------------
struct philox4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                          uint32_t c3, uint32_t k0, uint32_t k1)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

__host__ __device__ __forceinline__ float u32_to_unif(uint32_t x) // (0,1)
{
    return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f);
}
__host__ __device__ __forceinline__ double u64_to_unif(uint32_t hi, uint32_t lo) // (0,1)
{
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return ((double)(v >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
