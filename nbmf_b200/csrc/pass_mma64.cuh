// pass_mma64.cuh -- the FP64 contraction pass on the FP64 tensor pipe (DMMA, mma.sync m8n8k4).
//
// Same "owner / stream" form as pass_simt_kernel (kernels_basic.cuh) and the tcgen05 pass
// (tensor_pass.cu), for Kp in {8, 16, 32, 64, 128}:
//   S[l][c] = u_l . w_c                     64 owner lanes x 64 streamed rows per block step   (GEMM 1)
//   R[l][c] = sel(l, c) / S[l][c]           in the accumulator fragments, written to shared memory
//   G[l][:] += sum_c R[l][c] w_c            accumulators stay in registers over the whole sweep (GEMM 2)
// Every arithmetic operation is an IEEE fp64 FMA / division, so this path keeps the 1e-10 parity of
// the FP64 mode (VB_Aspect_Rcpp's q(Z) loop, collapsed_gibbs_z_VB.cpp:447-465, with the E_Z cube
// eliminated -- SURVEY.md Appendix A); it replaces the shuffle-reduction kernel, which ran at 2.6 %
// of the FP64 peak, as the K <= 32 product path and as the validator of the f16 tensor path.
//
// Both branch rows of an entry are multiplied (the unselected one meets R = 0), as on the f16 path:
// selecting the operand row per entry would cost two 32-bit selects per FMA on the ALU pipe.
//
// Block = 256 threads (8 warps).  Shared memory: U tile [64][Kp+4], W tile [64][Kp+4], R tile [64][68]
// (fp64; the +4 padding makes every fragment load / store conflict-free at two wavefronts per 256 B).
// The next W tile is prefetched into registers while GEMM 2 runs.
#pragma once
#include "common.cuh"

struct Mma64Args {
    const double *U, *W;            // [L][Kp], [S_rows][Kp]
    double *G;                      // [L][Kp]
    const uint32_t *vbits, *mbits;  // [L / ob][words], packed along the streamed entity (mbits may be NULL)
    int64_t words, L;
    int64_t S_rows;                 // streamed rows (= entities * sb)
    int64_t rows_per_split;         // multiple of 64
    int ob, sb;                     // rows pass: ob = 1, sb = 2; cols pass: ob = 2, sb = 1
    double *sumlogD;                // += sum over selected pairs of log S (NULL: not wanted)
};

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int KP> struct Mma64Cfg {
    static constexpr int TL = 64, TS = 64;
    static constexpr int LDU = KP + 4, LDW = KP + 4, LDR = TS + 4;
    static constexpr int NTK = KP / 8;                    // k tiles of GEMM 2
    static constexpr int WK = NTK >= 4 ? 4 : NTK;         // warps along k
    static constexpr int WL = 8 / WK;                     // warps along the owner lanes
    static constexpr int MI = 8 / WL;                     // l tiles per warp
    static constexpr int NI = NTK / WK;                   // k tiles per warp
    static constexpr int PF = (TS * KP) / (256 * 2);      // double2 prefetch registers per thread
    static constexpr size_t SMEM = (size_t)(TL * LDU + TS * LDW + TL * LDR) * 8 + 2 * TL * 2 * 4;
    static_assert(KP % 8 == 0 && KP >= 8 && KP <= 128, "Kp");
    static_assert((TS * KP) % 512 == 0, "prefetch layout");
};

template <int KP, bool ROWS, bool WANT_LOG>
__global__ void __launch_bounds__(256) pass_mma64_kernel(Mma64Args a)
{
    using C = Mma64Cfg<KP>;
    extern __shared__ double sm64[];
    double *Us = sm64;                                   // [TL][LDU]
    double *Ws = Us + C::TL * C::LDU;                    // [TS][LDW]
    double *Rs = Ws + C::TS * C::LDW;                    // [TL][LDR]
    uint32_t *sbv = (uint32_t *)(Rs + C::TL * C::LDR);   // [TL][2] value words of the current step
    uint32_t *sbm = sbv + C::TL * 2;                     // [TL][2] mask words
    __shared__ double sh_red[32];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int64_t l0 = (int64_t)blockIdx.x * C::TL;
    const int64_t r_begin = (int64_t)blockIdx.y * a.rows_per_split;
    const int64_t r_end = min(a.S_rows, r_begin + a.rows_per_split);
    constexpr int WPS = ROWS ? 1 : 2;                    // bit words per lane and step (32 n or 64 f)

    // owner tile
    for (int i = tid; i < C::TL * KP; i += 256) {
        const int r = i / KP, k = i % KP;
        Us[r * C::LDU + k] = (l0 + r < a.L) ? a.U[(l0 + r) * KP + k] : 0.0;
    }
    // GEMM 1 tiling: warp = 4 l tiles x 2 c tiles; GEMM 2 tiling: warp = MI l tiles x NI k tiles
    const int w1l = (warp >> 2) * 32, w1c = (warp & 3) * 16;
    const int w2l = (warp / C::WK) * (C::MI * 8), w2k = (warp % C::WK) * (C::NI * 8);
    double acc[C::MI][C::NI][2];
#pragma unroll
    for (int i = 0; i < C::MI; i++)
#pragma unroll
        for (int j = 0; j < C::NI; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    double ell = 0.0;

    double2 pf[C::PF];
    auto prefetch = [&](int64_t r0) {
#pragma unroll
        for (int j = 0; j < C::PF; j++) {
            const int e = (tid + 256 * j) * 2;
            const int64_t row = r0 + e / KP;
            pf[j] = (row < r_end) ? *reinterpret_cast<const double2 *>(a.W + row * KP + (e % KP)) : make_double2(0.0, 0.0);
        }
    };
    auto commit = [&]() {
#pragma unroll
        for (int j = 0; j < C::PF; j++) {
            const int e = (tid + 256 * j) * 2;
            *reinterpret_cast<double2 *>(Ws + (e / KP) * C::LDW + (e % KP)) = pf[j];
        }
    };
    if (r_begin < r_end) prefetch(r_begin);
    for (int64_t r0 = r_begin; r0 < r_end; r0 += C::TS) {
        __syncthreads();               // GEMM 2 of the previous step is done with Ws and Rs
        commit();
        if (tid < C::TL * WPS) {       // this step's bit words of the 64 owner lanes
            const int li = tid / WPS, wi = tid % WPS;
            const int64_t l = l0 + li;
            uint32_t v = 0, m = 0;
            if (l < a.L) {
                const int64_t e0 = r0 / a.sb;                       // first streamed entity of the step
                const int64_t idx = (l / a.ob) * a.words + (e0 >> 5) + wi;
                v = a.vbits[idx];
                m = a.mbits ? a.mbits[idx] : 0xffffffffu;
            }
            sbv[li * 2 + wi] = v; sbm[li * 2 + wi] = m;
        }
        __syncthreads();
        if (r0 + C::TS < r_end) prefetch(r0 + C::TS);
        // ---- GEMM 1: S = U W^T
        double s[4][2][2];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 2; j++) s[i][j][0] = s[i][j][1] = 0.0;
#pragma unroll 4
        for (int k0 = 0; k0 < KP; k0 += 4) {
            double fa[4], fb[2];
#pragma unroll
            for (int i = 0; i < 4; i++) fa[i] = Us[(w1l + i * 8 + g) * C::LDU + k0 + t];
#pragma unroll
            for (int j = 0; j < 2; j++) fb[j] = Ws[(w1c + j * 8 + g) * C::LDW + k0 + t];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 2; j++) dmma_m8n8k4(s[i][j][0], s[i][j][1], fa[i], fb[j]);
        }
        // ---- R = sel / S in the accumulator layout: element e of tile (i, j) is
        //      lane w1l + 8 i + g, streamed row w1c + 8 j + 2 t + e
        double dprod = 1.0;   // product of the step's selected D (16 per thread, each >= ~1e-9): one logarithm per step
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int li = w1l + i * 8 + g;
            const bool own_b = ((l0 + li) & 1) != 0;                // cols pass: branch of this owner lane
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int c = w1c + j * 8 + 2 * t;
                double r[2];
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    // rows pass: streamed row c = 2 n + b -> entity c / 2, branch e; cols pass: entity c + e
                    const int ent = ROWS ? (c >> 1) : (c + e);
                    const uint32_t vw = sbv[li * 2 + (ent >> 5)], mw = sbm[li * 2 + (ent >> 5)];
                    const bool vb = (vw >> (ent & 31)) & 1u;
                    bool act = ((mw >> (ent & 31)) & 1u) && (vb == (ROWS ? (e == 1) : own_b));
                    act = act && (r0 + c + e < r_end) && (l0 + li < a.L);
                    const double d = s[i][j][e];
                    r[e] = act ? 1.0 / d : 0.0;
                    if (WANT_LOG) dprod *= act ? d : 1.0;
                }
                *reinterpret_cast<double2 *>(Rs + li * C::LDR + c) = make_double2(r[0], r[1]);
            }
            if (WANT_LOG && dprod < 1e-200) { ell += log(dprod); dprod = 1.0; }   // four factors between checks: no underflow for D > 1e-27
        }
        if (WANT_LOG) ell += log(dprod);
        __syncthreads();
        // ---- GEMM 2: G += R W
#pragma unroll 4
        for (int c0 = 0; c0 < C::TS; c0 += 4) {
            double fa[C::MI], fb[C::NI];
#pragma unroll
            for (int i = 0; i < C::MI; i++) fa[i] = Rs[(w2l + i * 8 + g) * C::LDR + c0 + t];
#pragma unroll
            for (int j = 0; j < C::NI; j++) fb[j] = Ws[(c0 + t) * C::LDW + w2k + j * 8 + g];
#pragma unroll
            for (int i = 0; i < C::MI; i++)
#pragma unroll
                for (int j = 0; j < C::NI; j++) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
        }
    }
    // ---- accumulators: lane w2l + 8 i + g, components w2k + 8 j + 2 t + {0, 1}
#pragma unroll
    for (int i = 0; i < C::MI; i++) {
        const int64_t l = l0 + w2l + i * 8 + g;
        if (l < a.L) {
#pragma unroll
            for (int j = 0; j < C::NI; j++) {
                double *dst = a.G + l * KP + w2k + j * 8 + 2 * t;
                if (gridDim.y == 1) *reinterpret_cast<double2 *>(dst) = make_double2(acc[i][j][0], acc[i][j][1]);
                else { atomicAdd(dst, acc[i][j][0]); atomicAdd(dst + 1, acc[i][j][1]); }
            }
        }
    }
    if (WANT_LOG) {
        const double tsum = block_sum(ell, sh_red);
        if (tid == 0) atomicAdd(a.sumlogD, tsum);
    }
}
