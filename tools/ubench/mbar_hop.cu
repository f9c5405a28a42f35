// Micro-benchmark: latency of mbarrier hand-offs between warps of one CTA on sm_100a.
//   mode 0: thread A arrive -> thread B try_wait wake -> B arrive -> A try_wait wake (ping-pong)
//   mode 1: same, but A signals with tcgen05.commit (no MMAs outstanding)
//   mode 2: A: arrive, B: 128 threads arrive (count 128), i.e. a 4-warp group answering
//   mode 3: successful try_wait on an already completed phase (cost of the instruction itself)
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o mbar_hop mbar_hop.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_test_spin(uint32_t bar, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }

__global__ void hop_kernel(int mode, int iters, long long *out)
{
    __shared__ __align__(8) unsigned long long bars[4];
    __shared__ uint32_t tmem_base;
    const uint32_t b0 = smem_u32(&bars[0]), b1 = smem_u32(&bars[1]);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool spin = mode >= 10;
    if (spin) mode -= 10;
    if (threadIdx.x == 0) {
        mbar_init(b0, 1);
        mbar_init(b1, mode == 2 ? 128 : 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    long long t0 = clock64();
    if (mode == 3) {
        if (threadIdx.x == 0) {
            mbar_arrive(b0);               // complete phase 0
            for (int i = 0; i < iters; i++) mbar_wait(b0, 0);
            out[0] = (clock64() - t0) / iters;
        }
    } else if (warp == 0) {
        if (lane == 0) {
            uint32_t ph = 0;
            for (int i = 0; i < iters; i++) {
                if (mode == 1) tc_commit(b0); else mbar_arrive(b0);
                if (spin) mbar_test_spin(b1, ph); else mbar_wait(b1, ph);
                ph ^= 1;
            }
            out[0] = (clock64() - t0) / iters;
        }
    } else if (warp >= 4) {
        if (mode == 2 || (warp == 4 && lane == 0)) {
            uint32_t ph = 0;
            for (int i = 0; i < iters; i++) {
                if (spin) mbar_test_spin(b0, ph); else mbar_wait(b0, ph);
                ph ^= 1;
                mbar_arrive(b1);
            }
        }
    }
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem_base) : "memory");
}

int main()
{
    long long *d, h;
    cudaMalloc(&d, 8);
    const int modes[] = {0, 1, 2, 3, 10, 11, 12};
    for (int m : modes) {
        for (int rep = 0; rep < 2; rep++) {
            hop_kernel<<<1, 256>>>(m, 20000, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d: %s\n", m, cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("mode %2d: %lld cycles per round trip (or per try_wait for mode 3)\n", m, h);
    }
    // the same with all 148 SMs busy
    for (int m : modes) {
        hop_kernel<<<148, 256>>>(m, 20000, d);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("mode %2d (148 CTAs): %lld\n", m, h);
    }
    return 0;
}
