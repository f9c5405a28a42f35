// Micro-benchmark behind the two-pass decision of DESIGN.md section 3: how fast can 148 CTAs add fp32
// tiles into a shared global accumulator?  A one-pass contraction (both statistics from one visit of each
// entry, as collapsed_gibbs_z_VB.cpp:447-465 does) keeps one side's accumulator in TMEM and has to add the
// other side's 128 x 128 fp32 partial (64 KB) into global memory once per 128 x 128 tile of entries, i.e.
// every ~1.3 us per SM at the tensor rate of the two-pass kernel: 148 x 64 KB / 1.3 us = 7.3 TB/s of
// reduction payload.  This program measures what the memory system sustains:
//   mode 0  red.global.add.f32        one element per thread, coalesced (128 B per warp instruction)
//   mode 1  red.global.add.v4.f32     four elements per thread (sm_90+ vector reduction)
//   mode 2  cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32   16 KB bulk reductions from shared memory
//   mode 3  plain st.global.v4.f32 of the same bytes (the no-reduction ceiling)
// Every CTA adds `tiles` 64 KB tiles; tile t of CTA c lands at offset ((c * 7 + t * 13) % slots) * 64 KB of a
// region of `region_mb` MB, so that different CTAs hit the same addresses at different times (as the f-blocks
// of a one-pass sweep would) and the region may or may not fit the 126 MB L2.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o red_throughput red_throughput.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int TILE_FLOATS = 128 * 128;

__global__ void __launch_bounds__(512) red_kernel(float *dst, int slots, int tiles, int mode)
{
    extern __shared__ __align__(128) float tile[];   // 64 KB
    for (int i = threadIdx.x; i < TILE_FLOATS; i += blockDim.x) tile[i] = 1.0f;
    __syncthreads();
    for (int t = 0; t < tiles; t++) {
        float *base = dst + (size_t)(((size_t)blockIdx.x * 7 + (size_t)t * 13) % slots) * TILE_FLOATS;
        if (mode == 0) {
            for (int i = threadIdx.x; i < TILE_FLOATS; i += blockDim.x)
                asm volatile("red.global.add.f32 [%0], %1;" ::"l"(base + i), "f"(tile[i]) : "memory");
        } else if (mode == 1) {
            for (int i = threadIdx.x * 4; i < TILE_FLOATS; i += blockDim.x * 4) {
                const float4 v = *reinterpret_cast<const float4 *>(tile + i);
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(base + i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
            }
        } else if (mode == 2) {
            if (threadIdx.x < 4) {
                const uint32_t src = (uint32_t)__cvta_generic_to_shared(tile + threadIdx.x * 4096);
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                             ::"l"(base + threadIdx.x * 4096), "r"(src), "r"(16384) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
        } else {
            for (int i = threadIdx.x * 4; i < TILE_FLOATS; i += blockDim.x * 4)
                *reinterpret_cast<float4 *>(base + i) = *reinterpret_cast<const float4 *>(tile + i);
        }
    }
    if (mode == 2 && threadIdx.x < 4) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char **argv)
{
    int dev = 0;
    cudaSetDevice(dev);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, dev);
    const int ctas = p.multiProcessorCount;
    const int tiles = argc > 1 ? atoi(argv[1]) : 2048;
    cudaFuncSetAttribute(red_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_FLOATS * 4);
    const char *names[4] = {"red.global.add.f32", "red.global.add.v4.f32", "cp.reduce.async.bulk add.f32 (16 KB)", "st.global.v4.f32 (no reduction)"};
    printf("%s, %d SMs, %d tiles of 64 KB per CTA\n", p.name, ctas, tiles);
    for (int region_mb : {32, 512}) {
        const int slots = region_mb * 16;   // 64 KB slots
        float *d = nullptr;
        if (cudaMalloc(&d, (size_t)slots * TILE_FLOATS * 4) != cudaSuccess) { printf("alloc failed\n"); return 1; }
        cudaMemset(d, 0, (size_t)slots * TILE_FLOATS * 4);
        for (int mode = 0; mode < 4; mode++) {
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            red_kernel<<<ctas, 512, TILE_FLOATS * 4>>>(d, slots, 16, mode);   // warm-up
            cudaEventRecord(e0);
            red_kernel<<<ctas, 512, TILE_FLOATS * 4>>>(d, slots, tiles, mode);
            cudaEventRecord(e1);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            const double bytes = (double)ctas * tiles * TILE_FLOATS * 4;
            printf("region %4d MB  %-40s %8.1f GB/s of payload  (%.2f us per 64 KB tile and CTA)\n", region_mb, names[mode],
                   bytes / (ms * 1e-3) / 1e9, ms * 1e3 / tiles);
            cudaEventDestroy(e0); cudaEventDestroy(e1);
        }
        cudaFree(d);
    }
    printf("need for a one-pass contraction at the two-pass kernel's tensor rate: 148 x 64 KB / 1.3 us = 7300 GB/s\n");
    return 0;
}
