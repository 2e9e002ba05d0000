// Measures the tensor-pipe rate of the exact tcgen05.mma forms the Hamming engines issue (B200, sm_100a):
//   kind::mxf4.block_scale.block32, cta_group::2, M = 256 (128 per CTA), K = 64, N = 240 (A, B in shared memory) and
//   N = 160 with A in tensor memory.  One elected thread per CTA pair issues MMAs back to back into two alternating
//   accumulator stages - no loads, no epilogue - so the time is the tensor pipe's.  Operand contents are irrelevant.
// Output: one JSON line with TFLOP/s per variant (2 * M * N * K FLOP per MMA).  bench.py reads profiles/tensor_peaks.json.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_peak tools/mma_peak.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int N, bool TA>
__global__ void __launch_bounds__(128, 1) mma_peak_kernel(int iters)
{
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint32_t tmem_ptr;
    __shared__ uint64_t bar;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = raw + ((1024u - (raw & 1023u)) & 1023u);
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_ptr;
    {   // unit block scales in the last 32 columns, all 128 lanes
        const uint32_t t = tb + ((uint32_t)(warp * 32) << 16) + 480u;
        const uint32_t v = 0x7F7F7F7Fu;
        for (int h = 0; h < 2; ++h)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
                         ::"r"(t + 16u * h), "r"(v) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (rank == 0 && threadIdx.x == 0) {
        constexpr uint32_t idesc = (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (1u << 23) | ((uint32_t)(256 >> 4) << 24);
        const uint64_t hi = (uint64_t)0x40004040u << 32;
        const uint64_t adesc = hi | (uint64_t)(((base) >> 4) | 0x10000u);
        const uint64_t bdesc = hi | (uint64_t)(((base + 32768u) >> 4) | 0x10000u);
        const uint32_t sfa = tb + 480u, sfb = tb + 496u;
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = tb + (uint32_t)((i & 1) * N);
            if constexpr (TA) {
                const uint32_t a = tb + 2u * N + (uint32_t)((i & 3) * 8);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.block32 [%0], [%1], %2, %3, [%5], [%6], p;\n\t}"
                             ::"r"(d), "r"(a), "l"(bdesc + (uint64_t)(2 * (i & 3))), "r"(idesc), "r"(1u), "r"(sfa), "r"(sfb) : "memory");
            } else {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
                             ::"r"(d), "l"(adesc + (uint64_t)(2 * (i & 3))), "l"(bdesc + (uint64_t)(2 * (i & 3))), "r"(idesc), "r"(1u), "r"(sfa), "r"(sfb) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u), "r"(1000000u) : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
    }
}

template <int N, bool TA>
static double run(int iters)
{
    const size_t smem = 1024 + 65536 + 32768;
    cudaFuncSetAttribute(mma_peak_kernel<N, TA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; ++w) cudaLaunchKernelEx(&cfg, mma_peak_kernel<N, TA>, iters);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        cudaLaunchKernelEx(&cfg, mma_peak_kernel<N, TA>, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(e)); return -1.0; }
    const double flop = 74.0 * (double)iters * 2.0 * 256.0 * N * 64.0;
    return flop / (best * 1e-3) / 1e12;
}

int main()
{
    const int iters = 40000;
    const double a = run<240, false>(iters);
    const double b = run<160, true>(iters);
    const double c = run<160, false>(iters);
    printf("{\"mxf4_2cta_n240_ss_tflops\": %.1f, \"mxf4_2cta_n160_ts_tflops\": %.1f, \"mxf4_2cta_n160_ss_tflops\": %.1f, \"iters\": %d}\n", a, b, c, iters);
    return 0;
}
