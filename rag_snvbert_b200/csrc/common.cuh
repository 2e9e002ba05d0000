// Shared device/host helpers for libsnvknn (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

namespace snv {

// ---- error plumbing -------------------------------------------------------------------
void set_error(const std::string& msg);
extern long long g_launch_count;  // kernels launched by this library (snv_launch_count)
// measurement hook: events around the dominant kernel of a search (snv_profile_enable)
void profile_begin(cudaStream_t stream);
void profile_end(cudaStream_t stream);

#define SNV_CUDA_CHECK(expr)                                                             \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            ::snv::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));        \
            return SNV_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

#define SNV_LAUNCH_CHECK()                                                               \
    do {                                                                                 \
        ++::snv::g_launch_count;                                                         \
        cudaError_t _e = cudaGetLastError();                                             \
        if (_e != cudaSuccess) {                                                         \
            ::snv::set_error(std::string("kernel launch: ") + cudaGetErrorString(_e));   \
            return SNV_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

constexpr int kNumSMs = 148;  // B200

#ifdef __CUDACC__
// ---- mbarrier / bulk-copy (TMA 1-D) PTX --------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// try_wait suspends the warp in hardware until the phase completes or the time hint (ns) expires, so a waiting
// warp does not poll: it takes no issue slots from the other warps of its SM sub-partition while it waits.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
#ifdef TC_DEBUG_SPIN_WAIT
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
#else
    while (!mbar_try_wait(bar, parity)) {
    }
#endif
}

// For waits that are normally long and not latency critical (a producer running ahead of its
// ring): back off between polls so the spinning warp does not take issue slots from its SM sub-partition.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) __nanosleep(128);
}

// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP); dst/src 16-byte aligned,
// bytes a multiple of 16; completion is signalled on `bar` as transaction bytes.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                         uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
#endif  // __CUDACC__

}  // namespace snv
