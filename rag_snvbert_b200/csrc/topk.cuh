// Register-resident exact top-k on totally ordered integer keys.
//
// Every selection in the engine (thread-local insert, cross-split merge, cross-GPU merge)
// compares ONE integer key that encodes (distance, row id), so the result order is the
// canonical (distance ascending, id ascending) by construction — which is what makes the
// Hamming path bit-exact against the oracle including ties.
#pragma once
#include <stdint.h>

namespace snv {

constexpr uint32_t kSent32 = 0xFFFFFFFFu;
constexpr uint64_t kSent64 = 0xFFFFFFFFFFFFFFFFull;

// best[] ascending; drop the largest, insert key.  best'[i] = max(best[i-1], min(best[i], key)).
template <int KT, typename T>
__device__ __forceinline__ void topk_insert(T (&best)[KT], T key)
{
#pragma unroll
    for (int i = KT - 1; i > 0; --i) {
        T lo = best[i] < key ? best[i] : key;
        best[i] = best[i - 1] > lo ? best[i - 1] : lo;
    }
    best[0] = best[0] < key ? best[0] : key;
}

// non-negative float -> order-preserving uint32 (IEEE bits are monotone for x >= 0)
__device__ __forceinline__ uint32_t f32_key(float x) { return __float_as_uint(x < 0.f ? 0.f : x); }

}  // namespace snv
