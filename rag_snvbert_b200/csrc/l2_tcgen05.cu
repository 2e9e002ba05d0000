// Float squared-L2 search:  D = |q|^2 + |r|^2 - 2 q.r  with the cross term on the 5th-gen tensor
// cores (tcgen05.mma kind::tf32, accumulators in TMEM) and the norm-add + clamp + exact top-k
// fused into the epilogue (sm_100a).
//
// Replaces faiss IndexFlatL2.search on float rows (src/dataset/rag_train_dataset.py:281 token
// vectors, src/dataset/embedding_rag_infer_dataset.py:284-285) and torch.cdist + topk
// (src/dataset/embedding_rag_dataset.py:392-402).
//
// Layout of one CTA (192 threads), canonical warp-specialised Blackwell GEMM:
//   warp 0      TMA producer: 2-D tiled loads (SWIZZLE_128B) of the query tile A [128 x 32] and the
//               panel tile B [256 x 32] fp32 per k-block into a kStages-deep shared-memory ring
//   warp 1      TMEM allocator + MMA issuer: one elected lane issues 4 x tcgen05.mma (M128 N256 K8)
//               per k-block into one of two 256-column accumulator stages; tcgen05.commit frees
//               the smem slot / publishes the accumulator
//   warps 2..9  epilogue: thread = query row (TMEM lane), two warps per lane quarter taking
//               alternate 32-column chunks; tcgen05.ld (double-buffered), d = fma(-2, acc, qn + rn),
//               clamp, candidate lists in smem folded in lockstep into a register top-k on
//               (float bits, id)
// A CTA owns one 128-query tile and a contiguous range of panel tiles, so the running top-k
// stays in registers across tiles; partial results go to a [nq][nsplit][kt] key buffer that
// merge_keys_kernel reduces (same total order as everywhere else).
//
// Precision: operands are fp32 bit patterns read as tf32 (10-bit mantissa).  Mode TF32X3 feeds
// the hi/lo split  A' = [q_lo | q_hi | q_hi],  B' = [r_hi | r_lo | r_hi]  so that one GEMM over
// K' = 3d yields q_lo.r_hi + q_hi.r_lo + q_hi.r_hi (the dropped lo.lo term and the truncation
// of lo are both ~2^-22 relative).  The tensor core's fp32 accumulation truncates (measured
// ~1 ulp of the running sum per MMA step), which bounds the distance error at about
// (d/8) ulp(|q.r|): the stated tolerance is 1e-5 (|q|^2 + |r|^2).  Small-integer inputs (tokens,
// genotypes) are exact in either mode.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "kernels.cuh"
#include "tcgen05.cuh"
#include "topk.cuh"

namespace snv {

namespace {

constexpr int BM = 128;        // queries per CTA tile (TMEM lanes)
constexpr int BN = 256;        // panel rows per MMA tile (TMEM columns per accumulator stage)
constexpr int BK = 32;         // fp32 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 8;      // tf32: 32 bytes per MMA
constexpr int kAccStages = 2;
constexpr int kTmemCols = kAccStages * BN;  // 512
constexpr int kEpiWarps = 8;    // two per TMEM lane quarter, alternating 32-column chunks
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = 64 + kEpiThreads;
constexpr uint32_t kABytes = BM * BK * 4;   // 16 KB
constexpr uint32_t kBBytes = BN * BK * 4;   // 32 KB
constexpr uint32_t kStageBytes = kABytes + kBBytes;
// Ring geometry per mode.  The k-loop is bound by bytes in flight (ring depth x stage bytes against the L2 latency), not
// by L2 bandwidth: single CTAs hold 3 stages of 48 KB; a CTA of a pair loads only half of every panel tile (32 KB per
// stage), so the same shared memory holds 5 k-blocks in flight.
template <bool PAIR>
struct Ring2 {
    static constexpr int kStages = PAIR ? 5 : 3;
    static constexpr uint32_t kStageBytes = PAIR ? kABytes + kBBytes / 2 : kABytes + kBBytes;
};
#ifndef SNV_L2_SLOT_EPI
#define SNV_L2_SLOT_EPI 1  // epilogue: per-column slots + 32-column masks (0: the round-1 per-thread candidate lists)
#endif
constexpr int kListCap = 24;   // per-thread candidate list (epilogue): folded when > 8 are pending, checked every 16 columns
constexpr size_t kListBytes = SNV_L2_SLOT_EPI ? (size_t)32 * kEpiThreads * 4 : (size_t)kListCap * kEpiThreads * 8;
template <bool PAIR>
constexpr size_t smem_bytes()
{
    return 1024 /*align slack*/ + (size_t)Ring2<PAIR>::kStages * Ring2<PAIR>::kStageBytes + 2 * BN * 4 /*rn*/ + kListBytes + 256 /*barriers*/;
}
static_assert(smem_bytes<false>() <= 232448 && smem_bytes<true>() <= 232448, "shared memory budget");

using namespace tc;

// f(std::integral_constant<int, 0>{}) ... f(std::integral_constant<int, 31>{}): compile-time column index
template <int I = 0, typename F>
__device__ __forceinline__ void static_for32(F&& f)
{
    if constexpr (I < 32) {
        f(std::integral_constant<int, I>{});
        static_for32<I + 1>(f);
    }
}

template <int KT, bool SPLITK, bool PAIR = false>
__global__ void __launch_bounds__(kThreads, 1)
l2_topk_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_r,
               const L2SearchParams p)
{
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment for SWIZZLE_128B tiles
    const uint32_t raw_addr = smem_u32(smem_raw);
    unsigned char* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    constexpr int kStages = Ring2<PAIR>::kStages;
    constexpr uint32_t kStageBytes = Ring2<PAIR>::kStageBytes;
    unsigned char* tiles = smem;
    float* rn_s = reinterpret_cast<float*>(smem + (size_t)kStages * kStageBytes);  // [2][BN]
    uint64_t* lists = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * kStageBytes + 2 * BN * 4);  // [kListCap][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * kStageBytes + 2 * BN * 4 + kListBytes);
    uint64_t* full_bar = bars;                          // [kStages]   TMA -> MMA
    uint64_t* empty_bar = bars + kStages;               // [kStages]   MMA -> TMA
    uint64_t* tmem_full = bars + 2 * kStages;           // [2]         MMA -> epilogue
    uint64_t* tmem_empty = bars + 2 * kStages + 2;      // [2]         epilogue -> MMA
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int n_tiles_total = (int)((p.n + BN - 1) / BN);
    // fused mode: CTA = (query tile, contiguous range of panel tiles), full depth.
    // split-K mode (skinny problems, e.g. 48 x 2008 x 197,760): CTA = (query tile, ONE panel tile,
    // a range of k-blocks); every CTA stores its partial dot products into its own slice and the
    // selection pass adds the slices in a fixed order (round-to-nearest, deterministic; the truncating
    // tensor-core accumulation chains stay short).
    int split, mt, t0, my_tiles, kb0, num_kb;
    if constexpr (SPLITK) {
        int b = blockIdx.x;
        const int ks = b % p.ksplit;
        b /= p.ksplit;
        t0 = b % n_tiles_total;
        mt = b / n_tiles_total;
        my_tiles = 1;
        split = 0;
        kb0 = ks * p.kb_per_split;
        const int total_kb = p.kp / BK;
        num_kb = (kb0 + p.kb_per_split < total_kb) ? p.kb_per_split : total_kb - kb0;
    } else {
        // CTA pair (PAIR): the two CTAs of a cluster take query tiles 2 m and 2 m + 1 against the same panel tiles;
        // each loads half (128 rows) of every panel tile and the leader issues one M = 256 MMA over both halves
        const int b = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
        split = b % p.nsplit;
        mt = PAIR ? 2 * (b / p.nsplit) + (int)cluster_ctarank() : b / p.nsplit;
        t0 = split * p.tiles_per_split;
        const int t1 = (t0 + p.tiles_per_split < n_tiles_total) ? t0 + p.tiles_per_split : n_tiles_total;
        my_tiles = t1 - t0;  // >= 1 by construction of the plan
        kb0 = 0;
        num_kb = p.kp / BK;
    }
    const int m0 = mt * BM;
    const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0u;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_q);
        prefetch_tensormap(&map_r);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < kAccStages; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], (PAIR ? 2 : 1) * kEpiThreads);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        if constexpr (PAIR) tmem_alloc_2cta(tmem_ptr, kTmemCols);
        else tmem_alloc(tmem_ptr, kTmemCols);
    }
    tcgen05_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync();  // the peer's barriers exist before anything arrives on them
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = 0; t < my_tiles; ++t) {
                const int n0 = (t0 + t) * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1u);
#ifdef L2_DEBUG_NO_TMA
                    mbar_arrive(&full_bar[stage]);
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                    continue;
#endif
                    unsigned char* a_dst = tiles + (size_t)stage * kStageBytes;
                    unsigned char* b_dst = a_dst + kABytes;
                    if constexpr (PAIR) {
                        // both CTAs' bytes (query tile + half panel tile each) complete on the leader's barrier
                        if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * (kABytes + kBBytes / 2));
                        tma_load_2d_2cta(a_dst, &map_q, (kb0 + kb) * BK, m0, &full_bar[stage]);
                        tma_load_2d_2cta(b_dst, &map_r, (kb0 + kb) * BK, n0 + (int)cta_rank * (BN / 2), &full_bar[stage]);
                    } else {
                        mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
                        tma_load_2d(a_dst, &map_q, (kb0 + kb) * BK, m0, &full_bar[stage]);
                        tma_load_2d(b_dst, &map_r, (kb0 + kb) * BK, n0, &full_bar[stage]);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (the leader CTA of a pair) =================
        constexpr uint32_t idesc = make_idesc_tf32(PAIR ? 2 * BM : BM, BN);
        if (leader) {
        int stage = 0;
        uint32_t phase = 0;
        for (int t = 0; t < my_tiles; ++t) {
            const int as = t & 1;
            const uint32_t acc_phase = (uint32_t)(t >> 1) & 1u;
            mbar_wait(&tmem_empty[as], acc_phase ^ 1u);
            tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tcgen05_fence_after();
                if (elect_one()) {
#ifndef L2_DEBUG_NO_MMA
                    const uint32_t a_addr = smem_u32(tiles + (size_t)stage * kStageBytes);
                    const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t adesc = make_kmajor_sw128_desc(a_addr + k * UMMA_K * 4);
                        const uint64_t bdesc = make_kmajor_sw128_desc(b_addr + k * UMMA_K * 4);
                        if constexpr (PAIR) umma_tf32_2cta(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
                        else umma_tf32(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
#endif
                    if constexpr (PAIR) {
                        umma_commit_2cta(&empty_bar[stage]);                   // both CTAs' slots reusable once read
                        if (kb == num_kb - 1) umma_commit_2cta(&tmem_full[as]);  // both CTAs' accumulators complete
                    } else {
                        umma_commit(&empty_bar[stage]);                   // smem slot reusable once read
                        if (kb == num_kb - 1) umma_commit(&tmem_full[as]);  // accumulator complete
                    }
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
        }
        }  // leader
    } else {
        // ================= epilogue: thread = query row, 2 warps per lane quarter =================
        if constexpr (SPLITK) {
            const int quarter = warp & 3;
            const int half = (warp - 2) >> 2;
            const int64_t q = (int64_t)m0 + quarter * 32 + lane;
            const int n0 = t0 * BN;
            mbar_wait(&tmem_full[0], 0);
            tcgen05_fence_after();
            const int ncols = (p.n - n0 < BN) ? (int)(p.n - n0) : BN;
            const int nchunks = (ncols + 31) >> 5;
            for (int ci = half; ci < nchunks; ci += 2) {
                uint32_t acc[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ci * 32), acc);
                tmem_ld_wait(acc);
                if (q < p.nq) {
                    // this CTA's own slice dot[k-split][q][n0 ..]: plain stores, summed in a fixed order afterwards
                    // (deterministic, unlike fp32 atomics; pad columns of the last tile are written too and never read)
                    float4* dst = reinterpret_cast<float4*>(p.dot + ((int64_t)(kb0 / p.kb_per_split) * p.nq + q) * p.dot_ld + n0 + ci * 32);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        dst[j] = make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]), __uint_as_float(acc[4 * j + 2]),
                                             __uint_as_float(acc[4 * j + 3]));
                }
            }
            tcgen05_fence_before();
            mbar_arrive(&tmem_empty[0]);
        } else {
        const int quarter = warp & 3;              // TMEM lanes [32*quarter, 32*quarter + 32)
        const int half = (warp - 2) >> 2;          // which alternate 32-column chunks this warp takes
        const int row = quarter * 32 + lane;       // row inside the 128-query tile
        const int et = (warp - 2) * 32 + lane;     // 0..255 among the epilogue threads
        const int64_t q = (int64_t)m0 + row;
        const bool active = q < p.nq;
        const float qn = active ? p.q_norm[q] : 0.f;
        uint64_t best[KT];
#pragma unroll
        for (int i = 0; i < KT; ++i) best[i] = kSent64;
#if SNV_L2_SLOT_EPI
        // Selection (slot + mask, the scheme of the tensor-core Hamming kernel's first generation): per column ONE fused
        // multiply-add  d' = |r|^2 - 2 q.r  and ONE compare against  thr - |q|^2  (so |q|^2 never enters the per-column
        // work), then - predicated - a store of d' into the column's own slot and a bit in a 32-column mask (four partial
        // masks; the bit is a multiply-add by a runtime 1: FMA pipe).  After each 32-column chunk the lanes pop their
        // mask bits in lockstep, two per round, and only then clamp, build the 64-bit key (float bits, id) and insert
        // into the sorted register top-k.  The compare threshold carries a few ulp of slack: the key compare decides.
        float* my_slot = reinterpret_cast<float*>(lists) + et;  // column j of the chunk at my_slot[j * kEpiThreads]
        float thr = 3.4028234663852886e38f, thrq = 3.4028234663852886e38f;
        const uint32_t one = (uint32_t)p.one;
        auto refresh = [&]() {
            thr = best[KT - 1] == kSent64 ? 3.4028234663852886e38f : __uint_as_float((uint32_t)(best[KT - 1] >> 32));
            const float t = thr - qn;
            thrq = best[KT - 1] == kSent64 ? 3.4028234663852886e38f : t + fabsf(t) * 9.5367431640625e-7f + 1e-30f;  // + 2^-20 relative
        };
        auto fold = [&](uint32_t mask, int col0) {
            if (__any_sync(0xffffffffu, mask != 0u)) {
                do {
                    if (mask != 0u) {
                        const int j1 = __ffs((int)mask) - 1;
                        mask &= mask - 1u;
                        const bool two = mask != 0u;
                        const int j2 = two ? __ffs((int)mask) - 1 : j1;
                        mask &= mask - 1u;
                        const float d1 = fmaxf(my_slot[j1 * kEpiThreads] + qn, 0.f);   // tiny negative round-off clamps to 0 like faiss
                        const float d2 = fmaxf(my_slot[j2 * kEpiThreads] + qn, 0.f);
                        const uint64_t key1 = ((uint64_t)__float_as_uint(d1) << 32) | (uint64_t)(uint32_t)(col0 + j1);
                        const uint64_t key2 = two ? ((uint64_t)__float_as_uint(d2) << 32) | (uint64_t)(uint32_t)(col0 + j2) : kSent64;
                        if (key1 < best[KT - 1]) topk_insert<KT, uint64_t>(best, key1);
                        if (key2 < best[KT - 1]) topk_insert<KT, uint64_t>(best, key2);
                    }
                } while (__any_sync(0xffffffffu, mask != 0u));
                refresh();
            }
        };
        auto process = [&](uint32_t (&acc)[32], const float* rn, int c0, int n0) {
            uint32_t m4[4] = {0u, 0u, 0u, 0u};
            const uint32_t slot0 = smem_u32(my_slot);
            const float4* rn4 = reinterpret_cast<const float4*>(rn + c0);
            float4 v = rn4[0];
            static_for32([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                const float r = (j & 3) == 0 ? v.x : ((j & 3) == 1 ? v.y : ((j & 3) == 2 ? v.z : v.w));
                const float dp = fmaf(-2.f, __uint_as_float(acc[j]), r);
                if constexpr ((j & 3) == 3 && j < 31) v = rn4[(j + 1) >> 2];
                if (dp < thrq) {
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(slot0 + (uint32_t)(j * kEpiThreads * 4)), "f"(dp) : "memory");
                    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(m4[j & 3]) : "r"(one), "n"(1u << j));
                }
            });
            fold((m4[0] | m4[1]) | (m4[2] | m4[3]), n0 + c0);
        };
        auto fold_final = [&]() {};
#else
        // Selection: candidates that beat the (possibly stale, hence looser) threshold are appended
        // to a per-thread list in shared memory; lists are folded into the sorted register top-k
        // only when some lane has more than 8 pending — all lanes insert in lockstep, instead of
        // one divergent ~50-instruction insertion per column.
        uint64_t* my_list = lists + et;  // slot s at my_list[s * kEpiThreads]
        float thr = 3.4028234663852886e38f;
        int cnt = 0;
        auto fold = [&]() {
            const int maxc = __reduce_max_sync(0xffffffffu, cnt);
            for (int s2 = 0; s2 < maxc; ++s2) {
                if (s2 < cnt) {
                    const uint64_t key = my_list[s2 * kEpiThreads];
                    if (key < best[KT - 1]) topk_insert<KT, uint64_t>(best, key);
                }
            }
            cnt = 0;
            thr = best[KT - 1] == kSent64 ? 3.4028234663852886e38f : __uint_as_float((uint32_t)(best[KT - 1] >> 32));
        };
        // 16 columns: distances, threshold test, append.  `bias` = |q|^2 + |r|^2 was read from shared
        // memory before any list store (the compiler cannot hoist shared loads across those stores).
        auto score16 = [&](const uint32_t* acc, const float* bias, int col0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float d = fmaf(-2.f, __uint_as_float(acc[j]), bias[j]);
                if (d < thr) {
                    // ids ascend along the scan, so on equal distance the earlier id stays;
                    // tiny negative round-off clamps to 0 like faiss
                    const float dc = d < 0.f ? 0.f : d;
                    my_list[cnt * kEpiThreads] = ((uint64_t)__float_as_uint(dc) << 32) | (uint64_t)(uint32_t)(col0 + j);
                    ++cnt;
                }
            }
            if (__any_sync(0xffffffffu, cnt > kListCap - 16)) fold();
        };
        auto process = [&](uint32_t (&acc)[32], const float* rn, int c0, int n0) {
            float bias[32];
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
                const float4 v = reinterpret_cast<const float4*>(rn + c0)[j4];
                bias[4 * j4] = qn + v.x; bias[4 * j4 + 1] = qn + v.y;
                bias[4 * j4 + 2] = qn + v.z; bias[4 * j4 + 3] = qn + v.w;
            }
            score16(acc, bias, n0 + c0);
            score16(acc + 16, bias + 16, n0 + c0 + 16);
        };
        auto fold_final = [&]() { fold(); };
#endif
        // |r|^2 of the first tile (1 per thread; +inf past the panel end so those columns never
        // pass the threshold); later tiles are prefetched one tile ahead
        float rn_next = (t0 * BN + et < p.n) ? p.ref_norm[t0 * BN + et] : __int_as_float(0x7f800000);
        for (int t = 0; t < my_tiles; ++t) {
            const int as = t & 1;
            const uint32_t acc_phase = (uint32_t)(t >> 1) & 1u;
            const int n0 = (t0 + t) * BN;
            float* rn = rn_s + as * BN;
            rn[et] = rn_next;
            if (t + 1 < my_tiles) {
                const int64_t c = (int64_t)n0 + BN + et;
                rn_next = c < p.n ? p.ref_norm[c] : __int_as_float(0x7f800000);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
            mbar_wait(&tmem_full[as], acc_phase);
            tcgen05_fence_after();
#ifdef L2_DEBUG_NO_EPILOGUE
            const int nchunks = 0;
#else
            const int ncols = (p.n - n0 < BN) ? (int)(p.n - n0) : BN;
            const int nchunks = (ncols + 31) >> 5;
#endif
            const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN);
            uint32_t accA[32], accB[32];
            int ci = half;
            if (ci < nchunks) tmem_ld_32x32b_x32(tbase + (uint32_t)(ci * 32), accA);
            while (ci < nchunks) {
                tmem_ld_wait(accA);
                if (ci + 2 < nchunks) tmem_ld_32x32b_x32(tbase + (uint32_t)((ci + 2) * 32), accB);
                process(accA, rn, ci * 32, n0);
                ci += 2;
                if (ci >= nchunks) break;
                tmem_ld_wait(accB);
                if (ci + 2 < nchunks) tmem_ld_32x32b_x32(tbase + (uint32_t)((ci + 2) * 32), accA);
                process(accB, rn, ci * 32, n0);
                ci += 2;
            }
            tcgen05_fence_before();
            if (PAIR && !leader) mbar_arrive_cluster(&tmem_empty[as], 0u);
            else mbar_arrive(&tmem_empty[as]);
        }
        fold_final();
        if (active) {
            uint64_t* out = p.partial + (((int64_t)q * p.nsplit + split) * 2 + half) * KT;
#pragma unroll
            for (int i = 0; i < KT; ++i) out[i] = best[i];
        }
        }  // fused mode
    }

    tcgen05_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync();  // the leader's MMAs may still read this CTA's operands
    if (warp == 1) {
        tcgen05_fence_after();
        if constexpr (PAIR) tmem_dealloc_2cta(tmem_base, kTmemCols);
        else tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---- split-K selection input: dot products -> ordered keys ----------------------------------
__global__ void __launch_bounds__(256)
l2_keys_kernel(const float* __restrict__ dot, int64_t dot_ld, int ksplit, const float* __restrict__ q_norm,
               const float* __restrict__ ref_norm, int64_t nq, int64_t n, uint64_t* __restrict__ keys)
{
    const int64_t total = nq * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = i / n, c = i % n;
        float dotv = 0.f;
        for (int ks = 0; ks < ksplit; ++ks) dotv += dot[((int64_t)ks * nq + q) * dot_ld + c];  // fixed order
        float d = fmaf(-2.f, dotv, q_norm[q] + ref_norm[c]);
        d = d < 0.f ? 0.f : d;
        keys[i] = ((uint64_t)__float_as_uint(d) << 32) | (uint64_t)(uint32_t)c;
    }
}

// ---- centering: column means of the first rows added (L2 is translation invariant) ------------
// Deterministic: every (column, block of rows) partial sum is produced by one thread in row order, and the partials of a
// column are added in block order (double accumulation) by one thread - no atomics, so two indexes built from the same
// rows centre identically, bit for bit.
constexpr int kMeanRowsPerBlock = 256;
__global__ void __launch_bounds__(256)
l2_colsum_kernel(const float* __restrict__ x, int64_t rows, int64_t d, float* __restrict__ parts)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d) return;
    const int64_t r0 = (int64_t)blockIdx.y * kMeanRowsPerBlock;
    const int64_t r1 = r0 + kMeanRowsPerBlock < rows ? r0 + kMeanRowsPerBlock : rows;
    float acc = 0.f;
    for (int64_t r = r0; r < r1; ++r) acc += x[r * d + c];
    parts[(int64_t)blockIdx.y * d + c] = acc;
}

__global__ void __launch_bounds__(256)
l2_colmean_reduce_kernel(const float* __restrict__ parts, int64_t nblocks, int64_t d, double inv_rows, float* __restrict__ mean)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d) return;
    double acc = 0.0;
    for (int64_t b = 0; b < nblocks; ++b) acc += (double)parts[b * d + c];
    mean[c] = (float)(acc * inv_rows);
}

// ---- operand preparation: one warp per (row, 2048-column chunk) ------------------------------
// (few, very long rows — 48 queries x 197,760 — must still fill the machine)
constexpr int kPrepChunk = 2048;
__global__ void __launch_bounds__(256)
l2_prep_kernel(const float* __restrict__ x, const float* __restrict__ mean, int64_t rows, int64_t d, int mode,
               bool is_query, int kp, int chunks_per_row, float* __restrict__ ops, float* __restrict__ norms,
               float* __restrict__ norm_parts)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t items = rows * chunks_per_row;
    const int64_t used = (mode == SNV_L2_TF32X3) ? 3 * d : d;
    for (int64_t item = warp; item < items; item += nwarps) {
        const int64_t r = item / chunks_per_row;
        const int64_t c0 = (item % chunks_per_row) * kPrepChunk;
        const int64_t c1 = c0 + kPrepChunk < d ? c0 + kPrepChunk : d;
        const float* xr = x + r * d;
        float* o = ops + r * kp;
        float acc = 0.f;
        for (int64_t c = c0 + lane; c < c1; c += 32) {
            const float v = mean ? xr[c] - mean[c] : xr[c];
            acc = fmaf(v, v, acc);
            if (mode == SNV_L2_TF32X3) {
                const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
                const float lo = v - hi;
                // small cross terms first, hi.hi last: the tensor core accumulates in fp32 with
                // truncation (~1 ulp of the running sum per MMA step), so the long chain at full
                // magnitude is kept as short as possible
                o[c] = is_query ? lo : hi;
                o[d + c] = is_query ? hi : lo;
                o[2 * d + c] = hi;
            } else {
                o[c] = v;
            }
        }
        if (c1 == d)
            for (int64_t c = used + lane; c < kp; c += 32) o[c] = 0.f;  // zero the k padding once per row
#pragma unroll
        for (int s2 = 16; s2 > 0; s2 >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s2);
        if (lane == 0) {
            if (chunks_per_row == 1) norms[r] = acc;
            else norm_parts[r * chunks_per_row + item % chunks_per_row] = acc;  // summed in order by l2_norm_reduce_kernel
        }
    }
}

// Are all values integral and small (tokens, genotypes)?  Their tf32 products are exact, so centering would only hurt;
// anything else (embeddings) gets the column means subtracted by default.  flag &= 0 on the first offender.
__global__ void __launch_bounds__(256)
l2_integral_check_kernel(const float* __restrict__ x, int64_t count, int* __restrict__ flag)
{
    bool ok = true;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        ok = ok && (v == rintf(v)) && fabsf(v) <= 2048.f;
    }
    if (!__all_sync(0xffffffffu, ok) && (threadIdx.x & 31) == 0) atomicAnd(flag, 0);
}

// |x|^2 of long rows: the per-chunk partial sums added in a fixed order (double accumulation), one thread per row
__global__ void __launch_bounds__(128)
l2_norm_reduce_kernel(const float* __restrict__ parts, int64_t rows, int chunks, float* __restrict__ norms)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    double acc = 0.0;
    for (int c = 0; c < chunks; ++c) acc += (double)parts[r * chunks + c];
    norms[r] = (float)acc;
}

// ---- host: tensor maps ---------------------------------------------------------------------
int make_map(CUtensorMap* map, const float* base, int64_t rows, int kp, int box_rows)
{
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return SNV_ERR_CUDA;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)kp, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)kp * 4};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
        return SNV_ERR_CUDA;
    }
    return SNV_OK;
}

}  // namespace

int l2_operand_depth(int64_t d, int mode)
{
    const int64_t k = (mode & 0xF) == SNV_L2_TF32X3 ? 3 * d : d;
    return (int)round_up(k, BK);
}

int l2_prep_chunks(int64_t d) { return (int)ceil_div(d, kPrepChunk); }

int l2_prep_launch(const float* x, const float* mean, int64_t rows, int64_t d, int mode, bool is_query, int kp,
                   float* ops, float* norms, float* norm_parts, cudaStream_t stream)
{
    if (rows <= 0) return SNV_OK;
    const int block = 256;
    const int chunks = l2_prep_chunks(d);
    if (chunks > 1 && !norm_parts) { set_error("l2_prep: rows deeper than one chunk need the partial-norm scratch"); return SNV_ERR_INVALID; }
    int64_t grid = ceil_div(rows * chunks, block / 32);
    if (grid > (int64_t)kNumSMs * 16) grid = (int64_t)kNumSMs * 16;
    l2_prep_kernel<<<(unsigned)grid, block, 0, stream>>>(x, mean, rows, d, mode & 0xF, is_query, kp, chunks, ops, norms, norm_parts);
    SNV_LAUNCH_CHECK();
    if (chunks > 1) {
        l2_norm_reduce_kernel<<<(unsigned)ceil_div(rows, 128), 128, 0, stream>>>(norm_parts, rows, chunks, norms);
        SNV_LAUNCH_CHECK();
    }
    return SNV_OK;
}

int l2_integral_check_launch(const float* x, int64_t count, int* flag, cudaStream_t stream)
{
    if (count <= 0) return SNV_OK;
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(count, 256 * 8), (int64_t)kNumSMs * 16);
    l2_integral_check_kernel<<<grid, 256, 0, stream>>>(x, count, flag);
    SNV_LAUNCH_CHECK();
    return SNV_OK;
}

size_t l2_colmean_scratch_bytes(int64_t rows, int64_t d) { return (size_t)ceil_div(std::max<int64_t>(rows, 1), kMeanRowsPerBlock) * (size_t)d * 4; }

int l2_colmean_launch(const float* x, int64_t rows, int64_t d, float* mean, float* scratch, cudaStream_t stream)
{
    if (rows <= 0) {
        SNV_CUDA_CHECK(cudaMemsetAsync(mean, 0, (size_t)d * 4, stream));
        return SNV_OK;
    }
    const int block = 256;
    const int64_t nblocks = ceil_div(rows, kMeanRowsPerBlock);
    dim3 grid((unsigned)ceil_div(d, block), (unsigned)nblocks);
    l2_colsum_kernel<<<grid, block, 0, stream>>>(x, rows, d, scratch);
    SNV_LAUNCH_CHECK();
    l2_colmean_reduce_kernel<<<(unsigned)ceil_div(d, block), block, 0, stream>>>(scratch, nblocks, d, 1.0 / (double)rows, mean);
    SNV_LAUNCH_CHECK();
    return SNV_OK;
}

size_t l2_plan(L2SearchParams& p)
{
    if (p.k < 1 || p.k > 32) {
        set_error("L2 search: k must be in [1, 32] (got " + std::to_string(p.k) + ")");
        return (size_t)-1;
    }
    if (p.n >= ((int64_t)1 << 32)) {
        set_error("L2 search: panel too large for 32-bit ids; shard rows");
        return (size_t)-1;
    }
    p.kt = p.k <= 8 ? 8 : 32;
    const int64_t n_tiles = p.n > 0 ? ceil_div(p.n, BN) : 0;
    const int64_t m_tiles = ceil_div(p.nq, BM);
    // CTA pairs (cta_group::2, needs two query tiles): each CTA of a pair streams only half of every panel tile, so its
    // ring holds 5 k-blocks in flight instead of 3 in the same shared memory.  Round 1 measured the pair kernel with the
    // single-CTA ring geometry (3 stages) and found no gain; with 5 stages cfg 4 runs 59.6 vs 65.9 us (tensor pipe 70 %
    // of active cycles), so pairs are the default wherever a window has two query tiles.
    p.pair = m_tiles >= 2;   // SNV_L2_PAIR=0 forces single CTAs
    if (const char* e = getenv("SNV_L2_PAIR")) p.pair = m_tiles >= 2 && atoi(e) != 0;
    const int64_t m_items = p.pair ? ceil_div(m_tiles, 2) : m_tiles;
    const int64_t units = p.pair ? kNumSMs / 2 : kNumSMs;
    // pick the row split that minimises (waves x tiles per CTA): one CTA per SM is resident
    int best_s = 1;
    int64_t best_cost = -1;
    for (int s = 1; s <= n_tiles && s <= 64; ++s) {
        const int64_t per = ceil_div(n_tiles, s);
        const int64_t real_s = ceil_div(n_tiles, per);
        const int64_t waves = ceil_div(m_items * real_s, units);
        const int64_t cost = waves * (per * 8 + 1);  // +1: fixed prologue/epilogue weight per CTA
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_s = (int)real_s; }
    }
    p.nsplit = n_tiles > 0 ? best_s : 1;
    p.tiles_per_split = n_tiles > 0 ? (int)ceil_div(n_tiles, p.nsplit) : 0;
    p.nsplit = n_tiles > 0 ? (int)ceil_div(n_tiles, p.tiles_per_split) : 1;
    // split-K when the (query tile x panel tile) grid cannot fill the machine and the depth is large
    p.ksplit = 1;
    p.kb_per_split = p.kp / BK;
    const int64_t total_kb = p.kp / BK;
    const int64_t ctas = m_tiles * n_tiles;
    if (n_tiles > 0 && ctas * 2 <= kNumSMs && total_kb >= 64 && p.nq * p.n <= ((int64_t)1 << 26)) {
        int64_t ks = ceil_div(2 * kNumSMs, ctas);
        const int64_t max_ks = total_kb / 16;  // at least 16 k-blocks (64 MMAs) per CTA
        if (ks > max_ks) ks = max_ks;
        // Short accumulation chains: the tensor core adds into its fp32 accumulator by truncation (~1 ulp of the running
        // sum per MMA step), so a k-split never runs more than 128 k-blocks (512 MMA steps); the slices are then added in
        // fp32 round-to-nearest.  At the reference's depth (197,760 x 3 operand columns) that is 145 slices.
        if (ks >= 2 && ceil_div(total_kb, ks) > 128) ks = ceil_div(total_kb, 128);
        if (ks >= 2) {
            p.kb_per_split = (int)ceil_div(total_kb, ks);
            p.ksplit = (int)ceil_div(total_kb, p.kb_per_split);
            p.dot_ld = round_up(p.n, 32);
            p.pair = false;  // split-K runs on single CTAs
            // workspace: dot [ksplit][nq][dot_ld] fp32 (one slice per k-split), then keys [nq][n] u64
            return (size_t)round_up((int64_t)p.ksplit * p.nq * p.dot_ld * 4, 256) + (size_t)p.nq * p.n * 8 + 256;
        }
    }
    return (size_t)p.nq * p.nsplit * 2 * p.kt * sizeof(uint64_t);  // two epilogue warps per row
}

int l2_launch(const L2SearchParams& p, cudaStream_t stream)
{
    if (p.nq <= 0) return SNV_OK;
    if (p.n <= 0) {
        // empty panel: everything is padding
        SNV_CUDA_CHECK(cudaMemsetAsync(p.partial, 0xFF, (size_t)p.nq * p.kt * 8, stream));
        return merge_keys_launch(p.partial, 1, p.kt, p.nq, p.k, p.id_offset, true, nullptr, p.D_f32, p.I, stream);
    }
    CUtensorMap map_q, map_r;
    int rc = make_map(&map_q, p.q_ops, p.nq, p.kp, BM);
    if (rc) return rc;
    rc = make_map(&map_r, p.ref_ops, p.n, p.kp, p.pair && p.ksplit <= 1 ? BN / 2 : BN);
    if (rc) return rc;
    const int64_t m_tiles = ceil_div(p.nq, BM);
    if (p.ksplit > 1) {
        // split-K: zero the dot matrix, accumulate partial products, then select
        L2SearchParams ps = p;
        ps.dot = reinterpret_cast<float*>(p.partial);
        uint64_t* keys = reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(p.partial) + round_up((int64_t)p.ksplit * p.nq * p.dot_ld * 4, 256));
        // (the opt-in is per device / context: set on every launch, never cached process-wide)
        SNV_CUDA_CHECK(cudaFuncSetAttribute(l2_topk_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<false>()));
        const int64_t n_tiles = ceil_div(p.n, BN);
        const unsigned grid_sk = (unsigned)(m_tiles * n_tiles * p.ksplit);
        profile_begin(stream);
        l2_topk_kernel<8, true><<<grid_sk, kThreads, smem_bytes<false>(), stream>>>(map_q, map_r, ps);
        profile_end(stream);
        SNV_LAUNCH_CHECK();
        const int64_t total = p.nq * p.n;
        const unsigned gk = (unsigned)std::min<int64_t>(ceil_div(total, 256), (int64_t)kNumSMs * 16);
        l2_keys_kernel<<<gk, 256, 0, stream>>>(ps.dot, p.dot_ld, p.ksplit, p.q_norm, p.ref_norm, p.nq, p.n, keys);
        SNV_LAUNCH_CHECK();
        if (p.n > 0x7fffffff) { set_error("L2 split-K: panel too large"); return SNV_ERR_UNSUPPORTED; }
        return merge_keys_launch(keys, 1, (int)p.n, p.nq, p.k, p.id_offset, true, nullptr, p.D_f32, p.I, stream);
    }
    const unsigned grid = p.pair ? (unsigned)(2 * ceil_div(m_tiles, 2) * p.nsplit) : (unsigned)(m_tiles * p.nsplit);
    auto launch = [&](auto kern) -> int {
        const size_t kSmemBytes = p.pair ? smem_bytes<true>() : smem_bytes<false>();
        // the dynamic shared-memory opt-in is per device / context: set on every launch, never cached process-wide
        SNV_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
        profile_begin(stream);
        if (p.pair) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid);
            cfg.blockDim = dim3(kThreads);
            cfg.dynamicSmemBytes = kSmemBytes;
            cfg.stream = stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            SNV_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, map_q, map_r, p));
        } else {
            kern<<<grid, kThreads, kSmemBytes, stream>>>(map_q, map_r, p);
        }
        profile_end(stream);
        return SNV_OK;
    };
    if (p.kt == 8) rc = p.pair ? launch(l2_topk_kernel<8, false, true>) : launch(l2_topk_kernel<8, false, false>);
    else rc = p.pair ? launch(l2_topk_kernel<32, false, true>) : launch(l2_topk_kernel<32, false, false>);
    if (rc) return rc;
    SNV_LAUNCH_CHECK();
    return merge_keys_launch(p.partial, p.nsplit * 2, p.kt, p.nq, p.k, p.id_offset, true, nullptr, p.D_f32, p.I, stream);
}

}  // namespace snv
