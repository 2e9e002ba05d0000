// Bit-packed Hamming / observed-site masked Hamming search on the 5th-gen tensor cores (sm_100a).
//
// Same contract as hamming.cu (exact (distance, id)-lexicographic top-k, bit-exact in the parity
// tests; reference call sites batch_test_faiss_l2.py:110, src/dataset/rag_train_dataset.py:281,
// partial_faiss_intersect.py:82-111) for the shapes where a window's panel is scanned by many
// queries: there the scan is a dense contraction — the reference itself runs it as an sgemm inside
// faiss IndexFlatL2 — and the popcount kernel is bound by the integer pipes, not by HBM.
//
//   d(q, r) = popc((q ^ r) & m) = popc(q & m) + sum_s a_s * r_s,   a_s = m_s * (1 - 2 q_s) in {-1, 0, +1}
//
// The panel stays BIT-PACKED in HBM (132 B per haplotype).  Inside the SM, expander warps turn each
// 256-row x 128-site panel k-block into an fp8 (E4M3) operand tile directly in the SWIZZLE_128B
// shared-memory layout the MMA reads: one LOP3 per 4 sites, because the site -> K position map is
// free to choose (a dot product does not care about the order of K) and a set bit at ANY of bit
// positions 3..6 of a byte is a power of two in E4M3 (2^-6, 2^-5, 2^-3, 2^1).  The query operand
// carries the inverse magnitude (64, 32, 8, 0.5) with the sign / mask folded in, so every product
// is exactly -1, 0 or +1 and the fp32 accumulation in TMEM is exact.
//
// K layout of one k-block (128 sites = packed words w_0..w_3 of the row): byte position
// 32 i + 4 j + b  <->  site 32 i + 8 b + j  (word i, bit 8 b + j); MMA k-step i (32 bytes) is exactly
// packed word i, so a row of `words` packed words costs `words` MMAs (33 for 1030 sites).
//
// CTA (480 threads, one per SM, persistent over (window, query tile, row split) items):
//   warp 0       TMA producer of the query operand tile A [128 x 128 B] (fp8, SWIZZLE_128B)
//   warp 1       TMEM allocator + MMA issuer: tcgen05.mma kind::f8f6f4 M128 N256 K32 into one of two
//                256-column accumulator stages
//   warp 2       TMA producer of the raw packed panel k-blocks [256 rows x 16 B]
//   warps 3-6    expanders: packed bits -> fp8 B tile (32 KB per k-block)
//   warps 7-14   epilogue: thread = query (TMEM lane), two warps per lane quarter on alternate
//                32-column chunks; threshold test on the raw accumulator, candidates appended to
//                per-thread lists in shared memory and folded in lockstep into a register top-k of
//                32-bit keys (distance << idx_bits | row); the two halves are merged in shared memory
//                and the final (D, I) rows are written by the kernel itself.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "common.cuh"
#include "kernels.cuh"
#include "tcgen05.cuh"
#include "topk.cuh"

namespace snv {

namespace {

using namespace tc;

constexpr int BM = 128;          // queries per tile (TMEM lanes)
constexpr int BN = 256;          // panel rows per tile (TMEM columns per accumulator stage)
constexpr int KBLK = 128;        // sites (= fp8 bytes) per k-block: one 128-byte swizzle row
constexpr int kAccStages = 2;
constexpr int kTmemCols = kAccStages * BN;  // 512
constexpr int kExpWarps = 4;
constexpr int kExpThreads = kExpWarps * 32;
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kFirstExpWarp = 3;
constexpr int kFirstEpiWarp = kFirstExpWarp + kExpWarps;  // 7
constexpr int kThreads = 32 * (kFirstEpiWarp + kEpiWarps);  // 480
constexpr uint32_t kABytes = BM * KBLK;   // 16 KB
constexpr uint32_t kBBytes = BN * KBLK;   // 32 KB
constexpr uint32_t kRawBytes = BN * 16;   // 4 KB: 256 rows x 4 packed words
// Three rings: the query operand comes from L2 (long latency: deep ring), the panel operand is made
// in the SM (expander latency: 3 slots), raw packed k-blocks are tiny.
constexpr int kAStages = 5, kBStages = 3, kRawStages = 4;
constexpr int kListCap = 24;              // per-thread candidate list, checked every 16 columns
constexpr uint32_t kListStride = kEpiThreads * 4;  // bytes between consecutive slots of one thread
constexpr size_t kListBytes = (size_t)kListCap * kListStride;
constexpr size_t kSmemBytes = 1024 /*align slack*/ + (size_t)kAStages * kABytes + (size_t)kBStages * kBBytes +
                              (size_t)kRawStages * kRawBytes + kListBytes + 512 /*barriers*/;
static_assert(32 * BM * 4 <= kListBytes, "the half-exchange buffer aliases the candidate lists");
static_assert(kSmemBytes <= 232448, "shared memory budget");

// E4M3 codes.  Panel side: the bit itself, moved (if needed) to one of bit positions 3..6 of its byte;
// query side: the inverse power of two, sign bit = allele 1, zero = unobserved site.
//   j (bit inside the byte):   0     1     2     3     4     5     6     7
//   panel code              0x10  0x20  0x40  0x08  0x10  0x20  0x40  0x08   (2^-5 2^-3 2^1 2^-6 ...)
//   query code              0x60  0x50  0x30  0x68  0x60  0x50  0x30  0x68   (32   8    0.5 64   ...)
__device__ __forceinline__ void expand_panel_word(uint32_t w, uint4& c0, uint4& c1)
{
    const uint32_t lo = w << 4, hi = w >> 4;
    c0 = make_uint4(lo & 0x10101010u, lo & 0x20202020u, lo & 0x40404040u, w & 0x08080808u);
    c1 = make_uint4(w & 0x10101010u, w & 0x20202020u, w & 0x40404040u, hi & 0x08080808u);
}

__device__ __forceinline__ uint32_t query_code(int j)
{
    switch (j & 3) {
        case 0: return 0x60u;
        case 1: return 0x50u;
        case 2: return 0x30u;
        default: return 0x68u;
    }
}

struct TcParams {
    int nw, nq, qtiles;
    int64_t n;               // panel rows per window
    int words, kblocks;      // packed words in use (= MMAs per tile), k-blocks of 4 words
    int n_tiles, nsplit, tiles_per_split;
    int items;               // nw * qtiles * nsplit
    int idx_bits, k;
    int64_t id_offset;
    const int32_t* q_bias;   // [nw * nq] popc(q & m)
    int32_t* D_i32;
    float* D_f32;
    int64_t* I;
    uint64_t* partial;       // [nw * nq][nsplit][kt] when nsplit > 1
};

struct Item {
    int w, qt, split, t0, ntiles;
};
__device__ __forceinline__ Item decode_item(const TcParams& p, int item)
{
    Item it;
    it.split = item % p.nsplit;
    item /= p.nsplit;
    it.qt = item % p.qtiles;
    it.w = item / p.qtiles;
    it.t0 = it.split * p.tiles_per_split;
    const int t1 = it.t0 + p.tiles_per_split < p.n_tiles ? it.t0 + p.tiles_per_split : p.n_tiles;
    it.ntiles = t1 - it.t0;
    return it;
}

template <int N>
struct Ring {
    int i = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void next()
    {
        if (++i == N) { i = 0; phase ^= 1u; }
    }
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

template <int KT, bool EXPAND>
__global__ void __launch_bounds__(kThreads, 1)
hamming_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_r, const TcParams p)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    unsigned char* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);  // SWIZZLE_128B tiles: 1024-byte aligned
    unsigned char* a_tiles = smem;
    unsigned char* b_tiles = a_tiles + (size_t)kAStages * kABytes;
    unsigned char* raws = b_tiles + (size_t)kBStages * kBBytes;
    uint32_t* lists = reinterpret_cast<uint32_t*>(raws + (size_t)kRawStages * kRawBytes);  // [kListCap][256]
    uint32_t* xchg = lists;                                                               // [KT][128], after the lists are folded
    uint64_t* bars = reinterpret_cast<uint64_t*>(lists + kListCap * kEpiThreads);
    uint64_t* full_a = bars;                        // [kAStages]   TMA -> MMA
    uint64_t* empty_a = full_a + kAStages;          // [kAStages]   MMA -> TMA
    uint64_t* full_b = empty_a + kAStages;          // [kBStages]   expanders (or TMA) -> MMA
    uint64_t* empty_b = full_b + kBStages;          // [kBStages]   MMA -> expanders (or TMA)
    uint64_t* raw_full = empty_b + kBStages;        // [kRawStages] TMA -> expanders
    uint64_t* raw_empty = raw_full + kRawStages;    // [kRawStages] expanders -> TMA
    uint64_t* tmem_full = raw_empty + kRawStages;   // [2] MMA -> epilogue
    uint64_t* tmem_empty = tmem_full + kAccStages;  // [2] epilogue -> MMA
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + kAccStages);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_q);
        prefetch_tensormap(&map_r);
        for (int s = 0; s < kAStages; ++s) {
            mbar_init(&full_a[s], 1);
            mbar_init(&empty_a[s], 1);
        }
        for (int s = 0; s < kBStages; ++s) {
            mbar_init(&full_b[s], EXPAND ? kExpWarps : 1);
            mbar_init(&empty_b[s], 1);
        }
        for (int s = 0; s < kRawStages; ++s) {
            mbar_init(&raw_full[s], 1);
            mbar_init(&raw_empty[s], kExpWarps);
        }
        for (int s = 0; s < kAccStages; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], kEpiThreads);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int KB = p.kblocks;

    if (warp == 0) {
        // ================= TMA producer: query operand tiles =================
        if (lane == 0) {
            Ring<kAStages> ra;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
                const Item it = decode_item(p, item);
                const int row_a = it.w * p.nq + it.qt * BM;
                for (int t = 0; t < it.ntiles; ++t) {
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait(&empty_a[ra.i], ra.phase ^ 1u);
                        mbar_arrive_expect_tx(&full_a[ra.i], kABytes);
                        tma_load_2d(a_tiles + (size_t)ra.i * kABytes, &map_q, kb * KBLK, row_a, &full_a[ra.i]);
                        ra.next();
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (one thread) =================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_e4m3(BM, BN);
            Ring<kAStages> ra;
            Ring<kBStages> rb;
            uint32_t tcount = 0;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
                const Item it = decode_item(p, item);
                for (int t = 0; t < it.ntiles; ++t, ++tcount) {
                    const uint32_t as = tcount & 1u;
                    mbar_wait(&tmem_empty[as], ((tcount >> 1) & 1u) ^ 1u);
                    tcgen05_fence_after();
                    const uint32_t d_tmem = tmem_base + as * BN;
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait(&full_a[ra.i], ra.phase);
                        mbar_wait(&full_b[rb.i], rb.phase);
                        tcgen05_fence_after();
                        const uint32_t a_addr = smem_u32(a_tiles + (size_t)ra.i * kABytes);
                        const uint32_t b_addr = smem_u32(b_tiles + (size_t)rb.i * kBBytes);
                        const int nm = (kb == KB - 1) ? p.words - 4 * kb : 4;  // one MMA per packed word
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (k < nm) {
                                const uint64_t adesc = make_kmajor_sw128_desc(a_addr + k * 32);
                                const uint64_t bdesc = make_kmajor_sw128_desc(b_addr + k * 32);
                                umma_f8(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
                            }
                        }
                        umma_commit(&empty_a[ra.i]);
                        umma_commit(&empty_b[rb.i]);
                        if (kb == KB - 1) umma_commit(&tmem_full[as]);
                        ra.next();
                        rb.next();
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 2) {
        // ================= TMA producer: raw packed panel k-blocks (fp8 panel tiles when !EXPAND) =================
        if (lane == 0) {
            Ring<kRawStages> rr;
            Ring<kBStages> rb;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
                const Item it = decode_item(p, item);
                for (int t = 0; t < it.ntiles; ++t) {
                    const int n0 = (it.t0 + t) * BN;
                    for (int kb = 0; kb < KB; ++kb) {
                        if constexpr (EXPAND) {
                            mbar_wait(&raw_empty[rr.i], rr.phase ^ 1u);
                            mbar_arrive_expect_tx(&raw_full[rr.i], kRawBytes);
                            tma_load_3d(raws + (size_t)rr.i * kRawBytes, &map_r, kb * 4, n0, it.w, &raw_full[rr.i]);
                            rr.next();
                        } else {
                            mbar_wait(&empty_b[rb.i], rb.phase ^ 1u);
                            mbar_arrive_expect_tx(&full_b[rb.i], kBBytes);
                            tma_load_3d(b_tiles + (size_t)rb.i * kBBytes, &map_r, kb * KBLK, n0, it.w, &full_b[rb.i]);
                            rb.next();
                        }
                    }
                }
            }
        }
    } else if (warp < kFirstEpiWarp) {
        // ================= expanders: packed bits -> fp8 operand tile (SWIZZLE_128B, K-major) =================
        if constexpr (EXPAND) {
            const int et = (warp - kFirstExpWarp) * 32 + lane;
            Ring<kRawStages> rr;
            Ring<kBStages> rb;
            int pending = -1;  // B slot whose stores still need the proxy fence + arrive (deferred by one k-block)
            const int sw = et & 7;  // rows et and et + 128 share the swizzle phase
            for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
                const Item it = decode_item(p, item);
                const int nkb = it.ntiles * KB;
                for (int g = 0; g < nkb; ++g) {
                    mbar_wait(&raw_full[rr.i], rr.phase);
                    const unsigned char* src = raws + (size_t)rr.i * kRawBytes;
                    const uint4 w0 = *reinterpret_cast<const uint4*>(src + et * 16);
                    const uint4 w1 = *reinterpret_cast<const uint4*>(src + (et + kExpThreads) * 16);
                    if (pending >= 0) {
                        // the previous k-block's stores have had time to drain: make them visible to the
                        // tensor core (async proxy) and publish the slot
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&full_b[pending]);
                    }
                    mbar_wait(&empty_b[rb.i], rb.phase ^ 1u);
                    unsigned char* d0 = b_tiles + (size_t)rb.i * kBBytes + et * 128;
                    unsigned char* d1 = d0 + kExpThreads * 128;
                    uint4 c0, c1;
                    expand_panel_word(w0.x, c0, c1);
                    *reinterpret_cast<uint4*>(d0 + ((0 ^ sw) << 4)) = c0;
                    *reinterpret_cast<uint4*>(d0 + ((1 ^ sw) << 4)) = c1;
                    expand_panel_word(w0.y, c0, c1);
                    *reinterpret_cast<uint4*>(d0 + ((2 ^ sw) << 4)) = c0;
                    *reinterpret_cast<uint4*>(d0 + ((3 ^ sw) << 4)) = c1;
                    expand_panel_word(w0.z, c0, c1);
                    *reinterpret_cast<uint4*>(d0 + ((4 ^ sw) << 4)) = c0;
                    *reinterpret_cast<uint4*>(d0 + ((5 ^ sw) << 4)) = c1;
                    expand_panel_word(w0.w, c0, c1);
                    *reinterpret_cast<uint4*>(d0 + ((6 ^ sw) << 4)) = c0;
                    *reinterpret_cast<uint4*>(d0 + ((7 ^ sw) << 4)) = c1;
                    expand_panel_word(w1.x, c0, c1);
                    *reinterpret_cast<uint4*>(d1 + ((0 ^ sw) << 4)) = c0;
                    *reinterpret_cast<uint4*>(d1 + ((1 ^ sw) << 4)) = c1;
                    expand_panel_word(w1.y, c0, c1);
                    *reinterpret_cast<uint4*>(d1 + ((2 ^ sw) << 4)) = c0;
                    *reinterpret_cast<uint4*>(d1 + ((3 ^ sw) << 4)) = c1;
                    expand_panel_word(w1.z, c0, c1);
                    *reinterpret_cast<uint4*>(d1 + ((4 ^ sw) << 4)) = c0;
                    *reinterpret_cast<uint4*>(d1 + ((5 ^ sw) << 4)) = c1;
                    expand_panel_word(w1.w, c0, c1);
                    *reinterpret_cast<uint4*>(d1 + ((6 ^ sw) << 4)) = c0;
                    *reinterpret_cast<uint4*>(d1 + ((7 ^ sw) << 4)) = c1;
                    __syncwarp();  // every lane has consumed its raw words
                    if (lane == 0) mbar_arrive(&raw_empty[rr.i]);
                    pending = rb.i;
                    rr.next();
                    rb.next();
                }
            }
            if (pending >= 0) {
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_b[pending]);
            }
        }
    } else {
        // ================= epilogue: thread = query row, 2 warps per lane quarter =================
        const int quarter = warp & 3;                   // TMEM lanes [32 * quarter, +32)
        const int half = (warp - kFirstEpiWarp) >> 2;   // columns [128 * half, +128) of every tile
        const int row = quarter * 32 + lane;
        const int et = (warp - kFirstEpiWarp) * 32 + lane;
        const uint32_t list_base = smem_u32(lists + et);  // slot s at list_base + s * kListStride
        const int idx_bits = p.idx_bits;
        uint32_t tcount = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            const Item it = decode_item(p, item);
            const int qi = it.qt * BM + row;
            const bool active = qi < p.nq;
            const int64_t q = (int64_t)it.w * p.nq + (active ? qi : 0);
            const int32_t qb = active ? p.q_bias[q] : 0;
            uint32_t best[KT];
#pragma unroll
            for (int i = 0; i < KT; ++i) best[i] = kSent32;
            // A candidate is scored as ONE float, kf = 128 * acc + jj (jj = column inside this half of the
            // tile, 0..127; exact: |acc| < 2^12): one FFMA + one compare per column, and `kf < thr` with
            // thr = 128 * (acc of the k-th best) also rejects equal distances at later columns (ids ascend
            // along the scan).  Survivors are appended to a per-thread list in shared memory and folded
            // into the sorted register top-k in lockstep.
            float thr = 3.0e38f;
            uint32_t lp = list_base;
            auto fold = [&](uint32_t col_base) {
                const int cnt = (int)((lp - list_base) / kListStride);
                const int maxc = __reduce_max_sync(0xffffffffu, cnt);
                for (int s2 = 0; s2 < maxc; ++s2) {
                    if (s2 < cnt) {
                        const int ki = __float2int_rn(__uint_as_float(lists[et + s2 * kEpiThreads]));
                        const uint32_t key = ((uint32_t)((ki >> 7) + qb) << idx_bits) | (col_base + (uint32_t)(ki & 127));
                        if (key < best[KT - 1]) topk_insert<KT, uint32_t>(best, key);
                    }
                }
                lp = list_base;
                thr = best[KT - 1] == kSent32 ? 3.0e38f : (float)(((int32_t)(best[KT - 1] >> idx_bits) - qb) * 128);
            };
            for (int t = 0; t < it.ntiles; ++t, ++tcount) {
                const uint32_t as = tcount & 1u;
                const int n0 = (it.t0 + t) * BN;
                mbar_wait(&tmem_full[as], (tcount >> 1) & 1u);
                tcgen05_fence_after();
                int cols = (p.n - n0 < BN ? (int)(p.n - n0) : BN) - 128 * half;  // columns of this half in use
                cols = cols < 0 ? 0 : (cols > 128 ? 128 : cols);
                const int nch = (cols + 31) >> 5;
                const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * BN + (uint32_t)(128 * half);
                const uint32_t col_base = (uint32_t)(t * BN + 128 * half);  // columns count from the split's first row
                uint32_t accA[32], accB[32];
                auto process = [&](uint32_t (&acc)[32], auto u_tag) {
                    constexpr int u = decltype(u_tag)::value;
                    if (32 * u + 32 > cols) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (32 * u + j >= cols) acc[j] = 0x7F800000u;  // past the panel end (+inf): never a candidate
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float kf = fmaf(__uint_as_float(acc[j]), 128.0f, (float)(32 * u + j));
                        if (kf < thr) {
                            asm volatile("st.shared.f32 [%0], %1;" ::"r"(lp), "f"(kf) : "memory");
                            lp += kListStride;
                        }
                        if (j == 15 || j == 31) {
                            if (__any_sync(0xffffffffu, lp > list_base + (kListCap - 16) * kListStride)) fold(col_base);
                        }
                    }
                };
                if (nch > 0) tmem_ld_32x32b_x32(tbase, accA);
                if (nch > 0) {
                    tmem_ld_wait(accA);
                    if (nch > 1) tmem_ld_32x32b_x32(tbase + 32u, accB);
                    process(accA, std::integral_constant<int, 0>{});
                }
                if (nch > 1) {
                    tmem_ld_wait(accB);
                    if (nch > 2) tmem_ld_32x32b_x32(tbase + 64u, accA);
                    process(accB, std::integral_constant<int, 1>{});
                }
                if (nch > 2) {
                    tmem_ld_wait(accA);
                    if (nch > 3) tmem_ld_32x32b_x32(tbase + 96u, accB);
                    process(accA, std::integral_constant<int, 2>{});
                }
                if (nch > 3) {
                    tmem_ld_wait(accB);
                    process(accB, std::integral_constant<int, 3>{});
                }
                tcgen05_fence_before();
                mbar_arrive(&tmem_empty[as]);
                fold(col_base);  // list entries carry tile-relative columns: fold before the next tile
            }
            // ---- merge the two halves of each query through shared memory, then write the result
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");  // every list is folded: the buffer is free
            if (half == 1) {
#pragma unroll
                for (int i = 0; i < KT; ++i) xchg[i * BM + row] = best[i];
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
            if (half == 0) {
#pragma unroll 1
                for (int i = 0; i < KT; ++i) {
                    const uint32_t key = xchg[i * BM + row];
                    if (key < best[KT - 1]) topk_insert<KT, uint32_t>(best, key);
                }
                if (active) {
                    const uint32_t idx_mask = (1u << idx_bits) - 1u;
                    const int64_t r0 = (int64_t)it.t0 * BN;
                    if (p.nsplit == 1) {
#pragma unroll
                        for (int i = 0; i < KT; ++i) {
                            if (i < p.k) {
                                const uint32_t key = best[i];
                                const bool none = key == kSent32;
                                const int32_t dist = none ? 0x7FFFFFFF : (int32_t)(key >> idx_bits);
                                const int64_t o = q * p.k + i;
                                if (p.D_i32) p.D_i32[o] = dist;
                                if (p.D_f32) p.D_f32[o] = none ? 3.4028234663852886e38f : (float)dist;
                                p.I[o] = none ? -1 : (int64_t)(key & idx_mask) + r0 + p.id_offset;
                            }
                        }
                    } else {
                        uint64_t* out = p.partial + (q * p.nsplit + it.split) * KT;
#pragma unroll
                        for (int i = 0; i < KT; ++i) {
                            const uint32_t key = best[i];
                            out[i] = key == kSent32 ? kSent64
                                                    : ((uint64_t)(key >> idx_bits) << 32) | (uint64_t)((int64_t)(key & idx_mask) + r0);
                        }
                    }
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");  // exchange reads done before lists are reused
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---- query operand: packed (q, observed mask) -> fp8 rows [rows][kblocks * 128] + popc(q & m) ----------------
__global__ void __launch_bounds__(256)
tc_expand_queries_kernel(const uint32_t* __restrict__ q, const uint32_t* __restrict__ mask, int64_t mask_win_stride,
                         int64_t mask_q_stride, int nq, int64_t rows, int stride, int words, int d, int kblocks,
                         uint8_t* __restrict__ ops, int32_t* __restrict__ bias)
{
    const int cpr = kblocks * 8;  // 16-byte chunks per row
    const int64_t total = rows * cpr;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = idx / cpr;
        const int ch = (int)(idx % cpr);
        const int wi = (ch >> 3) * 4 + ((ch & 7) >> 1);
        const int h = ch & 1;
        const uint32_t* qr = q + row * stride;
        const uint32_t* mr = mask ? mask + (row / nq) * mask_win_stride + (row % nq) * mask_q_stride : nullptr;
        auto valid_bits = [&](int w) -> uint32_t {
            const int lo = w * 32;
            return lo + 32 <= d ? 0xFFFFFFFFu : (lo < d ? (1u << (d - lo)) - 1u : 0u);
        };
        uint32_t wm = 0, wq = 0;
        if (wi < words) {
            wm = valid_bits(wi);
            if (mr) wm &= mr[wi];
            wq = qr[wi] & wm;
        }
        uint32_t out[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = 4 * h + jj;
            const uint32_t mb = (wm >> j) & 0x01010101u;
            const uint32_t sb = (wq >> j) & 0x01010101u;
            out[jj] = mb * query_code(j) | sb * 0x80u;
        }
        *reinterpret_cast<uint4*>(ops + (row * cpr + ch) * 16) = make_uint4(out[0], out[1], out[2], out[3]);
        if (ch == 0) {
            int32_t b = 0;
            for (int w = 0; w < words; ++w) {
                uint32_t m = valid_bits(w);
                if (mr) m &= mr[w];
                b += __popc(qr[w] & m);
            }
            bias[row] = b;
        }
    }
}

// ---- bring-up variant (EXPAND = false): the panel expanded to fp8 rows in HBM ------------------------------
__global__ void __launch_bounds__(256)
tc_expand_panel_kernel(const uint32_t* __restrict__ panel, int64_t panel_win_stride, int nw, int64_t n, int stride,
                       int words, int kblocks, uint8_t* __restrict__ ops)
{
    const int cpr = kblocks * 8;
    const int64_t total = (int64_t)nw * n * cpr;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = idx / cpr;  // w * n + r
        const int ch = (int)(idx % cpr);
        const int wi = (ch >> 3) * 4 + ((ch & 7) >> 1);
        const uint32_t w = wi < words ? panel[(row / n) * panel_win_stride + (row % n) * stride + wi] : 0u;
        uint4 c0, c1;
        expand_panel_word(w, c0, c1);
        *reinterpret_cast<uint4*>(ops + (row * cpr + ch) * 16) = (ch & 1) ? c1 : c0;
    }
}

int encode_map(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const cuuint64_t* gdim,
               const cuuint64_t* gstride, const cuuint32_t* box, CUtensorMapSwizzle swz)
{
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return SNV_ERR_CUDA;
    }
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
        return SNV_ERR_CUDA;
    }
    return SNV_OK;
}

int bit_length64(int64_t v)
{
    int b = 0;
    while (v > 0) { ++b; v >>= 1; }
    return b;
}

template <int KT, bool EXPAND>
int launch_kernel(const CUtensorMap& map_q, const CUtensorMap& map_r, const TcParams& tp, int grid, cudaStream_t stream)
{
    constexpr size_t smem = kSmemBytes;
    static bool attr = false;
    if (!attr) {
        SNV_CUDA_CHECK(cudaFuncSetAttribute(hamming_tc_kernel<KT, EXPAND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    profile_begin(stream);
    hamming_tc_kernel<KT, EXPAND><<<grid, kThreads, smem, stream>>>(map_q, map_r, tp);
    profile_end(stream);
    SNV_LAUNCH_CHECK();
    return SNV_OK;
}

}  // namespace

// 0 = popcount kernel, 1 = tensor cores with in-SM expansion, 2 = tensor cores, panel pre-expanded (bring-up)
int hamming_engine_for(const HammingSearchParams& p)
{
    int mode = -1;  // auto
    if (const char* e = getenv("SNV_HAMMING_ENGINE")) {
        if (!strcmp(e, "popc")) mode = 0;
        else if (!strcmp(e, "tc")) mode = 1;
        else if (!strcmp(e, "tc_hbm")) mode = 2;
    }
    const bool can = !p.work && p.n > 0 && p.nq > 0 && p.k >= 1 && p.k <= 32 && p.d < (1 << 12) &&
                     (int64_t)p.nw * p.nq < ((int64_t)1 << 31) && p.n < ((int64_t)1 << 31);
    if (!can || mode == 0) return 0;
    if (mode > 0) return mode;
    // auto: enough queries per window to fill a useful part of the 128-lane tile, and a panel worth a tile
    return (p.nq >= 32 && p.n >= 2 * BN) ? 1 : 0;
}

size_t hamming_tc_plan(const HammingSearchParams& p, HammingTcPlan& plan)
{
    plan = HammingTcPlan{};
    plan.engine = hamming_engine_for(p);
    if (!plan.engine) return 0;
    plan.kt = p.k <= 8 ? 8 : 32;
    plan.kblocks = (int)ceil_div(p.words, 4);
    plan.qtiles = (int)ceil_div(p.nq, BM);
    plan.n_tiles = (int)ceil_div(p.n, BN);
    plan.idx_bits = 32 - bit_length64((int64_t)p.d + 1);
    const int64_t max_tiles_per_split = ((int64_t)1 << plan.idx_bits) / BN;
    // row splits only when (window, query tile) items cannot fill the machine, or for the key's id field
    const int64_t base = (int64_t)p.nw * plan.qtiles;
    int64_t nsplit = 1;
    if (base < kNumSMs) nsplit = std::min<int64_t>(ceil_div(kNumSMs, base), std::max<int64_t>(1, plan.n_tiles / 2));
    nsplit = std::max<int64_t>(nsplit, ceil_div(plan.n_tiles, max_tiles_per_split));
    plan.tiles_per_split = (int)ceil_div(plan.n_tiles, nsplit);
    plan.nsplit = (int)ceil_div(plan.n_tiles, plan.tiles_per_split);
    if (base * plan.nsplit > 0x7fffffffLL) {
        set_error("hamming search: grid too large");
        return (size_t)-1;
    }
    const int64_t rows = (int64_t)p.nw * p.nq;
    plan.off_bias = round_up(rows * plan.kblocks * KBLK, 256);
    plan.off_partial = plan.off_bias + round_up(rows * 4, 256);
    plan.off_panel = plan.off_partial + (plan.nsplit > 1 ? round_up(rows * plan.nsplit * plan.kt * 8, 256) : 0);
    size_t total = plan.off_panel;
    if (plan.engine == 2) total += (size_t)p.nw * p.n * plan.kblocks * KBLK;
    return total + 1024;  // + slack so that the last query tile's TMA box stays inside the allocation
}

int hamming_tc_launch(const HammingSearchParams& p, const HammingTcPlan& plan, void* ws, cudaStream_t stream)
{
    if (p.nw <= 0 || p.nq <= 0) return SNV_OK;
    const int64_t rows = (int64_t)p.nw * p.nq;
    uint8_t* q_ops = static_cast<uint8_t*>(ws);
    int32_t* q_bias = reinterpret_cast<int32_t*>(q_ops + plan.off_bias);
    uint64_t* partial = reinterpret_cast<uint64_t*>(q_ops + plan.off_partial);
    uint8_t* panel_ops = q_ops + plan.off_panel;
    const int kbytes = plan.kblocks * KBLK;
    {
        const int64_t total = rows * plan.kblocks * 8;
        const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(total, 256), (int64_t)kNumSMs * 32);
        tc_expand_queries_kernel<<<grid, 256, 0, stream>>>(p.q, p.mask, p.mask_win_stride, p.mask_q_stride, p.nq, rows, p.stride,
                                                          p.words, p.d, plan.kblocks, q_ops, q_bias);
        SNV_LAUNCH_CHECK();
    }
    CUtensorMap map_q, map_r;
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)kbytes, (cuuint64_t)rows};
        const cuuint64_t gstride[1] = {(cuuint64_t)kbytes};
        const cuuint32_t box[2] = {KBLK, BM};
        int rc = encode_map(&map_q, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, q_ops, gdim, gstride, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    if (plan.engine == 1) {
        // raw packed panel [nw][n][stride] words: box = 4 words x 256 rows of one window
        const cuuint64_t gdim[3] = {(cuuint64_t)p.stride, (cuuint64_t)p.n, (cuuint64_t)p.nw};
        const cuuint64_t gstride[2] = {(cuuint64_t)p.stride * 4, (cuuint64_t)p.panel_win_stride * 4};
        const cuuint32_t box[3] = {4, BN, 1};
        int rc = encode_map(&map_r, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, p.panel, gdim, gstride, box, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc) return rc;
    } else {
        const int64_t total = (int64_t)p.nw * p.n * plan.kblocks * 8;
        const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(total, 256), (int64_t)kNumSMs * 32);
        tc_expand_panel_kernel<<<grid, 256, 0, stream>>>(p.panel, p.panel_win_stride, p.nw, p.n, p.stride, p.words, plan.kblocks, panel_ops);
        SNV_LAUNCH_CHECK();
        const cuuint64_t gdim[3] = {(cuuint64_t)kbytes, (cuuint64_t)p.n, (cuuint64_t)p.nw};
        const cuuint64_t gstride[2] = {(cuuint64_t)kbytes, (cuuint64_t)p.n * kbytes};
        const cuuint32_t box[3] = {KBLK, BN, 1};
        int rc = encode_map(&map_r, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, panel_ops, gdim, gstride, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    TcParams tp{};
    tp.nw = p.nw; tp.nq = p.nq; tp.qtiles = plan.qtiles;
    tp.n = p.n;
    tp.words = p.words; tp.kblocks = plan.kblocks;
    tp.n_tiles = plan.n_tiles; tp.nsplit = plan.nsplit; tp.tiles_per_split = plan.tiles_per_split;
    tp.items = p.nw * plan.qtiles * plan.nsplit;
    tp.idx_bits = plan.idx_bits; tp.k = p.k;
    tp.id_offset = p.id_offset;
    tp.q_bias = q_bias;
    tp.D_i32 = p.D_i32; tp.D_f32 = p.D_f32; tp.I = p.I;
    tp.partial = partial;
    const int grid = std::min(tp.items, kNumSMs);
    int rc;
    if (plan.engine == 1) rc = plan.kt == 8 ? launch_kernel<8, true>(map_q, map_r, tp, grid, stream) : launch_kernel<32, true>(map_q, map_r, tp, grid, stream);
    else rc = plan.kt == 8 ? launch_kernel<8, false>(map_q, map_r, tp, grid, stream) : launch_kernel<32, false>(map_q, map_r, tp, grid, stream);
    if (rc) return rc;
    if (plan.nsplit > 1)
        return merge_keys_launch(partial, plan.nsplit, plan.kt, rows, p.k, p.id_offset, false, p.D_i32, p.D_f32, p.I, stream);
    return SNV_OK;
}

}  // namespace snv
