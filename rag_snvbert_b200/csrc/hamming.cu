// Bit-packed Hamming / observed-site masked Hamming scan with a fused exact top-k (sm_100a).
//
// Replaces faiss IndexFlatL2.search on 0/1 rows and IndexBinaryFlat.search
// (reference call sites: batch_test_faiss_l2.py:110, src/dataset/rag_train_dataset.py:281,
// test_faiss_intersect.py:160,181, partial_faiss_intersect.py:82-111).
//
// Mapping: one thread = one query haplotype.  The query's packed words (and its observed-site
// mask) live in registers for the whole scan; the window's panel streams through shared
// memory in row tiles fetched by 1-D bulk async copies (TMA engine, mbarrier completion,
// kStages deep) issued by one thread, and every warp reads the current panel row as a
// shared-memory broadcast.  Distances use a carry-save-adder tree (LOP3) in front of POPC so
// that the quarter-rate POPC pipe and the ALU pipe are balanced.
// Candidates are ranked on ONE 32-bit key  (distance << idx_bits | row_in_split)  so the
// result is the exact (distance, id)-lexicographic top-k, ties included.
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"
#include "topk.cuh"

#ifndef SNV_CSA_DEPTH
#define SNV_CSA_DEPTH 2
#endif
#ifndef SNV_ROW_UNROLL
#define SNV_ROW_UNROLL 2
#endif
#ifndef SNV_POPC_MODE
#define SNV_POPC_MODE 1  // 0: plain adds (ALU pipe), 1: one IMAD chain, 2: four IMAD chains (all within 3% on B200, profiles/r1_tuning_hamming.txt)
#endif
#ifndef SNV_BLOCK
#define SNV_BLOCK 128
#endif
#ifndef SNV_ROW_GROUP
#define SNV_ROW_GROUP 1
#endif
#define SNV_PRAGMA_(x) _Pragma(#x)
#define SNV_UNROLL(n) SNV_PRAGMA_(unroll n)

namespace snv {

namespace {

constexpr int kMaxBlock = SNV_BLOCK;
constexpr int kListKT = 32;    // top-k sizes from here on select through shared-memory candidate lists
constexpr int kListCap = 24;
constexpr int kBarBytes = 128;  // smem reserved for the stage mbarriers

// acc + w * popc(x): the multiply-add goes to the FMA pipe (IMAD) because `w` is a runtime
// register (1, 2 or 4 from the kernel parameters), keeping the saturated ALU pipe for LOP3 only.
__device__ __forceinline__ uint32_t popc_mad(uint32_t x, uint32_t w, uint32_t acc)
{
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"((uint32_t)__popc(x)), "r"(w), "r"(acc));
    return r;
}

// Accumulates sum_i wt[LEVEL] * popcount(x[i]) over the multiset x[0..N) into acc[]: DEPTH levels
// of 3:2 carry-save compressors (carries move to weight level LEVEL+1), then POPC.  The adds are
// spread round-robin over NACC independent accumulators (short dependency chains).
template <int N, int DEPTH, int LEVEL, int NACC>
struct WeightedPopc {
    template <int SLOT>
    static __device__ __forceinline__ void run(const uint32_t (&x)[N], const uint32_t (&wt)[3], uint32_t (&acc)[NACC])
    {
        if constexpr (DEPTH == 0 || N < 3) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
#if SNV_POPC_MODE == 0
                acc[(SLOT + i) % NACC] += (uint32_t)__popc(x[i]) << LEVEL;
#else
                acc[(SLOT + i) % NACC] = popc_mad(x[i], wt[LEVEL], acc[(SLOT + i) % NACC]);
#endif
            }
        } else {
            constexpr int T = N / 3, R = N % 3;
            uint32_t sum[T + R], carry[T];
#pragma unroll
            for (int i = 0; i < T; ++i) {
                const uint32_t a = x[3 * i], b = x[3 * i + 1], c = x[3 * i + 2];
                sum[i] = a ^ b ^ c;                    // LOP3 0x96
                carry[i] = (a & b) | (c & (a ^ b));    // LOP3 0xE8 (majority)
            }
#pragma unroll
            for (int r = 0; r < R; ++r) sum[T + r] = x[3 * T + r];
            WeightedPopc<T + R, DEPTH - 1, LEVEL, NACC>::template run<SLOT>(sum, wt, acc);
            WeightedPopc<T, DEPTH - 1, LEVEL + 1, NACC>::template run<(SLOT + T + R) % NACC>(carry, wt, acc);
        }
    }
};

// Which (window, query, row split) a thread works on.  Uniform mode: blockIdx = ((w * qtiles) + qt)
// * nsplit + split, queries laid out [nw][nq].  Grouped mode (p.work != nullptr): one work item
// (window, start, count) per CTA group; queries stay in caller order and are reached through the
// window-sorted permutation p.order (ragged per-window batches of a training step).
struct Slot {
    int w, split;
    bool active;
    int64_t qrow;    // row of this query in q / mask / outputs
    int64_t mrow;    // word offset of its mask row
};
__device__ __forceinline__ Slot decode_slot(const HammingSearchParams& p)
{
    Slot s;
    int b = blockIdx.x;
    s.split = b % p.nsplit;
    b /= p.nsplit;
    if (p.work) {
        const int w = p.work[3 * b], start = p.work[3 * b + 1], count = p.work[3 * b + 2];
        s.w = w;
        s.active = (int)threadIdx.x < count;
        s.qrow = s.active ? p.order[start + threadIdx.x] : 0;
        s.mrow = p.mask_q_stride ? s.qrow * p.mask_q_stride : (int64_t)w * p.mask_win_stride;
    } else {
        const int qt = b % p.qtiles;
        s.w = b / p.qtiles;
        const int qi = qt * blockDim.x + threadIdx.x;
        s.active = qi < p.nq;
        s.qrow = (int64_t)s.w * p.nq + (s.active ? qi : 0);
        s.mrow = (int64_t)s.w * p.mask_win_stride + (int64_t)(s.active ? qi : 0) * p.mask_q_stride;
    }
    return s;
}

template <int NW>
__device__ __forceinline__ void load_row_regs(uint32_t (&dst)[NW], const uint32_t* __restrict__ src)
{
    constexpr int V = NW / 4;
#pragma unroll
    for (int c = 0; c < V; ++c) {
        const uint4 v = reinterpret_cast<const uint4*>(src)[c];
        dst[4 * c] = v.x; dst[4 * c + 1] = v.y; dst[4 * c + 2] = v.z; dst[4 * c + 3] = v.w;
    }
#pragma unroll
    for (int i = 4 * V; i < NW; ++i) dst[i] = src[i];
}

template <int NW, bool MASKED, int KT>
__global__ void __launch_bounds__(kMaxBlock)
hamming_topk_kernel(const HammingSearchParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
    uint32_t* tiles = reinterpret_cast<uint32_t*>(smem_raw + kBarBytes);

    const int tid = threadIdx.x;
    const Slot slot = decode_slot(p);
    const int split = slot.split;
    const int w = slot.w;

    const int64_t r0 = (int64_t)split * p.rows_per_split;
    const int64_t r1 = (r0 + p.rows_per_split < p.n) ? r0 + p.rows_per_split : p.n;
    const int nrows = (int)(r1 - r0);
    const int TR = p.tile_rows;
    const int ntiles = (nrows + TR - 1) / TR;
    const int stages = p.stages;
    const uint32_t tile_words = (uint32_t)TR * p.stride;
    const uint32_t* pw = p.panel + (int64_t)w * p.panel_win_stride + r0 * p.stride;

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
        fence_barrier_init();
    }
    __syncthreads();

    auto issue = [&](int t) {
        const int rows = (nrows - t * TR < TR) ? nrows - t * TR : TR;
        const uint32_t bytes = (uint32_t)rows * p.stride * 4u;
        uint64_t* bar = &bars[t % stages];
        mbar_arrive_expect_tx(bar, bytes);
        bulk_g2s(tiles + (size_t)(t % stages) * tile_words, pw + (size_t)t * tile_words, bytes, bar);
    };
    if (tid == 0) {
        for (int t = 0; t < stages - 1 && t < ntiles; ++t) issue(t);
    }

    // ---- this thread's query (and observed-site mask) -> registers
    const bool active = slot.active;
    uint32_t q[NW];
    uint32_t m[MASKED ? NW : 1];
    if (active) {
        load_row_regs<NW>(q, p.q + slot.qrow * p.stride);
        if constexpr (MASKED) {
            load_row_regs<NW>(m, p.mask + slot.mrow);
#pragma unroll
            for (int i = 0; i < NW; ++i) q[i] &= m[i];
        }
    } else {
#pragma unroll
        for (int i = 0; i < NW; ++i) q[i] = 0;
        if constexpr (MASKED) {
#pragma unroll
            for (int i = 0; i < NW; ++i) m[i] = 0;
        }
    }

    uint32_t best[KT];
#pragma unroll
    for (int i = 0; i < KT; ++i) best[i] = kSent32;

    // candidate lists (KT >= kListKT only): [kListCap][kMaxBlock] keys after the panel tiles
    uint32_t* my_list = tiles + (size_t)stages * tile_words + tid;
    uint32_t thr = kSent32;
    int cnt = 0;
    auto fold = [&]() {
        const int maxc = __reduce_max_sync(0xffffffffu, cnt);
        for (int s2 = 0; s2 < maxc; ++s2) {
            if (s2 < cnt) {
                const uint32_t key = my_list[s2 * kMaxBlock];
                if (key < best[KT - 1]) topk_insert<KT, uint32_t>(best, key);
            }
        }
        cnt = 0;
        thr = best[KT - 1];
    };
    const int idx_bits = p.idx_bits;
    // popcount weights pre-scaled by 2^idx_bits (runtime values: see popc_mad)
    const uint32_t wt[3] = {p.wt1 << idx_bits, p.wt2 << idx_bits, p.wt4 << idx_bits};
    for (int t = 0; t < ntiles; ++t) {
        if (tid == 0 && t + stages - 1 < ntiles) issue(t + stages - 1);
        mbar_wait(&bars[t % stages], (uint32_t)(t / stages) & 1u);
        const uint32_t* tile = tiles + (size_t)(t % stages) * tile_words;
        const int rows = (nrows - t * TR < TR) ? nrows - t * TR : TR;
        const uint32_t row_base = (uint32_t)(t * TR);
        auto row_key = [&](int j) -> uint32_t {
            uint32_t x[NW];
            load_row_regs<NW>(x, tile + (size_t)j * p.stride);  // warp-uniform address: broadcast
#pragma unroll
            for (int i = 0; i < NW; ++i) {
                if constexpr (MASKED) x[i] = (x[i] & m[i]) ^ q[i];  // == (r ^ q) & m, one LOP3
                else                  x[i] ^= q[i];
            }
            // key = dist * 2^idx_bits + row (weights are pre-scaled; accumulators sum to dist << idx_bits)
#if SNV_POPC_MODE == 2
            constexpr int NACC = NW >= 16 ? 4 : 1;
#else
            constexpr int NACC = 1;
#endif
            uint32_t acc[NACC];
#pragma unroll
            for (int a = 0; a < NACC; ++a) acc[a] = 0;
            WeightedPopc<NW, SNV_CSA_DEPTH, 0, NACC>::template run<0>(x, wt, acc);
#if SNV_POPC_MODE == 0
            return (acc[0] << idx_bits) | (row_base + (uint32_t)j);
#else
            uint32_t key = row_base + (uint32_t)j;
#pragma unroll
            for (int a = 0; a < NACC; ++a) key += acc[a];
            return key;
#endif
        };
#if SNV_ROW_GROUP > 1
        // SNV_ROW_GROUP rows are scored back to back before any (rare, divergent) insertion, so the
        // LOP3 work of one row overlaps the POPC burst of the previous one inside a warp.
        for (int j = 0; j < rows; j += SNV_ROW_GROUP) {
            uint32_t keys[SNV_ROW_GROUP];
#pragma unroll
            for (int g = 0; g < SNV_ROW_GROUP; ++g) {
                keys[g] = row_key(j + g);                       // rows past the tile end read stale smem...
                if (j + g >= rows) keys[g] = kSent32;            // ...and are discarded here
            }
#pragma unroll
            for (int g = 0; g < SNV_ROW_GROUP; ++g)
                if (keys[g] < best[KT - 1]) topk_insert<KT, uint32_t>(best, keys[g]);
        }
#else
        if constexpr (KT >= kListKT) {
            // Large k: a 2*KT-instruction insertion whenever ANY lane of the warp finds a candidate
            // would cost ~30 % of the scan.  Candidates go to a per-thread list in shared memory
            // instead and are folded into the sorted registers in lockstep, when some lane has more
            // than kListCap - 8 pending (checked every 8 rows).  thr lags best[KT-1] between folds,
            // which only admits extra candidates; the fold compares exact keys.
            for (int j0 = 0; j0 < rows; j0 += 8) {
                const int j1 = j0 + 8 < rows ? j0 + 8 : rows;
SNV_UNROLL(SNV_ROW_UNROLL)
                for (int j = j0; j < j1; ++j) {
                    const uint32_t key = row_key(j);
                    if (key < thr) {
                        my_list[cnt * kMaxBlock] = key;
                        ++cnt;
                    }
                }
                if (__any_sync(0xffffffffu, cnt > kListCap - 8)) fold();
            }
        } else {
SNV_UNROLL(SNV_ROW_UNROLL)
            for (int j = 0; j < rows; ++j) {
                const uint32_t key = row_key(j);
                if (key < best[KT - 1]) topk_insert<KT, uint32_t>(best, key);
            }
        }
#endif
        __syncthreads();  // everyone is done with this stage before it is refilled
    }
    if constexpr (KT >= kListKT) fold();

    if (!active) return;
    const uint32_t idx_mask = (1u << idx_bits) - 1u;
    const int64_t qrow = slot.qrow;
    if (p.nsplit == 1) {
#pragma unroll
        for (int i = 0; i < KT; ++i) {
            if (i < p.k) {
                const uint32_t key = best[i];
                const bool empty = key == kSent32;
                const int32_t dist = empty ? 0x7FFFFFFF : (int32_t)(key >> idx_bits);
                const int64_t id = empty ? -1 : (int64_t)(key & idx_mask) + r0 + p.id_offset;
                const int64_t o = qrow * p.k + i;
                if (p.D_i32) p.D_i32[o] = dist;
                if (p.D_f32) p.D_f32[o] = empty ? 3.4028234663852886e38f : (float)dist;
                p.I[o] = id;
            }
        }
    } else {
        uint64_t* out = p.partial + (qrow * p.nsplit + split) * KT;
#pragma unroll
        for (int i = 0; i < KT; ++i) {
            const uint32_t key = best[i];
            out[i] = key == kSent32
                         ? kSent64
                         : ((uint64_t)(key >> idx_bits) << 32) | (uint64_t)((key & idx_mask) + (uint32_t)r0);
        }
    }
}

// Generic fallback for rows wider than the register-resident instantiations: query words in
// shared memory ([word][thread], conflict-free), plain POPC per word.
template <bool MASKED, int KT>
__global__ void __launch_bounds__(32)
hamming_topk_generic_kernel(const HammingSearchParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
    uint32_t* tiles = reinterpret_cast<uint32_t*>(smem_raw + kBarBytes);
    const int tid = threadIdx.x;
    const int B = blockDim.x;
    const Slot slot = decode_slot(p);
    const int split = slot.split;
    const int w = slot.w;
    const int64_t r0 = (int64_t)split * p.rows_per_split;
    const int64_t r1 = (r0 + p.rows_per_split < p.n) ? r0 + p.rows_per_split : p.n;
    const int nrows = (int)(r1 - r0);
    const int TR = p.tile_rows;
    const int ntiles = (nrows + TR - 1) / TR;
    const int stages = p.stages;
    const uint32_t tile_words = (uint32_t)TR * p.stride;
    const uint32_t* pw = p.panel + (int64_t)w * p.panel_win_stride + r0 * p.stride;
    uint32_t* qs = tiles + (size_t)stages * tile_words;  // [words][B]
    uint32_t* ms = qs + (size_t)p.stride * B;

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
        fence_barrier_init();
    }
    __syncthreads();
    auto issue = [&](int t) {
        const int rows = (nrows - t * TR < TR) ? nrows - t * TR : TR;
        const uint32_t bytes = (uint32_t)rows * p.stride * 4u;
        uint64_t* bar = &bars[t % stages];
        mbar_arrive_expect_tx(bar, bytes);
        bulk_g2s(tiles + (size_t)(t % stages) * tile_words, pw + (size_t)t * tile_words, bytes, bar);
    };
    if (tid == 0) {
        for (int t = 0; t < stages - 1 && t < ntiles; ++t) issue(t);
    }
    const bool active = slot.active;
    for (int i = 0; i < p.stride; ++i) {
        uint32_t qv = 0, mv = 0;
        if (active) {
            qv = p.q[slot.qrow * p.stride + i];
            if constexpr (MASKED) {
                mv = p.mask[slot.mrow + i];
                qv &= mv;
            }
        }
        qs[(size_t)i * B + tid] = qv;
        if constexpr (MASKED) ms[(size_t)i * B + tid] = mv;
    }
    __syncthreads();

    uint32_t best[KT];
#pragma unroll
    for (int i = 0; i < KT; ++i) best[i] = kSent32;
    const int idx_bits = p.idx_bits;
    for (int t = 0; t < ntiles; ++t) {
        if (tid == 0 && t + stages - 1 < ntiles) issue(t + stages - 1);
        mbar_wait(&bars[t % stages], (uint32_t)(t / stages) & 1u);
        const uint32_t* tile = tiles + (size_t)(t % stages) * tile_words;
        const int rows = (nrows - t * TR < TR) ? nrows - t * TR : TR;
        for (int j = 0; j < rows; ++j) {
            const uint32_t* rp = tile + (size_t)j * p.stride;
            uint32_t dist = 0;
#pragma unroll 4
            for (int i = 0; i < p.words; ++i) {
                uint32_t x = rp[i];
                if constexpr (MASKED) x = (x & ms[(size_t)i * B + tid]) ^ qs[(size_t)i * B + tid];
                else                  x ^= qs[(size_t)i * B + tid];
                dist += __popc(x);
            }
            const uint32_t key = (dist << idx_bits) | (uint32_t)(t * TR + j);
            if (key < best[KT - 1]) topk_insert<KT, uint32_t>(best, key);
        }
        __syncthreads();
    }
    if (!active) return;
    const uint32_t idx_mask = (1u << idx_bits) - 1u;
    const int64_t qrow = slot.qrow;
    if (p.nsplit == 1) {
#pragma unroll
        for (int i = 0; i < KT; ++i) {
            if (i < p.k) {
                const uint32_t key = best[i];
                const bool empty = key == kSent32;
                const int32_t dist = empty ? 0x7FFFFFFF : (int32_t)(key >> idx_bits);
                const int64_t id = empty ? -1 : (int64_t)(key & idx_mask) + r0 + p.id_offset;
                const int64_t o = qrow * p.k + i;
                if (p.D_i32) p.D_i32[o] = dist;
                if (p.D_f32) p.D_f32[o] = empty ? 3.4028234663852886e38f : (float)dist;
                p.I[o] = id;
            }
        }
    } else {
        uint64_t* out = p.partial + (qrow * p.nsplit + split) * KT;
#pragma unroll
        for (int i = 0; i < KT; ++i) {
            const uint32_t key = best[i];
            out[i] = key == kSent32
                         ? kSent64
                         : ((uint64_t)(key >> idx_bits) << 32) | (uint64_t)((key & idx_mask) + (uint32_t)r0);
        }
    }
}

int bit_length(int64_t v)
{
    int b = 0;
    while (v > 0) { ++b; v >>= 1; }
    return b;
}

template <int NW, bool MASKED, int KT>
int launch_one(const HammingSearchParams& p, cudaStream_t stream)
{
    auto kern = hamming_topk_kernel<NW, MASKED, KT>;
    if (p.smem_bytes > 48 * 1024) {
        SNV_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes));
    }
    const int64_t grid = (p.work ? (int64_t)p.n_work : (int64_t)p.nw * p.qtiles) * p.nsplit;
    profile_begin(stream);
    kern<<<(unsigned)grid, p.block, p.smem_bytes, stream>>>(p);
    profile_end(stream);
    SNV_LAUNCH_CHECK();
    return SNV_OK;
}

template <bool MASKED, int KT>
int launch_generic(const HammingSearchParams& p, cudaStream_t stream)
{
    auto kern = hamming_topk_generic_kernel<MASKED, KT>;
    if (p.smem_bytes > 48 * 1024) {
        SNV_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes));
    }
    const int64_t grid = (p.work ? (int64_t)p.n_work : (int64_t)p.nw * p.qtiles) * p.nsplit;
    profile_begin(stream);
    kern<<<(unsigned)grid, p.block, p.smem_bytes, stream>>>(p);
    profile_end(stream);
    SNV_LAUNCH_CHECK();
    return SNV_OK;
}

template <int NW>
int launch_nw(const HammingSearchParams& p, cudaStream_t stream)
{
    const bool masked = p.mask != nullptr;
    if (p.kt == 8) return masked ? launch_one<NW, true, 8>(p, stream) : launch_one<NW, false, 8>(p, stream);
    return masked ? launch_one<NW, true, 32>(p, stream) : launch_one<NW, false, 32>(p, stream);
}

}  // namespace

// Register-resident instantiations: words -> NW template value (0 = generic fallback)
static int pick_nw(int words, int stride)
{
    if (words == 33 && stride == 36) return 33;
#ifdef SNV_TUNE_ONLY33
    return 0;
#endif
    switch (stride) {
        case 4: case 8: case 16: case 24: case 32: case 36: case 48: case 68: return stride;
        default: return 0;
    }
}

size_t hamming_plan(HammingSearchParams& p)
{
    if (p.k < 1 || p.k > 32) {
        set_error("hamming search: k must be in [1, 32] (got " + std::to_string(p.k) + ")");
        return (size_t)-1;
    }
    if (p.d >= (1 << 20)) {
        set_error("hamming search: d too large");
        return (size_t)-1;
    }
    p.kt = p.k <= 8 ? 8 : 32;
    p.wt1 = 1; p.wt2 = 2; p.wt4 = 4;
    p.nw_templ = pick_nw(p.words, p.stride);
    const int dist_bits = bit_length((int64_t)p.d + 1);
    p.idx_bits = 32 - dist_bits;
    const int64_t max_rows_per_split = (int64_t)1 << p.idx_bits;

    if (p.work) {
        p.block = p.work_block;  // chosen by the caller that built the work list
    } else if (p.nw_templ) {
        p.block = p.nq >= kMaxBlock ? kMaxBlock : (int)round_up(p.nq > 0 ? p.nq : 1, 32);
    } else {
        p.block = 32;
    }
    if (!p.nw_templ && (p.stride > 512 || p.block != 32)) {
        set_error(p.stride > 512 ? "hamming search: d > 16384 bits is not supported yet"
                                 : "hamming search: wide rows use 32-query blocks");
        return (size_t)-1;
    }
    p.qtiles = (int)ceil_div(p.nq > 0 ? p.nq : 1, p.block);

    // panel tile: ~16 KB per stage, 3 stages
    p.stages = 3;
    if (const char* e = getenv("SNV_STAGES")) p.stages = atoi(e) >= 2 ? atoi(e) : 3;
    int tr_budget = 16 * 1024;
    if (const char* e = getenv("SNV_TILE_BYTES")) tr_budget = atoi(e) >= 1024 ? atoi(e) : tr_budget;
    int tr = tr_budget / (p.stride * 4);
    tr = tr >= 256 ? 256 : (tr >= 128 ? 128 : (tr >= 64 ? 64 : (tr >= 32 ? 32 : (tr >= 16 ? 16 : 8))));
    p.tile_rows = tr;

    // row splits: several waves of CTAs (148 SMs x ~5-7 resident CTAs) so that the tail of the last
    // wave is short, and rows per split that fit the id field of the 32-bit key.
    const int64_t base = p.work ? (int64_t)p.n_work : (int64_t)p.nw * p.qtiles;
    const int64_t target = (int64_t)kNumSMs * 16;
    int64_t nsplit = 1;
    if (p.n > 0) {
        nsplit = base >= target ? 1 : ceil_div(target, base);
        const int64_t max_split = ceil_div(p.n, (int64_t)tr * 2);  // at least 2 tiles per split
        if (nsplit > max_split) nsplit = max_split;
        if (nsplit < 1) nsplit = 1;
        const int64_t need = ceil_div(p.n, max_rows_per_split);
        if (nsplit < need) nsplit = need;
        int64_t rps = round_up(ceil_div(p.n, nsplit), tr);
        if (rps > max_rows_per_split) rps = max_rows_per_split / tr * tr;
        if (rps < tr) {
            set_error("hamming search: d too large for the 32-bit key");
            return (size_t)-1;
        }
        nsplit = ceil_div(p.n, rps);
        p.rows_per_split = (int)rps;
    } else {
        p.rows_per_split = tr;
    }
    if (nsplit > 65535) {
        set_error("hamming search: panel too large for one call; shard rows");
        return (size_t)-1;
    }
    p.nsplit = (int)nsplit;
    p.smem_bytes = kBarBytes + (size_t)p.stages * p.tile_rows * p.stride * 4;
    if (!p.nw_templ) p.smem_bytes += (size_t)p.stride * p.block * 4 * (p.mask ? 2 : 1);
    else if (p.kt >= kListKT) p.smem_bytes += (size_t)kListCap * kMaxBlock * 4;
    if (base * p.nsplit > 0x7fffffffLL) {
        set_error("hamming search: grid too large");
        return (size_t)-1;
    }
    const int64_t nq_total = p.work ? p.nq_total : (int64_t)p.nw * p.nq;
    return p.nsplit > 1 ? (size_t)nq_total * p.nsplit * p.kt * sizeof(uint64_t) : 0;
}

int hamming_launch(const HammingSearchParams& p, cudaStream_t stream)
{
    if (!p.work && (p.nw <= 0 || p.nq <= 0)) return SNV_OK;
    if (p.work && p.n_work <= 0) return SNV_OK;
    int rc;
    switch (p.nw_templ) {
#ifndef SNV_TUNE_ONLY33
        case 4: rc = launch_nw<4>(p, stream); break;
        case 8: rc = launch_nw<8>(p, stream); break;
        case 16: rc = launch_nw<16>(p, stream); break;
        case 24: rc = launch_nw<24>(p, stream); break;
        case 32: rc = launch_nw<32>(p, stream); break;
#endif
        case 33: rc = launch_nw<33>(p, stream); break;
#ifndef SNV_TUNE_ONLY33
        case 36: rc = launch_nw<36>(p, stream); break;
        case 48: rc = launch_nw<48>(p, stream); break;
        case 68: rc = launch_nw<68>(p, stream); break;
#endif
        default: {
            const bool masked = p.mask != nullptr;
            if (p.kt == 8) rc = masked ? launch_generic<true, 8>(p, stream) : launch_generic<false, 8>(p, stream);
            else rc = masked ? launch_generic<true, 32>(p, stream) : launch_generic<false, 32>(p, stream);
        }
    }
    if (rc != SNV_OK) return rc;
    if (p.nsplit > 1) {
        return merge_keys_launch(p.partial, p.nsplit, p.kt, p.work ? p.nq_total : (int64_t)p.nw * p.nq, p.k, p.id_offset,
                                 false, p.D_i32, p.D_f32, p.I, stream);
    }
    return SNV_OK;
}

}  // namespace snv
