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
// panel k-block into a narrow-float operand tile directly in the SWIZZLE_128B shared-memory layout
// the MMA reads, with ONE LOP3 per output word: the site -> K position map is free to choose (a dot
// product does not care about the order of K), and a single set bit inside a narrow-float code is
// a power of two (E4M3 byte: bits 3..6 = 2^-6, 2^-5, 2^-3, 2^1; E2M1 nibble: bits 0..2 = 0.5, 1, 2).
// The query operand carries the inverse magnitude with the sign / mask folded in, so every product
// is exactly -1, 0 or +1 and the fp32 accumulation in TMEM is exact.
//
// Two operand formats (MODE):
//   fp8  kind::f8f6f4, E4M3, K = 32 sites per MMA, k-block = 4 packed words (128 sites), N = 256
//        byte position 32 i + 4 j + b  <->  word i, bit 8 b + j
//   fp4  kind::mxf4 (block scale = 1.0 everywhere), E2M1, K = 64 sites per MMA, k-block = 8 packed
//        words (256 sites), N = 240 (TMEM: 2 x 240 accumulator columns + 32 columns of unit scales)
//        nibble n of byte 16 i + 4 j + n / 2  <->  word i, bit 4 n + j
// Halving the operand bytes per site (fp4) is what pays: the single-CTA kernels are bound by shared-memory
// bandwidth (MMA operand reads + expander stores), not by the tensor pipe.
//
// CTA pairs (MODE_FP4_2CTA, the default when a window has two query tiles): two CTAs of a cluster hold 128 queries
// each (M = 256, tcgen05.mma.cta_group::2 issued by the leader) and each expands only half (120 rows) of every panel
// tile; slots are released and accumulators published with multicast tcgen05.commit, the peer's TMA completes on the
// leader's barrier, the peer's expanders / epilogue arrive on the leader's barriers through mapa.
//
// Query operand in tensor memory (MODE_FP4_2CTA_TA, engine 5: the DEFAULT for windows of up to 1216 sites; forced with
// SNV_HAMMING_ENGINE=tc4x2ta): the shared-memory port is what bounds the plain pair kernel (MMA operand reads + expander
// stores + TMA writes), and 38 % of that traffic is the query tile, re-fetched and re-read for every panel tile.  Here the
// epilogue warps build an item's query operand in TMEM once, straight from the PACKED query rows (global loads, in-register
// expansion, tcgen05.st; thread = query row = lane, 8 codes per column) and the MMAs take A from there (tcgen05.mma [d],
// [a], b-desc): no A ring, no A traffic, no expanded query operands in HBM, no side kernel.  TMEM = 2 x 160 accumulator
// columns + 160 operand columns + 32 scale columns, so tiles are 160 panel rows wide; the producers work per TILE (one
// raw TMA box of whole packed rows, one B slot of 5 k-block sub-tiles, one barrier wait and 18 MMAs per tile).
//
// CTA (480 threads, one per SM, persistent over (window, query tile [pair], row split) items):
//   warp 0       TMA producer of the query operand tile A [128 x 128 B] (SWIZZLE_128B; idle in the TMEM-A mode)
//   warp 1       TMEM allocator + MMA issuer (one elected lane): tcgen05.mma M128 / M256 into one of two accumulator stages
//   warp 2       TMA producer of the raw packed panel rows
//   warps 3-6    expanders: packed bits -> operand tile B (one 128-byte row per panel row and k-block) + the rows'
//                column-index code block (and, TMEM-A mode, the tile tag block)
//   warps 7-14   epilogue: thread = query (TMEM lane), two warps per lane quarter on the two column parts of a tile.
//                fp4 CTA-pair engines (candidate-list epilogue): one extra MMA per tile adds column / 256 (+ tile tag /
//                2048) to every accumulator, so a candidate is self-describing: the scan takes the columns in pairs
//                (minimum, compare, predicated 8-byte append to the thread's list in shared memory) and the lists are
//                folded into a sorted register top-k of 32-bit keys (distance << idx_bits | row) once per tile part
//                (k > 8: once per 4 tiles after the first 16).  Other engines: per column a compare, a predicated store
//                into the column's slot and a bit in a 32-column mask, popped in lockstep after each chunk.  The parts
//                exchange thresholds, are merged in shared memory and the final (D, I) rows are written by the kernel.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <type_traits>

#include "common.cuh"
#include "kernels.cuh"
#include "tcgen05.cuh"
#include "topk.cuh"

#ifndef SNV_TC_EPI16
#define SNV_TC_EPI16 0  // 1: 16 epilogue warps for k <= 8 in the CTA-pair kernel (measured slower on B200: 1.98 vs 1.76 ms per 296 windows)
#endif
#ifndef SNV_TC_LIST_EPI
#define SNV_TC_LIST_EPI 1  // fp4 engines: candidate-list epilogue (column index carried by the accumulator, folds deferred to tile ends)
#endif
#ifndef SNV_TC_LIST_CAP
#define SNV_TC_LIST_CAP 32  // pending entries (column pairs) a thread can hold (x 8 B x epilogue threads of shared memory)
#endif
#ifndef SNV_TC_FOLD_AT
#define SNV_TC_FOLD_AT 16   // fold inside a tile when some lane holds this many entries (<= SNV_TC_LIST_CAP - 16)
#endif
// Quad scan (four columns per test: FMNMX3 + FMNMX + FSETP, 16-byte list entries): 40 % fewer scan instructions, but measured
// SLOWER on B200 (profiles/r2_tc_quad_scan_variants.txt: cfg 2 1.518 -> 1.762 ms per 296 windows, k = 32 shard 1.244 -> 1.514 ms):
// every entry drags three more values through the fold's threshold checks.  Kept behind the flags, off.
#ifndef SNV_TC_QUAD_SCAN
#define SNV_TC_QUAD_SCAN 0
#endif
#ifndef SNV_TC_QUAD_SCAN_K32
#define SNV_TC_QUAD_SCAN_K32 0
#endif
#ifndef SNV_TC_DEFER_FOLD
#define SNV_TC_DEFER_FOLD 1  // TMEM-A engine: candidates of up to 8 tiles share one fold (each carries its tile's 3-bit tag, added by the index MMA)
#endif
#ifndef SNV_TC_TAG_SF
#define SNV_TC_TAG_SF 0x74777477u  // UE8M0 A scales of the index MMA's two 32-element blocks: 2^-8 (columns), 2^-11 (tile tag)
#endif
// Measured on B200 (profiles/r2_tc_fold_variants.txt): k = 32 on a 25,000-row shard 1.279 -> 1.236 ms with 16 warm tiles and
// windows of 4; every deferred setting LOSES at k <= 8 (cfg 2: 1.487 -> 1.50-1.57 ms; the insert is cheap there and a stale
// threshold admits more list entries), so the k <= 8 kernels keep one fold per tile.
#ifndef SNV_TC_FOLD_WARM
#define SNV_TC_FOLD_WARM 16   // ... after this many tiles of an item folded one by one (the threshold still falls fast there)
#endif
#ifndef SNV_TC_FOLD_WINDOW
#define SNV_TC_FOLD_WINDOW 4  // tiles per common fold: 1, 2, 4 or 8 (the tag has 3 bits)
#endif
#ifndef SNV_TC_FOLD_WARM_K8
#define SNV_TC_FOLD_WARM_K8 0   // the same two knobs for the k <= 8 kernels
#endif
#ifndef SNV_TC_FOLD_WINDOW_K8
#define SNV_TC_FOLD_WINDOW_K8 1
#endif
#ifndef SNV_TC_DEFER_PUBLISH
#define SNV_TC_DEFER_PUBLISH 1  // expanders publish a B slot one k-block late (its stores drain behind the next block's loads)
#endif
#ifndef SNV_TC_WARP_ARRIVE
#define SNV_TC_WARP_ARRIVE 1  // epilogue warps release an accumulator stage with one arrival per warp (0: one per thread)
#endif
#ifndef SNV_TC_DEFAULT_ENGINE
#define SNV_TC_DEFAULT_ENGINE 5  // what "auto" picks for tensor-core shapes: 1 = fp8, 3 = fp4, 4 = fp4 on CTA pairs, 5 = 4 with the query operand in tensor memory
#endif

namespace snv {

namespace {

using namespace tc;

constexpr int BM = 128;          // queries per tile (TMEM lanes)
constexpr int kRowBytes = 128;   // operand bytes per row and k-block: one 128-byte swizzle row
constexpr int kAccStages = 2;
constexpr int kTmemCols = 512;
constexpr int kExpWarps = 4;
constexpr int kExpThreads = kExpWarps * 32;
// Epilogue shape: 8 warps (2 per SM sub-partition) on 128-column parts of every tile, scored in 32-column groups.
// (16 warps on 64-column parts - SNV_TC_EPI16, CTA-pair kernel at k <= 8 only - were measured slower: the kernel is
// bound by issued instructions, more epilogue warps only add contention; profiles/r1_tc_breakdown_variants.txt.)
enum { MODE_FP8 = 0, MODE_FP8_HBM = 1, MODE_FP4 = 2, MODE_FP4_2CTA = 3, MODE_FP4_2CTA_TA = 4 };
#ifndef SNV_TC_TA_EPI12
#define SNV_TC_TA_EPI12 0  // TMEM-A engine at k <= 8: 12 epilogue warps (three column parts per lane quarter) instead of 8
#endif
template <int KT, int MODE>
struct Epi {
    static constexpr bool kPairMode = MODE == MODE_FP4_2CTA || MODE == MODE_FP4_2CTA_TA;
    static constexpr int kWarps = (MODE == MODE_FP4_2CTA_TA && KT == 8 && SNV_TC_TA_EPI12) ? 12
                                  : ((KT == 8 && kPairMode && SNV_TC_EPI16) ? 16 : 8);
    static constexpr int kThreads = kWarps * 32;
    static constexpr int kParts = kWarps / 4;          // warps per TMEM lane quarter = column parts of a tile
    static constexpr int kPartCols = 256 / (kParts == 3 ? 4 : kParts);     // 64 or 128 (smem-A engines)
    static constexpr int kGroup = 32;                  // columns scored between two folds
    static constexpr uint32_t kSlotStride = kThreads * 4;  // bytes between the candidate slots of consecutive columns
};
constexpr int kFirstExpWarp = 3;
constexpr int kFirstEpiWarp = kFirstExpWarp + kExpWarps;  // 7
template <int KT, int MODE>
constexpr int threads_of() { return 32 * (kFirstEpiWarp + Epi<KT, MODE>::kWarps); }  // 480 (608 / 736 with more epilogue warps)
constexpr uint32_t kABytes = BM * kRowBytes;   // 16 KB
constexpr uint32_t kBBytes = 256 * kRowBytes;  // 32 KB slot (240 or 256 rows in use)
#ifndef SNV_TC_BSTAGES
#define SNV_TC_BSTAGES 3
#endif
#ifndef SNV_TC_ASTAGES
#define SNV_TC_ASTAGES 4
#endif
#ifndef SNV_TC_RAWSTAGES
#define SNV_TC_RAWSTAGES 4
#endif
#ifndef SNV_TC_TA_BSTAGES
#define SNV_TC_TA_BSTAGES 2  // TMEM-A mode: B slots hold a whole tile (all k-blocks of the CTA's 80 rows, 50 KB each)
#endif
#ifndef SNV_TC_TA_RAWSTAGES
#define SNV_TC_TA_RAWSTAGES 3  // TMEM-A mode: raw slots hold the packed rows of a whole tile (80 rows x row stride)
#endif
constexpr size_t kListBytes = 32 * 1024;  // one slot per (epilogue thread, column of a group): 32 x 256 floats
constexpr size_t kListBytes16 = 64 * 1024; // 16 epilogue warps: 32 x 512 floats
static_assert((Epi<8, MODE_FP4_2CTA>::kParts - 1) * 8 * BM * 4 <= kListBytes && (Epi<32, MODE_FP8>::kParts - 1) * 32 * BM * 4 <= kListBytes,
              "the part-exchange buffer aliases the candidate slots");
static_assert((size_t)SNV_TC_LIST_CAP * 2048 >= (size_t)32 * BM * 4 && (size_t)24 * 8 * 384 >= (size_t)2 * 8 * BM * 4,
              "the part-exchange buffer aliases the candidate lists");

// Three rings: the query operand comes from L2 (long latency: deep ring), the panel operand is made
// in the SM (expander latency: 3 slots), raw packed k-blocks are small.
template <int MODE>
struct Cfg {
    static constexpr bool kFp4 = MODE == MODE_FP4 || MODE == MODE_FP4_2CTA || MODE == MODE_FP4_2CTA_TA;
    // CTA pair (cta_group::2): two CTAs of a cluster hold 128 queries each (M = 256) and each expands HALF of every
    // panel tile; the pair's MMA reads both halves.  Halves the expander work and the B operand bytes per SM.
    static constexpr bool kTwoCta = MODE == MODE_FP4_2CTA || MODE == MODE_FP4_2CTA_TA;
    // Query operand in TENSOR MEMORY (bring-up, SNV_HAMMING_ENGINE=tc4x2ta): the 128-query tile of an item (up to 5
    // k-blocks = 160 columns) is written to TMEM once per item by the epilogue warps and the MMAs take A from there,
    // so there is no A ring, no per-tile TMA traffic for it and no shared-memory operand reads for A.
    // TMEM: 2 x 160 accumulator columns | 160 operand columns | 32 columns of unit scales.
    static constexpr bool kTmemA = MODE == MODE_FP4_2CTA_TA;
    static constexpr int kMaxKbTmemA = 5;
    static constexpr bool kExpand = MODE != MODE_FP8_HBM;
    // Candidate-list epilogue (fp4 engines).  One extra MMA per tile adds (panel row within the tile) / 256 to every
    // accumulator - operand codes 1.0 on the query side, a constant per-row code block on the panel side, block scale
    // 2^-8 - so a candidate's value carries its own column: the scan is compare + predicated append to a per-thread
    // list, and the lists are folded into the sorted top-k once per tile part (or when a list fills) instead of after
    // every 32 columns.
    static constexpr bool kListEpi = kFp4 && kTwoCta && SNV_TC_LIST_EPI;  // (the single-CTA kernel has no shared memory left for lists)
    static constexpr bool kEpi12 = MODE == MODE_FP4_2CTA_TA && SNV_TC_TA_EPI12;   // (k <= 8 kernels; sizes below cover both)
    // Deferred folds (TMEM-A engine): the second 32-element block of the index MMA's k-chunk - all zero in the row codes,
    // column indices need 27 elements - carries  (tile & 7)  at an A scale of 2^-11, so an accumulator reads
    // a + column / 256 + (tile & 7) / 2048  and list entries of up to 8 consecutive tiles can wait for one common fold:
    // late in an item a tile leaves candidates in only a few lanes, and a lockstep fold per tile costs a full insertion
    // round for each of them.
    static constexpr bool kDeferFold = MODE == MODE_FP4_2CTA_TA && SNV_TC_LIST_EPI && SNV_TC_DEFER_FOLD && !(SNV_TC_TA_EPI12);
    static constexpr int kListCap = kEpi12 ? 24 : SNV_TC_LIST_CAP;
    static constexpr int kFoldAt = kEpi12 ? 8 : SNV_TC_FOLD_AT;
    static constexpr int kPartsMax = kEpi12 ? 3 : 2;
    static_assert(!kListEpi || kFoldAt + 16 <= kListCap, "a 32-column group (16 pairs) must always fit behind the fold mark");
    static constexpr int BN = kTmemA ? 160 : (kFp4 ? 240 : 256);   // panel rows per tile = TMEM columns per accumulator stage
    static constexpr int WPK = kFp4 ? 8 : 4;      // packed words per k-block
    static constexpr int WPM = kFp4 ? 2 : 1;      // packed words per MMA
    static constexpr int kAStages = kTmemA ? 0 : SNV_TC_ASTAGES;
    static constexpr int kBStages = kTmemA ? SNV_TC_TA_BSTAGES : SNV_TC_BSTAGES;
    // bytes in front of the B ring: the query tile ring (none in the TMEM-A mode)
    static constexpr size_t kAOpBytes = (size_t)kAStages * kABytes;
    static_assert(kAOpBytes % 1024 == 0, "the B ring stays 1024-byte aligned");
    static constexpr int kRawStages = kTmemA ? ((MODE == MODE_FP4_2CTA_TA && SNV_TC_TA_EPI12) ? 2 : SNV_TC_TA_RAWSTAGES) : SNV_TC_RAWSTAGES;
    static constexpr int kBRows = kTwoCta ? BN / 2 : BN;          // panel rows this CTA expands per tile
    static constexpr uint32_t kRawRow = WPK * 4;                // raw bytes per panel row and k-block
    // TMEM-A mode works per TILE, not per k-block: one raw TMA box brings the 80 packed rows whole (row stride <= 48
    // words), the expanders fill one B slot = kMaxKbTmemA sub-tiles of [80 rows x 128 B] (SWIZZLE_128B each), the MMA
    // warp issues every MMA of the tile behind one barrier wait
    static constexpr int kMaxStrideTmemA = 48;
    static constexpr uint32_t kBSub = (uint32_t)kBRows * kRowBytes;   // one k-block sub-tile of a per-tile B slot
    static_assert(!kTmemA || kBSub % 1024 == 0, "sub-tiles stay swizzle-atom aligned");
    static constexpr uint32_t kRawSlot = !kExpand ? 0 : (kTmemA ? (uint32_t)(kBRows * kMaxStrideTmemA * 4 + 1024) : (kTwoCta ? 128 : 256) * kRawRow);
    static constexpr uint32_t kBSlot = kTmemA ? kMaxKbTmemA * kBSub : (kTwoCta ? kBBytes / 2 : kBBytes);  // 16 KB holds the pair kernel's 120 rows
    static constexpr uint32_t kRawBytes = kBRows * kRawRow;     // what one TMA box brings (TMEM-A mode: rows x stride, a runtime value)
    static constexpr uint32_t kBBox = BN * kRowBytes;           // fp8-hbm variant: one TMA box of operand rows
    static constexpr uint32_t kACol = 2 * BN;                   // TMEM-A mode: first TMEM column of the query operand
    static constexpr uint32_t kSfCol = kTmemA ? 480 : 2 * BN;   // fp4: first TMEM column of the unit scales
    static_assert(!kTmemA || kACol + 32 * kMaxKbTmemA <= kSfCol, "TMEM budget");
    // mbarriers of the rings + TMEM stages, the TMEM base / runtime-one words and the a_full barrier
    static constexpr size_t kBarBytes = ((size_t)(2 * (kAStages + kBStages + kRawStages + kAccStages)) * 8 + 16 + 255) / 256 * 256;
    // candidate slots / lists of the epilogue threads (8 warps; 16 with SNV_TC_EPI16 in the pair kernels at k <= 8)
    static constexpr size_t kSlotBytes = kListEpi ? (size_t)kListCap * 8 * (kEpi12 ? 384 : ((kTwoCta && SNV_TC_EPI16) ? 512 : 256))
                                                  : ((kTwoCta && SNV_TC_EPI16) ? kListBytes16 : kListBytes);
    static constexpr size_t kQbBytes = kTmemA ? 2 * kPartsMax * BM * 4 + (size_t)kBRows * 32 : 0;  // TMEM-A mode: partial query biases
                                                                    // [item parity][part][row], then the column-index codes [row][32 B]
    static constexpr size_t kSmem = 1024 /*align slack*/ + kAOpBytes + (size_t)kBStages * kBSlot +
                                    (size_t)kRawStages * kRawSlot + kSlotBytes + (kTwoCta ? 4 : 2) * BM * 4 /*thresholds*/ + kQbBytes + kBarBytes;
    static_assert(kSmem <= 232448, "shared memory budget");
};

// E4M3 codes.  Panel side: the bit itself, moved (if needed) to one of bit positions 3..6 of its byte;
// query side: the inverse power of two, sign bit = allele 1, zero = unobserved site.
//   j (bit inside the byte):   0     1     2     3     4     5     6     7
//   panel code              0x10  0x20  0x40  0x08  0x10  0x20  0x40  0x08   (2^-5 2^-3 2^1 2^-6 ...)
//   query code              0x60  0x50  0x30  0x68  0x60  0x50  0x30  0x68   (32   8    0.5 64   ...)
__host__ __device__ __forceinline__ void expand_panel_word_fp8(uint32_t w, uint4& c0, uint4& c1)
{
    const uint32_t lo = w << 4, hi = w >> 4;
    c0 = make_uint4(lo & 0x10101010u, lo & 0x20202020u, lo & 0x40404040u, w & 0x08080808u);
    c1 = make_uint4(w & 0x10101010u, w & 0x20202020u, w & 0x40404040u, hi & 0x08080808u);
}

__host__ __device__ __forceinline__ uint32_t query_code_fp8(int j)
{
    switch (j & 3) {
        case 0: return 0x60u;
        case 1: return 0x50u;
        case 2: return 0x30u;
        default: return 0x68u;
    }
}

// E2M1 codes (nibble = sign | 2 exponent bits | 1 mantissa bit: 1 = 0.5, 2 = 1.0, 4 = 2.0).
//   j (bit inside the nibble):  0    1    2    3 (moved to bit 2)
//   panel code                  1    2    4    4        (0.5  1  2  2)
//   query code                  4    2    1    1        (2    1  0.5 0.5), | 8 for allele 1
__host__ __device__ __forceinline__ uint4 expand_panel_word_fp4(uint32_t w)
{
    return make_uint4(w & 0x11111111u, w & 0x22222222u, w & 0x44444444u, (w >> 1) & 0x44444444u);
}

__host__ __device__ __forceinline__ uint32_t query_code_fp4(int j) { return j == 0 ? 4u : (j == 1 ? 2u : 1u); }

// One 16-byte chunk of a query operand row: fp4 - chunk <- one packed word (wq = q & m, wm = m), h unused;
// fp8 - chunk pair (h = 0, 1) <- one packed word.  Code = magnitude where the site is observed, sign bit = allele 1.
template <bool FP4>
__host__ __device__ __forceinline__ uint4 expand_query_chunk(uint32_t wq, uint32_t wm, int h)
{
    uint32_t out[4];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int jj = 0; jj < 4; ++jj) {
        if constexpr (FP4) {
            const uint32_t mb = (wm >> jj) & 0x11111111u;
            const uint32_t sb = (wq >> jj) & 0x11111111u;
            out[jj] = mb * query_code_fp4(jj) | sb * 8u;
        } else {
            const int j = 4 * h + jj;
            const uint32_t mb = (wm >> j) & 0x01010101u;
            const uint32_t sb = (wq >> j) & 0x01010101u;
            out[jj] = mb * query_code_fp8(j) | sb * 0x80u;
        }
    }
    return make_uint4(out[0], out[1], out[2], out[3]);
}

struct TcParams {
    int nw, nq, qtiles;      // qtiles: query tiles per window (tile PAIRS in the 2-CTA mode)
    int64_t n;               // panel rows per window
    int words, kblocks;      // packed words in use (= MMAs per tile), k-blocks of 4 words
    int n_tiles, nsplit, tiles_per_split;
    int items;               // nw * qtiles * nsplit (+ the extra pieces of the tail items)
    // tail split (nsplit == 1 only): the last `items - items_a` pieces belong to trailing (window, query tile) items
    // that are cut into tail_split row ranges of tail_tiles tiles, so that the last round of items fills every SM
    int items_a, tail_split, tail_tiles;
    int64_t tail_row0;       // first query row of the tail items (their partial keys are indexed from it)
    int idx_bits, k, one;
    int idx_slot;            // list epilogue: MMA slot (32 operand bytes) of the column-index block, = ceil(words / 2); else -1
    int64_t id_offset;
    const int32_t* q_bias;   // [nw * nq] popc(q & m)
    int32_t* D_i32;
    float* D_f32;
    int64_t* I;
    uint64_t* partial;       // [nw * nq][nsplit][kt] when nsplit > 1
    const uint8_t* q_ops;    // [nw * nq][kblocks * 128 B] query operand rows (smem-A engines)
    // TMEM-A mode: the epilogue warps build the query operand from the packed rows themselves
    const uint32_t* q;       // [nw][nq][stride] packed queries
    const uint32_t* mask;    // nullptr or observed-site rows
    int64_t mask_win_stride, mask_q_stride;
    int stride, d;
};

struct Item {
    int w, qt, split, t0, ntiles;
    int out_split;  // 1: this piece covers the whole panel and writes final results; > 1: partial keys [row][out_split][kt]
};
__host__ __device__ __forceinline__ Item decode_item(const TcParams& p, int item, int qt_mul = 1, int qt_add = 0)
{
    Item it;
    if (p.tail_split > 1 && item >= p.items_a) {
        const int piece = item - p.items_a;
        it.split = piece % p.tail_split;
        const int base_item = p.items_a + piece / p.tail_split;
        it.qt = (base_item % p.qtiles) * qt_mul + qt_add;
        it.w = base_item / p.qtiles;
        it.t0 = it.split * p.tail_tiles;
        const int t1 = it.t0 + p.tail_tiles < p.n_tiles ? it.t0 + p.tail_tiles : p.n_tiles;
        it.ntiles = t1 - it.t0;
        it.out_split = p.tail_split;
        return it;
    }
    it.split = item % p.nsplit;
    item /= p.nsplit;
    it.qt = (item % p.qtiles) * qt_mul + qt_add;
    it.w = item / p.qtiles;
    it.t0 = it.split * p.tiles_per_split;
    const int t1 = it.t0 + p.tiles_per_split < p.n_tiles ? it.t0 + p.tiles_per_split : p.n_tiles;
    it.ntiles = t1 - it.t0;
    it.out_split = p.nsplit;
    return it;
}

__device__ __forceinline__ void umma_mxf4_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                               uint32_t tmem_sfa, uint32_t tmem_sfb)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
        : "memory");
}

// A operand from tensor memory (row = lane, 8 packed E2M1 codes per 32-bit column, 8 columns per K = 64 MMA)
__device__ __forceinline__ void umma_mxf4_2cta_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                                  uint32_t tmem_sfa, uint32_t tmem_sfb)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.block32 [%0], [%1], %2, %3, [%5], [%6], p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
        : "memory");
}

__device__ __forceinline__ void tmem_st_32x32b_x16v(uint32_t taddr, const uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}

// Column-index code block of panel row c of a tile (list epilogue): 64 E2M1 values that sum to c - c / 6 sixes (code 7),
// then the remainder as 0, 1, 2, 3, 4 or 4 + 1 - against 1.0 on the query side and an A scale of 2^-8.
__device__ __forceinline__ void index_code_words(int c, uint32_t (&wds)[8])
{
    const int n6 = c / 6, rem = c - 6 * n6;
    const uint32_t rcode = rem == 0 ? 0u : (rem == 1 ? 2u : (rem == 2 ? 4u : (rem == 3 ? 5u : 6u)));   // 0 1 2 3 4 (4)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int full = n6 - 8 * i;                                  // nibbles of this word below position n6
        uint32_t wv = full >= 8 ? 0x77777777u : (full > 0 ? 0x77777777u & ((1u << (4 * full)) - 1u) : 0u);   // 6.0 each
        if (full >= 0 && full < 8) wv |= rcode << (4 * full);                              // the remainder at position n6
        if (rem == 5 && full + 1 >= 0 && full + 1 < 8) wv |= 2u << (4 * (full + 1));       // + 1.0 at position n6 + 1
        wds[i] = wv;
    }
}

// compile-time unrolled loop: f(std::integral_constant<int, 0>{}), ..., f(std::integral_constant<int, N - 1>{})
template <int N, int I = 0, typename F>
__device__ __forceinline__ void static_for(F&& f)
{
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<N, I + 1>(f);
    }
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

// fp4: tcgen05.mma kind::mxf4 with block scaling; every scale factor in TMEM is 2^0
__device__ __forceinline__ void umma_mxf4(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                          uint32_t tmem_sfa, uint32_t tmem_sfb)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
        : "memory");
}

// block-scaled instruction descriptor (kind::mxf4): A = B = E2M1 (format 1 at bits 7 and 10), K-major,
// N >> 3 at bit 17, scale format UE8M0 (bit 23), M >> 4 at bit 24, scale-factor ids 0, K = 64 (bit 31 = 0)
__host__ __device__ constexpr uint32_t make_idesc_mxf4(int m, int n)
{
    return (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (1u << 23) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, uint32_t v)
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
        ::"r"(taddr), "r"(v) : "memory");
}

[[maybe_unused]] __device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
[[maybe_unused]] __device__ __forceinline__ void tmem_ld_wait16(uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.wait::ld.sync.aligned;"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
          "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
        :
        : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[N])
{
    if constexpr (N == 16) tmem_ld_32x32b_x16(taddr, r);
    else tmem_ld_32x32b_x32(taddr, r);
}
template <int N>
__device__ __forceinline__ void tmem_ld_wait_cols(uint32_t (&r)[N])
{
    if constexpr (N == 16) tmem_ld_wait16(r);
    else tmem_ld_wait(r);
}

template <int KT, int MODE>
__global__ void __launch_bounds__(threads_of<KT, MODE>(), 1)
hamming_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_r, const TcParams p)
{
    using C = Cfg<MODE>;
    constexpr bool EXPAND = C::kExpand;
    constexpr bool FP4 = C::kFp4;
    constexpr int BN = C::BN;
    constexpr int WPK = C::WPK;
    constexpr int kAStages = C::kAStages;
    constexpr int kRawStages = C::kRawStages;
    constexpr int kBStages = C::kBStages;
    constexpr bool TWO = C::kTwoCta;
    constexpr bool TA = C::kTmemA;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    unsigned char* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);  // SWIZZLE_128B tiles: 1024-byte aligned
    unsigned char* a_tiles = smem;
    unsigned char* b_tiles = a_tiles + C::kAOpBytes;
    unsigned char* raws = b_tiles + (size_t)kBStages * C::kBSlot;
    uint32_t* lists = reinterpret_cast<uint32_t*>(raws + (size_t)kRawStages * C::kRawSlot);  // [group columns][epilogue threads]
    using E = Epi<KT, MODE>;
    constexpr int kEpiThreads = E::kThreads;
    uint32_t* xchg = lists;                                                                 // [parts - 1][KT][128], after the slots are folded
    constexpr size_t kSlots = C::kSlotBytes;
    volatile float* thrx = reinterpret_cast<float*>(lists + kSlots / 4);               // [parts][128 queries] published thresholds
    [[maybe_unused]] volatile int32_t* qbx = reinterpret_cast<int32_t*>(lists + kSlots / 4 + (C::kTwoCta ? 4 : 2) * BM);  // TMEM-A: [2][2][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(lists + kSlots / 4 + (C::kTwoCta ? 4 : 2) * BM + C::kQbBytes / 4);
    uint64_t* full_a = bars;                        // [kAStages]   TMA -> MMA
    uint64_t* empty_a = full_a + kAStages;          // [kAStages]   MMA -> TMA
    uint64_t* full_b = empty_a + kAStages;          // [kBStages]   expanders (or TMA) -> MMA
    uint64_t* empty_b = full_b + kBStages;          // [kBStages]   MMA -> expanders (or TMA)
    uint64_t* raw_full = empty_b + kBStages;        // [kRawStages] TMA -> expanders
    uint64_t* raw_empty = raw_full + kRawStages;    // [kRawStages] expanders -> TMA
    uint64_t* tmem_full = raw_empty + kRawStages;   // [2] MMA -> epilogue
    uint64_t* tmem_empty = tmem_full + kAccStages;  // [2] epilogue -> MMA
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + kAccStages);
    [[maybe_unused]] uint64_t* a_full = reinterpret_cast<uint64_t*>(tmem_ptr + 2);  // TMEM-A mode: epilogue warps -> MMA, once per item

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // CTA pair: rank 0 is the leader (issues the MMAs; its barriers collect both CTAs' operands), items are
    // (window, query-tile PAIR, row split) and CTA r takes query tile 2 * pair + r
    const uint32_t cta_rank = TWO ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0u;
    const int item0 = TWO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int item_step = TWO ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    constexpr int kQtMul = TWO ? 2 : 1;
    constexpr uint32_t kPair = TWO ? 2u : 1u;

    if (warp == 0 && lane == 0) {
        if constexpr (!TA) prefetch_tensormap(&map_q);
        prefetch_tensormap(&map_r);
        for (int s = 0; s < kAStages; ++s) {
            mbar_init(&full_a[s], 1);
            mbar_init(&empty_a[s], 1);
        }
        for (int s = 0; s < kBStages; ++s) {
            mbar_init(&full_b[s], EXPAND ? kPair * kExpWarps : 1);
            mbar_init(&empty_b[s], 1);
        }
        for (int s = 0; s < kRawStages; ++s) {
            mbar_init(&raw_full[s], 1);
            mbar_init(&raw_empty[s], kExpWarps);
        }
        for (int s = 0; s < kAccStages; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], kPair * ((SNV_TC_WARP_ARRIVE || C::kListEpi) ? E::kWarps : kEpiThreads));
        }
        if constexpr (TA) mbar_init(a_full, kPair * E::kWarps);
        fence_barrier_init();
    }
    if (warp == 1) {
        if constexpr (TWO) tmem_alloc_2cta(tmem_ptr, kTmemCols);
        else tmem_alloc(tmem_ptr, kTmemCols);
    }
    if (warp == 2 && lane == 0) tmem_ptr[1] = (uint32_t)p.one;
    tcgen05_fence_before();
    __syncthreads();
    if constexpr (TWO) cluster_sync();  // the peer's barriers are initialised before anything arrives on them
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    if constexpr (FP4) {
        // unit block scales (UE8M0 0x7F = 2^0) in the 32 TMEM columns behind the accumulators, all 128 lanes
        if (warp >= kFirstExpWarp && warp < kFirstExpWarp + 4) {
            // columns [0, 8) and [16, 32): 2^0 (A / B scales of the packed-site MMAs); [8, 16): 2^-8 (UE8M0 0x77), the A
            // scale of the column-index MMA of the list epilogue
            const uint32_t t = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + C::kSfCol;
            uint32_t sfv[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) sfv[i] = (C::kListEpi && i >= 8) ? (C::kDeferFold ? SNV_TC_TAG_SF : 0x77777777u) : 0x7F7F7F7Fu;
            tmem_st_32x32b_x16v(t, sfv);
            tmem_st_32x32b_x16(t + 16u, 0x7F7F7F7Fu);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        tcgen05_fence_before();
        __syncthreads();
        tcgen05_fence_after();
    }
    const int KB = p.kblocks;

    if (warp == 0) {
        // ================= TMA producer: query operand tiles =================
        if (!TA && lane == 0) {
            Ring<(kAStages > 0 ? kAStages : 1)> ra;
            for (int item = item0; item < p.items; item += item_step) {
                const Item it = decode_item(p, item, kQtMul, (int)cta_rank);
                const int row_a = it.w * p.nq + it.qt * BM;
                for (int t = 0; t < it.ntiles; ++t) {
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait_relaxed(&empty_a[ra.i], ra.phase ^ 1u);
                        if constexpr (TWO) {
                            // both CTAs' tiles complete on the leader's barrier; only the leader arms it
                            if (leader) mbar_arrive_expect_tx(&full_a[ra.i], 2 * kABytes);
                            tma_load_2d_2cta(a_tiles + (size_t)ra.i * kABytes, &map_q, kb * kRowBytes, row_a, &full_a[ra.i]);
                        } else {
                            mbar_arrive_expect_tx(&full_a[ra.i], kABytes);
                            tma_load_2d(a_tiles + (size_t)ra.i * kABytes, &map_q, kb * kRowBytes, row_a, &full_a[ra.i]);
                        }
                        ra.next();
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer: the warp loops in lockstep, one elected lane issues =================
        if (leader) {
            constexpr int MM = TWO ? 2 * BM : BM;
            constexpr uint32_t idesc = FP4 ? make_idesc_mxf4(MM, BN) : make_idesc_e4m3(MM, BN);
            constexpr uint64_t kDescHi = (uint64_t)0x40004040u << 32;
            // K-major SWIZZLE_128B descriptors: low word = (address >> 4) | LBO 1 << 16, high word constant
            // (SBO 1024 B, descriptor version 1, layout SWIZZLE_128B); one MMA per 32 operand bytes = per packed
            // word (fp8) or per two packed words (fp4)
            const uint32_t a_lo0 = TA ? tmem_base + C::kACol : ((smem_u32(a_tiles) >> 4) | 0x10000u);  // TMEM-A: a column address
            const uint32_t b_lo0 = (smem_u32(b_tiles) >> 4) | 0x10000u;
            // MMAs of the last k-block over packed sites (list epilogue: 0..3, followed by the column-index MMA)
            const int nm_tail = (p.words - WPK * (KB - 1) + C::WPM - 1) / C::WPM;
            const uint32_t sf_a0 = tmem_base + C::kSfCol, sf_b = tmem_base + C::kSfCol + 16u;
            (void)sf_a0; (void)sf_b;
            auto mma = [&](uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t acc, uint32_t sf_a_off = 0u) {
                const uint32_t sf_a = sf_a0 + sf_a_off;
                (void)sf_a;
#ifdef TC_DEBUG_NO_MMA
                return;
#endif
                const uint64_t adesc = kDescHi | (uint64_t)a_lo;
                const uint64_t bdesc = kDescHi | (uint64_t)b_lo;
                if constexpr (TA) umma_mxf4_2cta_ta(d_tmem, a_lo, bdesc, idesc, acc, sf_a, sf_b);
                else if constexpr (TWO) umma_mxf4_2cta(d_tmem, adesc, bdesc, idesc, acc, sf_a, sf_b);
                else if constexpr (FP4) umma_mxf4(d_tmem, adesc, bdesc, idesc, acc, sf_a, sf_b);
                else umma_f8(d_tmem, adesc, bdesc, idesc, acc);
            };
            auto commit = [&](uint64_t* bar) {
#ifdef TC_DEBUG_PLAIN_COMMIT
                mbar_arrive(bar);
                if constexpr (TWO) mbar_arrive_cluster(bar, 1u);
                return;
#endif
                if constexpr (TWO) umma_commit_2cta(bar);
                else umma_commit(bar);
            };
            Ring<(kAStages > 0 ? kAStages : 1)> ra;
            Ring<kBStages> rb;
            uint32_t tcount = 0;
            [[maybe_unused]] uint32_t icount = 0;
            // operand step between the 4 MMAs of a k-block: 32 bytes of the swizzled row (descriptor units of 16 B),
            // or 8 TMEM columns
            constexpr uint32_t kAStep = TA ? 8u : 2u;
            if constexpr (TA) {
                // per-tile pipeline: one wait for the tile's operand slot, every MMA of the tile, two commits
                const int nslots = p.idx_slot + 1;  // MMA slots of 64 sites; the last one is the column-index block
                for (int item = item0; item < p.items; item += item_step) {
                    const Item it = decode_item(p, item, kQtMul, (int)cta_rank);
                    // both CTAs' epilogue warps have written this item's query tile to tensor memory
                    mbar_wait(a_full, icount & 1u);
                    ++icount;
                    tcgen05_fence_after();
                    for (int t = 0; t < it.ntiles; ++t, ++tcount) {
                        const uint32_t as = tcount & 1u;
                        mbar_wait(&tmem_empty[as], ((tcount >> 1) & 1u) ^ 1u);
                        mbar_wait(&full_b[rb.i], rb.phase);
                        tcgen05_fence_after();
                        if (elect_one()) {
                            const uint32_t d_tmem = tmem_base + as * BN;
                            const uint32_t b_lo = b_lo0 + (uint32_t)rb.i * (C::kBSlot >> 4);
                            for (int sl = 0; sl < nslots - 1; ++sl)
                                mma(d_tmem, a_lo0 + 8u * (uint32_t)sl, b_lo + (uint32_t)(sl >> 2) * (C::kBSub >> 4) + 2u * (uint32_t)(sl & 3), sl ? 1u : 0u);
                            const int sl = nslots - 1;
                            mma(d_tmem, a_lo0 + 8u * (uint32_t)sl, b_lo + (uint32_t)(sl >> 2) * (C::kBSub >> 4) + 2u * (uint32_t)(sl & 3), 1u, 8u);
                            commit(&empty_b[rb.i]);
                            commit(&tmem_full[as]);
                        }
                        __syncwarp();
                        rb.next();
                    }
                }
            } else
            for (int item = item0; item < p.items; item += item_step) {
                const Item it = decode_item(p, item, kQtMul, (int)cta_rank);
                for (int t = 0; t < it.ntiles; ++t, ++tcount) {
                    const uint32_t as = tcount & 1u;
                    mbar_wait(&tmem_empty[as], ((tcount >> 1) & 1u) ^ 1u);
                    tcgen05_fence_after();
                    const uint32_t d_tmem = tmem_base + as * BN;
                    for (int kb = 0; kb < KB; ++kb) {
                        if constexpr (!TA) mbar_wait(&full_a[ra.i], ra.phase);
                        mbar_wait(&full_b[rb.i], rb.phase);
                        tcgen05_fence_after();
                        if (elect_one()) {
                            const uint32_t a_lo = TA ? a_lo0 + (uint32_t)kb * 32u : a_lo0 + (uint32_t)ra.i * (kABytes >> 4);
                            const uint32_t b_lo = b_lo0 + (uint32_t)rb.i * (C::kBSlot >> 4);
                            if (kb != KB - 1) {
                                mma(d_tmem, a_lo, b_lo, kb != 0 ? 1u : 0u);
                                mma(d_tmem, a_lo + kAStep, b_lo + 2, 1u);
                                mma(d_tmem, a_lo + 2 * kAStep, b_lo + 4, 1u);
                                mma(d_tmem, a_lo + 3 * kAStep, b_lo + 6, 1u);
                                if constexpr (!TA) commit(&empty_a[ra.i]);
                                commit(&empty_b[rb.i]);
                            } else {
                                if constexpr (C::kListEpi) {
                                    // packed-site MMAs of the last k-block, then the column-index MMA (A scale 2^-8)
                                    uint32_t acc = kb != 0 ? 1u : 0u;
                                    if (nm_tail > 0) { mma(d_tmem, a_lo, b_lo, acc); acc = 1u; }
                                    if (nm_tail > 1) mma(d_tmem, a_lo + kAStep, b_lo + 2, 1u);
                                    if (nm_tail > 2) mma(d_tmem, a_lo + 2 * kAStep, b_lo + 4, 1u);
                                    mma(d_tmem, a_lo + (uint32_t)nm_tail * kAStep, b_lo + 2u * (uint32_t)nm_tail, acc, 8u);
                                } else {
                                    mma(d_tmem, a_lo, b_lo, kb != 0 ? 1u : 0u);
                                    if (nm_tail > 1) mma(d_tmem, a_lo + kAStep, b_lo + 2, 1u);
                                    if (nm_tail > 2) mma(d_tmem, a_lo + 2 * kAStep, b_lo + 4, 1u);
                                    if (nm_tail > 3) mma(d_tmem, a_lo + 3 * kAStep, b_lo + 6, 1u);
                                }
                                if constexpr (!TA) commit(&empty_a[ra.i]);
                                commit(&empty_b[rb.i]);
                                commit(&tmem_full[as]);
                            }
                        }
                        __syncwarp();
                        if constexpr (!TA) ra.next();
                        rb.next();
                    }
                }
            }
        }
    } else if (warp == 2) {
        // ================= TMA producer: raw packed panel k-blocks (fp8 panel tiles in the bring-up variant) =================
        if (lane == 0) {
            Ring<kRawStages> rr;
            Ring<kBStages> rb;
            for (int item = item0; item < p.items; item += item_step) {
                const Item it = decode_item(p, item, kQtMul, (int)cta_rank);
                for (int t = 0; t < it.ntiles; ++t) {
                    const int n0 = (it.t0 + t) * BN + (int)cta_rank * C::kBRows;  // this CTA's rows of the tile
                    if constexpr (TA) {
                        // the CTA's 80 packed rows of the tile, whole (box = row stride x 80 rows)
                        mbar_wait_relaxed(&raw_empty[rr.i], rr.phase ^ 1u);
                        mbar_arrive_expect_tx(&raw_full[rr.i], (uint32_t)(C::kBRows * p.stride * 4));
                        tma_load_3d(raws + (size_t)rr.i * C::kRawSlot, &map_r, 0, n0, it.w, &raw_full[rr.i]);
                        rr.next();
                    } else
                    for (int kb = 0; kb < KB; ++kb) {
                        if constexpr (EXPAND) {
                            mbar_wait_relaxed(&raw_empty[rr.i], rr.phase ^ 1u);
#ifdef TC_DEBUG_NO_RAW
                            mbar_arrive(&raw_full[rr.i]);
#else
                            mbar_arrive_expect_tx(&raw_full[rr.i], C::kRawBytes);
                            tma_load_3d(raws + (size_t)rr.i * C::kRawSlot, &map_r, kb * WPK, n0, it.w, &raw_full[rr.i]);
#endif
                            rr.next();
                        } else {
                            mbar_wait_relaxed(&empty_b[rb.i], rb.phase ^ 1u);
                            mbar_arrive_expect_tx(&full_b[rb.i], C::kBBox);
                            tma_load_3d(b_tiles + (size_t)rb.i * C::kBSlot, &map_r, kb * kRowBytes, n0, it.w, &full_b[rb.i]);
                            rb.next();
                        }
                    }
                }
            }
        }
    } else if (warp < kFirstEpiWarp) {
        // ================= expanders: packed bits -> operand tile (SWIZZLE_128B, K-major) =================
        if constexpr (EXPAND) {
            // thread et expands panel rows et and et + BN / 2 of every k-block (BN / 2 is a multiple of 8, so both
            // rows share the swizzle phase): 16-byte chunk c of a row lands at row * 128 + ((c ^ (row & 7)) << 4)
            const int et = (warp - kFirstExpWarp) * 32 + lane;
            constexpr bool kTwoRows = C::kBRows > kExpThreads;     // two rows per thread, or one (CTA pair: 120 rows)
            constexpr int kHalfRows = kTwoRows ? C::kBRows / 2 : C::kBRows;
            const bool act = et < kHalfRows;
            const int sw = et & 7;
            uint32_t off[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) off[c] = (uint32_t)(et * kRowBytes + ((c ^ sw) << 4));
            const uint32_t b_base = smem_u32(b_tiles);
            const uint32_t raw_base = smem_u32(raws) + (uint32_t)et * C::kRawRow;
            Ring<kRawStages> rr;
            Ring<kBStages> rb;
            int pending = -1;  // B slot whose stores still need the proxy fence + arrive (deferred by one k-block)
            const int tail_words = p.words - WPK * (KB - 1);  // packed words of the last k-block that an MMA reads
            const int tail_chunks = FP4 ? ((tail_words + 1) & ~1) : 2 * tail_words;  // 16-byte chunks to write there
            // list epilogue: the 32 operand bytes behind the last packed words hold the row's column-index code block:
            // 64 E2M1 values that sum to c = the row's position in the tile (c / 6 sixes, then the remainder as 0, 1, 2,
            // 3, 4 or 4 + 1); the query side holds 1.0 there and the MMA's A scale is 2^-8, so the accumulator gets c / 256
            [[maybe_unused]] uint4 ic0[2], ic1[2];      // chunk pair of row et (and et + kHalfRows)
            [[maybe_unused]] uint32_t ioff0 = 0u, ioff1 = 0u;
            if constexpr (C::kListEpi) {
                const int ich = 2 * p.idx_slot - 8 * (KB - 1);  // first of the two chunks, inside the last k-block
                ioff0 = (uint32_t)(et * kRowBytes + ((ich ^ sw) << 4));
                ioff1 = (uint32_t)(et * kRowBytes + (((ich + 1) ^ sw) << 4));
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    uint32_t wds[8];
                    index_code_words((int)cta_rank * C::kBRows + et + r * kHalfRows, wds);
                    ic0[r] = make_uint4(wds[0], wds[1], wds[2], wds[3]);
                    ic1[r] = make_uint4(wds[4], wds[5], wds[6], wds[7]);
                }
            }
            auto lds128 = [](uint32_t addr) {
                uint4 v;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
                return v;
            };
            auto sts128 = [](uint32_t addr, const uint4& v) {
#ifdef TC_DEBUG_NO_EXPAND_STS
                if (v.x != 0xdeadbeefu) return;
#endif
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
            };
            // chunks [0, nchunk) of one row; NCHUNK = 8 is the branch-free hot path
            auto expand_row = [&](const uint4 (&w)[WPK / 4], uint32_t dst, int nchunk) {
#pragma unroll
                for (int v = 0; v < WPK / 4; ++v) {
                    const uint32_t ww[4] = {w[v].x, w[v].y, w[v].z, w[v].w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if constexpr (FP4) {
                            const int c = 4 * v + i;
                            if (c < nchunk) sts128(dst + off[c], expand_panel_word_fp4(ww[i]));
                        } else {
                            const int c = 2 * i;
                            if (c < nchunk) {
                                uint4 c0, c1;
                                expand_panel_word_fp8(ww[i], c0, c1);
                                sts128(dst + off[c], c0);
                                sts128(dst + off[c + 1], c1);
                            }
                        }
                    }
                }
            };
            auto publish = [&](int slot) {  // the slot's stores are fenced: one arrival per warp on the (leader's) full barrier
                if (lane == 0) {
                    if (TWO && !leader) mbar_arrive_cluster(&full_b[slot], 0u);
                    else mbar_arrive(&full_b[slot]);
                }
            };
            if constexpr (TA) {
                // Per tile: 80 rows x nhalves half-k-blocks (4 packed words -> 4 operand chunks) spread over the 128
                // expander threads; unit u = (half h = u / 80, row = u % 80).  The last half holds the tail words, a zero
                // chunk up to the slot boundary and the row's column-index code.
                const int nslots = p.idx_slot + 1;
                const int nhalves = (nslots + 1) >> 1;
                const int units = nhalves * C::kBRows;
                const int ich = 2 * p.idx_slot;          // chunk (= packed word position) of the index block
                const uint32_t rstride = (uint32_t)p.stride * 4u;
                // the rows' column-index codes, once: table [row][32 B] in shared memory (thread = row), visible to the four
                // expander warps after their own named barrier
                const uint32_t idx_tab = smem_u32(const_cast<int32_t*>(qbx)) + 2u * (uint32_t)C::kPartsMax * BM * 4u;
                if (act) {
                    sts128(idx_tab + (uint32_t)et * 32u, ic0[0]);
                    sts128(idx_tab + (uint32_t)et * 32u + 16u, ic1[0]);
                }
                asm volatile("bar.sync 2, %0;" ::"n"(kExpThreads) : "memory");
                for (int item = item0; item < p.items; item += item_step) {
                    const Item it = decode_item(p, item, kQtMul, (int)cta_rank);
                    for (int t = 0; t < it.ntiles; ++t) {
                        mbar_wait(&raw_full[rr.i], rr.phase);
                        mbar_wait(&empty_b[rb.i], rb.phase ^ 1u);
                        const uint32_t src0 = smem_u32(raws) + (uint32_t)rr.i * C::kRawSlot;
                        const uint32_t dst0 = b_base + (uint32_t)rb.i * C::kBSlot;
                        // every raw load of the tile first (up to 7 units per thread), then the expansion: the loads'
                        // latencies overlap instead of adding up
                        constexpr int kMaxUnits = (2 * C::kMaxKbTmemA * C::kBRows + kExpThreads - 1) / kExpThreads;
                        uint4 wv[kMaxUnits];
#pragma unroll
                        for (int i = 0; i < kMaxUnits; ++i) {
                            const int u = et + i * kExpThreads;
                            if (u < units) {
                                const int h = u / C::kBRows, r = u - h * C::kBRows;
                                wv[i] = lds128(src0 + (uint32_t)r * rstride + (uint32_t)h * 16u);
                            }
                        }
#pragma unroll
                        for (int i = 0; i < kMaxUnits; ++i) {
                            const int u = et + i * kExpThreads;
                            if (u < units) {
                                const int h = u / C::kBRows, r = u - h * C::kBRows;
                                const int word0 = 4 * h;
                                const uint32_t dst = dst0 + (uint32_t)(h >> 1) * C::kBSub + (uint32_t)r * kRowBytes;
                                const uint32_t sw2 = (uint32_t)(r & 7), cb = (uint32_t)(h & 1) * 4u;
                                const uint32_t ww[4] = {wv[i].x, wv[i].y, wv[i].z, wv[i].w};
                                if (word0 + 4 <= p.words) {
#pragma unroll
                                    for (int j = 0; j < 4; ++j) sts128(dst + (((cb + j) ^ sw2) << 4), expand_panel_word_fp4(ww[j]));
                                } else {
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        const int wi = word0 + j;
                                        const uint32_t a = dst + (((cb + j) ^ sw2) << 4);
                                        if (wi < ich) sts128(a, expand_panel_word_fp4(wi < p.words ? ww[j] : 0u));
                                        else if (wi == ich) sts128(a, lds128(idx_tab + (uint32_t)r * 32u));
                                        else if (wi == ich + 1) {
                                            if constexpr (C::kDeferFold) {
                                                // tile tag t & 7 as E2M1 codes: 0 1 2 3 4 (4 + 1) 6 (6 + 1)
                                                const uint32_t tagw = (uint32_t)(0x2707260605040200ull >> (8 * (t & 7))) & 0xFFu;
                                                sts128(a, make_uint4(tagw, 0u, 0u, 0u));
                                            } else {
                                                sts128(a, lds128(idx_tab + (uint32_t)r * 32u + 16u));
                                            }
                                        }
                                    }
                                }
                            }
                        }
                        fence_proxy_async();
                        __syncwarp();
                        publish(rb.i);
                        if (lane == 0) mbar_arrive(&raw_empty[rr.i]);
                        rr.next();
                        rb.next();
                    }
                }
            } else
            for (int item = item0; item < p.items; item += item_step) {
                const Item it = decode_item(p, item, kQtMul, (int)cta_rank);
                for (int t = 0; t < it.ntiles; ++t) {
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait(&raw_full[rr.i], rr.phase);
                        uint4 w0[WPK / 4], w1[WPK / 4];
                        if (act) {
                            const uint32_t src = raw_base + (uint32_t)rr.i * C::kRawSlot;
#pragma unroll
                            for (int v = 0; v < WPK / 4; ++v) {
                                w0[v] = lds128(src + v * 16);
                                if constexpr (kTwoRows) w1[v] = lds128(src + kHalfRows * C::kRawRow + v * 16);
                            }
                        }
                        if (pending >= 0) {
                            // the previous k-block's stores have had time to drain: make them visible to the
                            // tensor core (async proxy) and publish the slot
#ifndef TC_DEBUG_NO_FENCE
                            fence_proxy_async();
#endif
                            __syncwarp();
                            publish(pending);
                        }
                        mbar_wait(&empty_b[rb.i], rb.phase ^ 1u);
                        if (act) {
                            const uint32_t dst = b_base + (uint32_t)rb.i * C::kBSlot;
                            if (kb != KB - 1 || (!C::kListEpi && tail_chunks == 8)) {
                                expand_row(w0, dst, 8);
                                if constexpr (kTwoRows) expand_row(w1, dst + kHalfRows * kRowBytes, 8);
                            } else {
                                expand_row(w0, dst, tail_chunks);
                                if constexpr (kTwoRows) expand_row(w1, dst + kHalfRows * kRowBytes, tail_chunks);
                                if constexpr (C::kListEpi) {
                                    sts128(dst + ioff0, ic0[0]);
                                    sts128(dst + ioff1, ic1[0]);
                                    if constexpr (kTwoRows) {
                                        sts128(dst + kHalfRows * kRowBytes + ioff0, ic0[1]);
                                        sts128(dst + kHalfRows * kRowBytes + ioff1, ic1[1]);
                                    }
                                }
                            }
                        }
                        __syncwarp();  // every lane has consumed its raw words
                        if (lane == 0) mbar_arrive(&raw_empty[rr.i]);
#if SNV_TC_DEFER_PUBLISH
                        pending = rb.i;
#else
                        fence_proxy_async();
                        __syncwarp();
                        publish(rb.i);
#endif
                        rr.next();
                        rb.next();
                    }
                }
            }
            if (!TA && pending >= 0) {
                fence_proxy_async();
                __syncwarp();
                publish(pending);
            }
        }
    } else {
        // ================= epilogue: thread = query row (TMEM lane), kParts warps per lane quarter =================
        constexpr int kParts = E::kParts, kPartCols = E::kPartCols, G = E::kGroup;
        static_assert(!TA || kParts == 2 || kParts == 3, "the TMEM-A tile split is written for two (96 + 64) or three (64 + 64 + 32) column parts");
        constexpr uint32_t kSlotStride = E::kSlotStride;
        const int quarter = warp & 3;                   // TMEM lanes [32 * quarter, +32)
        const int part = (warp - kFirstEpiWarp) >> 2;   // columns [kPartCols * part, +kPartCols) of every tile
        const int row = quarter * 32 + lane;
        const int et = (warp - kFirstEpiWarp) * 32 + lane;
        const uint32_t slot_base = smem_u32(lists + et);  // this thread's slot of column j at slot_base + j * kSlotStride
        const int idx_bits = p.idx_bits;
        // a runtime 1 the compiler cannot see through (read back from shared memory): keeps the mask adds on the
        // FMA pipe as IMAD with a register multiplier instead of ALU-pipe LOP3
        const uint32_t one = *reinterpret_cast<volatile uint32_t*>(tmem_ptr + 1);
        uint32_t tcount = 0;
        thrx[part * BM + row] = 3.0e38f;
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        // TMEM-A mode: the epilogue warps build an item's query operand in tensor memory straight from the PACKED query
        // rows (thread = query row = TMEM lane): 4 packed words (+ observed-site mask words) -> 4 operand chunks = 16
        // columns per tcgen05.st; the two column parts take alternate word quads, each adds up its share of the bias
        // popc(q & m) and leaves it in shared memory for the item's start.  One lane per warp then arrives on the
        // leader's a_full barrier.  No expansion kernel, no 640-byte operand rows in HBM.
        [[maybe_unused]] auto load_a = [&](int item_a, int slot) {
            const Item nx = decode_item(p, item_a, kQtMul, (int)cta_rank);
            const int qa = nx.qt * BM + row;
            const bool act = qa < p.nq;
            const int64_t qrow = (int64_t)nx.w * p.nq + (act ? qa : 0);
            const uint4* qsrc = reinterpret_cast<const uint4*>(p.q + qrow * p.stride);
            const uint4* msrc = p.mask ? reinterpret_cast<const uint4*>(p.mask + (int64_t)nx.w * p.mask_win_stride + (int64_t)(act ? qa : 0) * p.mask_q_stride)
                                       : nullptr;
            const uint32_t ta = tmem_base + ((uint32_t)(quarter * 32) << 16) + C::kACol;
            const int nquads = (2 * (p.idx_slot + 1) + 3) >> 2;   // word quads that hold operand chunks (index block included)
            const int ich = 2 * p.idx_slot;
            int32_t bias = 0;
#pragma unroll 1
            for (int qd = part; qd < nquads; qd += kParts) {
                const int w0 = 4 * qd;
                uint4 qv = make_uint4(0u, 0u, 0u, 0u), mv = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
                if (act && w0 < p.stride) {
                    qv = __ldg(qsrc + qd);
                    if (msrc) mv = __ldg(msrc + qd);
                }
                const uint32_t qw[4] = {qv.x, qv.y, qv.z, qv.w}, mw[4] = {mv.x, mv.y, mv.z, mv.w};
                uint32_t v[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int w = w0 + i, lo = 32 * w;
                    // observed-site bits of packed word w: sites < d, restricted by the mask row; 0 past the row end
                    uint32_t m = (w < p.words && act) ? (lo + 32 <= p.d ? 0xFFFFFFFFu : (1u << (p.d - lo)) - 1u) & mw[i] : 0u;
                    const uint32_t wq = qw[i] & m;
                    bias += __popc(wq);
                    uint4 x = expand_query_chunk<true>(wq, m, 0);
                    if (w == ich || w == ich + 1) x = make_uint4(0x22222222u, 0x22222222u, 0x22222222u, 0x22222222u);  // 1.0 x 64: index MMA
                    v[4 * i] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
                }
                tmem_st_32x32b_x16v(ta + (uint32_t)(16 * qd), v);
            }
            qbx[(slot * C::kPartsMax + part) * BM + row] = bias;
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (TWO && !leader) mbar_arrive_cluster(a_full, 0u);
                else mbar_arrive(a_full);
            }
        };
        [[maybe_unused]] uint32_t icount = 0;
        if constexpr (TA) {
            if (item0 < p.items) load_a(item0, 0);
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");  // the partial biases of the first item are visible
        }
        for (int item = item0; item < p.items; item += item_step) {
            const Item it = decode_item(p, item, kQtMul, (int)cta_rank);
            const int qi = it.qt * BM + row;
            const bool active = qi < p.nq;
            const int64_t q = (int64_t)it.w * p.nq + (active ? qi : 0);
            int32_t qb;
            if constexpr (TA) {
                // popc(q & m): the two parts' shares, written when the tile was built (ordered by the item-end barriers)
                const int sl = (int)(icount & 1u);
                qb = 0;
#pragma unroll
                for (int pp = 0; pp < kParts; ++pp) qb += qbx[(sl * C::kPartsMax + pp) * BM + row];
                ++icount;
            } else {
                qb = active ? p.q_bias[q] : 0;
            }
            uint32_t best[KT];
#pragma unroll
            for (int i = 0; i < KT; ++i) best[i] = kSent32;
            // Selection.  Per column: one compare of the raw accumulator against the threshold (strict, so equal
            // distances at later columns never displace: ids ascend along the scan), a predicated store into the
            // column's own slot in shared memory and a predicated bit in a per-group mask - no serial pointer
            // chain.  After each group of G columns the lanes pop their mask bits in lockstep (two per round) and
            // insert into the sorted register top-k.  acc + 1.5 * 2^23 has the integer value of acc in its low
            // mantissa bits:  key = (bits - 0x4B400000 + qb) << idx_bits | column  (mod 2^32), distance = qb + acc.
            // The warps of a query (column parts) publish their thresholds to each other: a candidate must
            // also not exceed the other parts' k-th best (non-strict: ids interleave between the parts).
            const uint32_t kconst = (uint32_t)(qb - 0x4B400000) << idx_bits;
            float thr_mine = 3.0e38f, thr = 3.0e38f;
            auto key_of = [&](float a, uint32_t col) { return (__float_as_uint(a + 12582912.0f) << idx_bits) + kconst + col; };
            float thr_other = 3.0e38f;  // min over the other parts of (published threshold + 1), re-read once per tile
            auto refresh_thr = [&]() {
                // (float)(k-th best distance - qb) without a conversion instruction (|value| < 2^22)
                const int32_t x = (int32_t)(best[KT - 1] >> idx_bits) - qb;
                thr_mine = best[KT - 1] == kSent32 ? 3.0e38f : __uint_as_float(0x4B400000u + (uint32_t)x) - 12582912.0f;
                thr = fminf(thr_mine, thr_other);
            };
            auto fold = [&](uint32_t mask, uint32_t col0) {
                if (__any_sync(0xffffffffu, mask != 0u)) {
                    do {
                        if (mask != 0u) {
                            const int j1 = __ffs((int)mask) - 1;
                            mask &= mask - 1u;
                            const bool two = mask != 0u;
                            const int j2 = two ? __ffs((int)mask) - 1 : j1;
                            mask &= mask - 1u;
                            float a1, a2;
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a1) : "r"(slot_base + (uint32_t)j1 * kSlotStride));
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a2) : "r"(slot_base + (uint32_t)j2 * kSlotStride));
                            const uint32_t key1 = key_of(a1, col0 + (uint32_t)j1);
                            const uint32_t key2 = two ? key_of(a2, col0 + (uint32_t)j2) : kSent32;
                            if (key1 < best[KT - 1]) topk_insert<KT, uint32_t>(best, key1);
                            if (key2 < best[KT - 1]) topk_insert<KT, uint32_t>(best, key2);
                        }
                    } while (__any_sync(0xffffffffu, mask != 0u));
                    refresh_thr();
                }
            };
            // ---- candidate-list epilogue (fp4 CTA-pair engines): every accumulator arrives as  a + c / 256  (a = distance -
            // bias, c = its column in the tile, added by the column-index MMA), so a candidate is self-describing.  The
            // scan takes the columns in pairs: one minimum, one compare, and - predicated - one 8-byte store of the pair
            // to the thread's list plus a count bump; the lists are folded into the sorted top-k once per tile part,
            // after the accumulator stage has gone back to the MMA warp, or earlier when some lane holds kFoldAt
            // entries.  a < thr  <=>  a + c / 256 < thr  for integer a, thr and c < 256; the threshold only ever
            // decreases, so a stale one admits extra candidates but never loses one, and the fold re-checks every
            // member of a pair against the current k-th best key.
            // quad scan: 16-byte entries (four adjacent columns), half as many of them in the same shared memory
            constexpr bool kQuad = TA && C::kListEpi && !C::kEpi12 && (KT <= 8 ? SNV_TC_QUAD_SCAN : SNV_TC_QUAD_SCAN_K32);
            constexpr uint32_t kEntry = kQuad ? 16u : 8u;
            constexpr uint32_t kFoldMark = kQuad ? (uint32_t)C::kFoldAt / 2u : (uint32_t)C::kFoldAt;
            [[maybe_unused]] constexpr uint32_t kLStride = (uint32_t)kEpiThreads * kEntry;  // bytes between a thread's entries
            [[maybe_unused]] const uint32_t list_base = smem_u32(lists) + (uint32_t)et * kEntry;
            [[maybe_unused]] uint32_t lcnt = 0u;
            // (deferred folds: the argument is the current TILE index of the item, else the tile's first column)
            [[maybe_unused]] auto fold_list = [&](uint32_t tile_col0) {
                const uint32_t mx = __reduce_max_sync(0xffffffffu, lcnt);
                if (mx == 0u) return;
                auto key_of_v = [&](float v) {
                    if constexpr (C::kDeferFold) {
                        // 2048 a + 8 c + tag as an integer (|.| < 2^22): the low mantissa bits of v * 2048 + 1.5 * 2^23.  The
                        // entry's tile is the last one at or before the current tile with that tag.
                        const int32_t iv = (int32_t)__float_as_uint(fmaf(v, 2048.0f, 12582912.0f)) - 0x4B400000;
                        const uint32_t tile = tile_col0 - ((tile_col0 - (uint32_t)iv) & 7u);
                        const uint32_t key = ((uint32_t)((iv >> 11) + qb) << idx_bits) + tile * (uint32_t)BN + ((uint32_t)(iv >> 3) & 255u);
                        return v < 3.0e38f ? key : kSent32;
                    }
                    // 256 a + c as an integer (|.| < 2^19): the low mantissa bits of v * 256 + 1.5 * 2^23; +inf (a column
                    // past the panel end) gives a key above every real one
                    const int32_t iv = (int32_t)__float_as_uint(fmaf(v, 256.0f, 12582912.0f)) - 0x4B400000;
                    const uint32_t key = ((uint32_t)((iv >> 8) + qb) << idx_bits) + tile_col0 + (uint32_t)(iv & 255);
                    return v < 3.0e38f ? key : kSent32;
                };
                // Two entries per round, loads predicated.  Of an entry's two values the smaller one is what qualified the
                // pair: it is inserted unguarded at k <= 8 (inserting a key that is not below the k-th best, or the empty
                // sentinel, leaves the list unchanged: straight-line code); the larger one is a candidate only if it is
                // below the threshold the scan used as well, which is rare, so its insert sits behind a branch the warp
                // seldom takes.  Values order like their keys: v = a + c / 256.
                if constexpr (kQuad) {
                    // one entry (four values) per round: the smallest is what qualified the quad and goes in unguarded
                    // (k <= 8) or behind the usual test; each of the others is a candidate only if it is below the scan's
                    // threshold as well (and is not the smallest again), which is rare
#pragma unroll 1
                    for (uint32_t i = 0; i < mx; ++i) {
                        float v0 = 3.0e38f, v1 = 3.0e38f, v2 = 3.0e38f, v3 = 3.0e38f;
                        if (i < lcnt) asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v0), "=f"(v1), "=f"(v2), "=f"(v3) : "r"(list_base + i * kLStride));
                        float lo;
                        asm("min.f32 %0, %1, %2, %3;" : "=f"(lo) : "f"(v0), "f"(v1), "f"(v2));
                        lo = fminf(lo, v3);
                        const uint32_t klo = key_of_v(lo);
                        if constexpr (KT <= 8) topk_insert<KT, uint32_t>(best, klo);
                        else if (klo < best[KT - 1]) topk_insert<KT, uint32_t>(best, klo);
                        const float vv[4] = {v0, v1, v2, v3};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (vv[e] < thr && vv[e] > lo) {
                                const uint32_t kh = key_of_v(vv[e]);
                                if (kh < best[KT - 1]) topk_insert<KT, uint32_t>(best, kh);
                            }
                        }
                    }
                } else {
#pragma unroll 1
                for (uint32_t i = 0; i < mx; i += 2) {
                    float v[4] = {3.0e38f, 3.0e38f, 3.0e38f, 3.0e38f};
                    if (i < lcnt) asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "r"(list_base + i * kLStride));
                    if (i + 1u < lcnt) asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v[2]), "=f"(v[3]) : "r"(list_base + (i + 1u) * kLStride));
#pragma unroll
                    for (int e = 0; e < 4; e += 2) {
                        const float lo = fminf(v[e], v[e + 1]), hi = fmaxf(v[e], v[e + 1]);
                        const uint32_t klo = key_of_v(lo);
                        if constexpr (KT <= 8) topk_insert<KT, uint32_t>(best, klo);
                        else if (klo < best[KT - 1]) topk_insert<KT, uint32_t>(best, klo);
                        if (hi < thr) {
                            const uint32_t khi = key_of_v(hi);
                            if (khi < best[KT - 1]) topk_insert<KT, uint32_t>(best, khi);
                        }
                    }
                }
                }
                lcnt = 0u;
                refresh_thr();
            };
            if constexpr (C::kListEpi) {
                for (int t = 0; t < it.ntiles; ++t, ++tcount) {
                    const uint32_t as = tcount & 1u;
                    const int n0 = (it.t0 + t) * BN;
                    mbar_wait(&tmem_full[as], (tcount >> 1) & 1u);
                    tcgen05_fence_after();
                    if constexpr (TA) {
                        if (t == it.ntiles - 1 && item + item_step < p.items) load_a(item + item_step, (int)(icount & 1u));
                    }
                    // columns [pstart, pstart + pwidth) of the tile belong to this part.  160-column tiles: two parts take
                    // 96 + 64 columns, three parts 64 + 64 + 32, the assignment rotating from tile to tile so that every part
                    // scores whole 32-column groups and carries the same load over two (three) tiles
                    int pstart, pwidth;
                    if constexpr (!TA) {
                        pstart = kPartCols * part;
                        pwidth = kPartCols;
                    } else if constexpr (kParts == 3) {
                        const int slot3 = (part + (int)(tcount % 3u)) % 3;
                        pstart = 64 * slot3;
                        pwidth = slot3 == 2 ? BN - 128 : 64;
                    } else {
                        const int c0 = (tcount & 1u) ? 64 : 96;
                        pstart = part ? c0 : 0;
                        pwidth = part ? BN - c0 : c0;
                    }
                    int cols = (p.n - n0 < BN ? (int)(p.n - n0) : BN) - pstart;  // columns of this part in use
                    cols = cols < 0 ? 0 : (cols > pwidth ? pwidth : cols);
                    const int nch = (cols + G - 1) / G;
                    const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * BN + (uint32_t)pstart;
                    const uint32_t tile_col0 = C::kDeferFold ? (uint32_t)t : (uint32_t)(t * BN);  // columns count from the split's first row
                    thrx[part * BM + row] = thr_mine;
                    thr_other = 3.0e38f;
#pragma unroll
                    for (int o = 1; o < kParts; ++o) thr_other = fminf(thr_other, thrx[((part + o) % kParts) * BM + row] + 1.0f);
                    thr = fminf(thr_mine, thr_other);
                    uint32_t accA[G], accB[G];
                    auto scan = [&](uint32_t (&acc)[G], int u) {
                        if (G * u + G > cols) {
#pragma unroll
                            for (int j = 0; j < G; ++j)
                                if (G * u + j >= cols) acc[j] = 0x7F800000u;  // past the panel end (+inf): never a candidate
                        }
#ifdef TC_DEBUG_NO_EPI
                        return;
#endif
                        if constexpr (kQuad) {
                            static_for<G / 4>([&](auto jc) {
                                constexpr int j = 4 * decltype(jc)::value;
                                const float a0 = __uint_as_float(acc[j]), a1 = __uint_as_float(acc[j + 1]);
                                const float a2 = __uint_as_float(acc[j + 2]), a3 = __uint_as_float(acc[j + 3]);
                                float m;
                                asm("min.f32 %0, %1, %2, %3;" : "=f"(m) : "f"(a0), "f"(a1), "f"(a2));
                                if (fminf(m, a3) < thr) {
                                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(list_base + lcnt * kLStride), "f"(a0), "f"(a1), "f"(a2), "f"(a3) : "memory");
                                    lcnt += 1u;
                                }
                            });
                        } else {
                        static_for<G / 2>([&](auto jc) {
                            constexpr int j = 2 * decltype(jc)::value;
                            const float a0 = __uint_as_float(acc[j]), a1 = __uint_as_float(acc[j + 1]);
                            if (fminf(a0, a1) < thr) {
                                asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(list_base + lcnt * kLStride), "f"(a0), "f"(a1) : "memory");
                                lcnt += 1u;
                            }
                        });
                        }
                        // keep every list behind the fold mark, so that the next group always fits
                        if (u + 1 < nch && __any_sync(0xffffffffu, lcnt >= kFoldMark)) fold_list(tile_col0);
                    };
                    if (nch > 0) tmem_ld_cols<G>(tbase, accA);
                    if (nch > 0) {
                        tmem_ld_wait_cols<G>(accA);
                        if (nch > 1) tmem_ld_cols<G>(tbase + (uint32_t)G, accB);
                        scan(accA, 0);
                    }
                    if (nch > 1) {
                        tmem_ld_wait_cols<G>(accB);
                        if (nch > 2) tmem_ld_cols<G>(tbase + (uint32_t)(2 * G), accA);
                        scan(accB, 1);
                    }
                    if (nch > 2) {
                        tmem_ld_wait_cols<G>(accA);
                        if (nch > 3) tmem_ld_cols<G>(tbase + (uint32_t)(3 * G), accB);
                        scan(accA, 2);
                    }
                    if (nch > 3) {
                        tmem_ld_wait_cols<G>(accB);
                        scan(accB, 3);
                    }
                    // the accumulator stage goes back to the MMA warp before the fold (which works from shared memory)
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (TWO && !leader) mbar_arrive_cluster(&tmem_empty[as], 0u);
                        else mbar_arrive(&tmem_empty[as]);
                    }
#ifndef TC_DEBUG_NO_FOLD
                    if constexpr (C::kDeferFold) {
                        // one by one while the item is young, then once per 8 tiles (before a tag repeats), at the item's end,
                        // and whenever some lane could not take another 32-column group
                        constexpr int kWarm = KT <= 8 ? SNV_TC_FOLD_WARM_K8 : SNV_TC_FOLD_WARM;
                        constexpr int kWin = KT <= 8 ? SNV_TC_FOLD_WINDOW_K8 : SNV_TC_FOLD_WINDOW;
                        static_assert(kWin == 1 || kWin == 2 || kWin == 4 || kWin == 8, "fold window");
                        if (t < kWarm || (t & (kWin - 1)) == kWin - 1 || t == it.ntiles - 1 ||
                            __any_sync(0xffffffffu, lcnt >= kFoldMark))
                            fold_list(tile_col0);
                    } else {
                        fold_list(tile_col0);
                    }
#else
                    lcnt = 0u;
#endif
                }
            } else
            for (int t = 0; t < it.ntiles; ++t, ++tcount) {
                const uint32_t as = tcount & 1u;
                const int n0 = (it.t0 + t) * BN;
                mbar_wait(&tmem_full[as], (tcount >> 1) & 1u);
                tcgen05_fence_after();
                if constexpr (TA) {
                    // the item's last accumulator is complete, so every MMA that reads its query tile has retired:
                    // tensor memory can take the next item's tile while this tile is still being scored
                    if (t == it.ntiles - 1 && item + item_step < p.items) load_a(item + item_step, (int)(icount & 1u));
                }
                // columns [pstart, pstart + pwidth) of the tile belong to this part.  TMEM-A mode (160-column tiles):
                // 96 + 64 columns, the wide side alternating from tile to tile so that both parts score whole
                // 32-column groups and carry the same load over two tiles
                const int c0 = TA ? ((tcount & 1u) ? 64 : 96) : kPartCols;
                const int pstart = TA ? (part ? c0 : 0) : kPartCols * part;
                const int pwidth = TA ? (part ? BN - c0 : c0) : kPartCols;
                int cols = (p.n - n0 < BN ? (int)(p.n - n0) : BN) - pstart;  // columns of this part in use
                cols = cols < 0 ? 0 : (cols > pwidth ? pwidth : cols);
                const int nch = (cols + G - 1) / G;
                const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * BN + (uint32_t)pstart;
                const uint32_t col_base = (uint32_t)(t * BN + pstart);  // columns count from the split's first row
                // exchange thresholds with the other column parts of this query (once per tile; stale is safe)
                thrx[part * BM + row] = thr_mine;
                thr_other = 3.0e38f;
#pragma unroll
                for (int o = 1; o < kParts; ++o) thr_other = fminf(thr_other, thrx[((part + o) % kParts) * BM + row] + 1.0f);
                thr = fminf(thr_mine, thr_other);
                uint32_t accA[G], accB[G];
                auto process = [&](uint32_t (&acc)[G], int u) {
                    if (G * u + G > cols) {
#pragma unroll
                        for (int j = 0; j < G; ++j)
                            if (G * u + j >= cols) acc[j] = 0x7F800000u;  // past the panel end (+inf): never a candidate
                    }
                    const uint32_t col0 = col_base + (uint32_t)(G * u);
#ifdef TC_DEBUG_NO_EPI
                    return;
#endif
                    if (t == 0 && u == 0) {
                        // first group of an item: no threshold yet, every column is a candidate - insert in lockstep
#pragma unroll
                        for (int j = 0; j < G; ++j) {
                            const float a = __uint_as_float(acc[j]);
                            const uint32_t key = a < 3.0e38f ? key_of(a, col0 + (uint32_t)j) : kSent32;
                            if (key < best[KT - 1]) topk_insert<KT, uint32_t>(best, key);
                        }
                        refresh_thr();
                        return;
                    }
                    // four independent partial masks (short dependency chains); the bit is added with a
                    // multiply-add by a runtime 1 so that it issues on the FMA pipe, not the busier ALU pipe
                    uint32_t m4[4] = {0u, 0u, 0u, 0u};
                    static_for<G>([&](auto jc) {
                        constexpr int j = decltype(jc)::value;
                        const float a = __uint_as_float(acc[j]);
                        if (a < thr) {
                            asm volatile("st.shared.f32 [%0], %1;" ::"r"(slot_base + (uint32_t)j * kSlotStride), "f"(a) : "memory");
                            asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(m4[j & 3]) : "r"(one), "n"(1u << j));
                        }
                    });
#ifdef TC_DEBUG_NO_FOLD
                    if (((m4[0] | m4[1]) | (m4[2] | m4[3])) == 0xdeadbeefu) fold(1u, col0);
#else
                    fold((m4[0] | m4[1]) | (m4[2] | m4[3]), col0);
#endif
                };
#ifdef TC_DEBUG_NO_EPI_LD
                if (tbase == 0xffffffffu)
#endif
                if (nch > 0) tmem_ld_cols<G>(tbase, accA);
#ifdef TC_DEBUG_NO_EPI_LD
                if (tbase == 0xffffffffu) {
#endif
                if (nch > 0) {
                    tmem_ld_wait_cols<G>(accA);
                    if (nch > 1) tmem_ld_cols<G>(tbase + (uint32_t)G, accB);
                    process(accA, 0);
                }
                if (nch > 1) {
                    tmem_ld_wait_cols<G>(accB);
                    if (nch > 2) tmem_ld_cols<G>(tbase + (uint32_t)(2 * G), accA);
                    process(accB, 1);
                }
                if (nch > 2) {
                    tmem_ld_wait_cols<G>(accA);
                    if (nch > 3) tmem_ld_cols<G>(tbase + (uint32_t)(3 * G), accB);
                    process(accA, 2);
                }
                if (nch > 3) {
                    tmem_ld_wait_cols<G>(accB);
                    process(accB, 3);
                }
#ifdef TC_DEBUG_NO_EPI_LD
                }
#endif
                tcgen05_fence_before();
#if SNV_TC_WARP_ARRIVE
                // one arrival per warp: a remote (peer -> leader) arrival per thread costs the pair ~2 k clocks per tile
                __syncwarp();
                if (lane == 0) {
                    if (TWO && !leader) mbar_arrive_cluster(&tmem_empty[as], 0u);
                    else mbar_arrive(&tmem_empty[as]);
                }
#else
                if (TWO && !leader) mbar_arrive_cluster(&tmem_empty[as], 0u);
                else mbar_arrive(&tmem_empty[as]);
#endif
            }
            // ---- merge the parts of each query through shared memory, then write the result
            thrx[part * BM + row] = 3.0e38f;  // reset for the next item (ordered by the barriers below)
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");  // every group is folded: the buffer is free
            if (part != 0) {
#pragma unroll
                for (int i = 0; i < KT; ++i) xchg[((part - 1) * KT + i) * BM + row] = best[i];
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
            if (part == 0) {
#pragma unroll 1
                for (int i = 0; i < (kParts - 1) * KT; ++i) {
                    const uint32_t key = xchg[i * BM + row];
                    if (key < best[KT - 1]) topk_insert<KT, uint32_t>(best, key);
                }
                if (active) {
                    const uint32_t idx_mask = (1u << idx_bits) - 1u;
                    const int64_t r0 = (int64_t)it.t0 * BN;
                    if (it.out_split == 1) {
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
                        // tail pieces index their partial keys from the first tail row
                        const int64_t qrel = (p.tail_split > 1 && p.nsplit == 1) ? q - p.tail_row0 : q;
                        uint64_t* out = p.partial + (qrel * it.out_split + it.split) * KT;
#pragma unroll
                        for (int i = 0; i < KT; ++i) {
                            const uint32_t key = best[i];
                            out[i] = key == kSent32 ? kSent64
                                                    : ((uint64_t)(key >> idx_bits) << 32) | (uint64_t)((int64_t)(key & idx_mask) + r0);
                        }
                    }
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");  // exchange reads done before the slots are reused
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if constexpr (TWO) cluster_sync();  // the peer may still read this CTA's operands / arrive on its barriers
    if (warp == 1) {
        tcgen05_fence_after();
        if constexpr (TWO) tmem_dealloc_2cta(tmem_base, kTmemCols);
        else tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---- query operand: packed (q, observed mask) -> operand rows [rows][kblocks * 128 B] + popc(q & m) ----------------
// One warp per query row: lane l owns packed word l (and l + 32, ...) for the bias, and writes the row's 16-byte
// chunks l, l + 32, ... (coalesced); no integer divisions on the hot path.  128-thread blocks at 32 registers fit
// in the 4096 registers the resident scan CTA (480 x 128) leaves free on an SM, so the expansion of the next window
// chunk overlaps the scan of the current one.
template <bool FP4>
__global__ void __launch_bounds__(128)
tc_expand_queries_kernel(const uint32_t* __restrict__ q, const uint32_t* __restrict__ mask, int64_t mask_win_stride,
                         int64_t mask_q_stride, int nq, int64_t rows, int stride, int words, int d, int kblocks, int idx_chunk,
                         uint8_t* __restrict__ ops, int32_t* __restrict__ bias)
{
    const int lane = threadIdx.x & 31;
    const int cpr = kblocks * 8;  // 16-byte chunks per row
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp0; row < rows; row += nwarps) {
        const uint32_t* qr = q + row * stride;
        const uint32_t* mr = mask ? mask + (row / nq) * mask_win_stride + (row % nq) * mask_q_stride : nullptr;
        auto observed = [&](int w) -> uint32_t {  // observed-site bits of packed word w (0 past the row end)
            const int lo = w * 32;
            uint32_t m = lo + 32 <= d ? 0xFFFFFFFFu : (lo < d ? (1u << (d - lo)) - 1u : 0u);
            if (mr && w < words) m &= mr[w];
            return w < words ? m : 0u;
        };
        int32_t b = 0;
        for (int w = lane; w < words; w += 32) b += __popc(qr[w] & observed(w));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
        if (lane == 0) bias[row] = b;
        uint4* out_row = reinterpret_cast<uint4*>(ops + row * (int64_t)cpr * 16);
        for (int ch = lane; ch < cpr; ch += 32) {
            // fp8: chunk pair (2 i, 2 i + 1) of a k-block <- packed word i; fp4: chunk i <- packed word i
            const int wi = FP4 ? ch : (ch >> 3) * 4 + ((ch & 7) >> 1);
            const int h = ch & 1;
            const uint32_t wm = observed(wi);
            const uint32_t wq = wi < words ? qr[wi] & wm : 0u;
            uint4 v = expand_query_chunk<FP4>(wq, wm, h);
            // list epilogue: 1.0 (E2M1 code 2) in the 64 positions of the column-index MMA
            if (FP4 && (ch == idx_chunk || ch == idx_chunk + 1)) v = make_uint4(0x22222222u, 0x22222222u, 0x22222222u, 0x22222222u);
            out_row[ch] = v;
        }
    }
}

// ---- bring-up variant (MODE_FP8_HBM): the panel expanded to fp8 rows in HBM ------------------------------
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
        expand_panel_word_fp8(w, c0, c1);
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

template <int KT, int MODE>
int launch_kernel(const CUtensorMap& map_q, const CUtensorMap& map_r, const TcParams& tp, int grid, cudaStream_t stream)
{
    constexpr size_t smem = Cfg<MODE>::kSmem;
    // the dynamic shared-memory opt-in is per device / context: set on every launch, never cached process-wide
    SNV_CUDA_CHECK(cudaFuncSetAttribute(hamming_tc_kernel<KT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    profile_begin(stream);
    if constexpr (Cfg<MODE>::kTwoCta) {
        // CTA pairs: a cluster of two CTAs on the two SMs of a TPC
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3((unsigned)threads_of<KT, MODE>());
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        SNV_CUDA_CHECK(cudaLaunchKernelEx(&cfg, hamming_tc_kernel<KT, MODE>, map_q, map_r, tp));
    } else {
        hamming_tc_kernel<KT, MODE><<<grid, threads_of<KT, MODE>(), smem, stream>>>(map_q, map_r, tp);
    }
    profile_end(stream);
    SNV_LAUNCH_CHECK();
    return SNV_OK;
}

int mode_of_engine(int engine)
{
    return engine == 5 ? MODE_FP4_2CTA_TA : (engine == 4 ? MODE_FP4_2CTA : (engine == 3 ? MODE_FP4 : (engine == 2 ? MODE_FP8_HBM : MODE_FP8)));
}
bool list_epi_engine(int engine) { return engine >= 4 && Cfg<MODE_FP4_2CTA>::kListEpi; }
int idx_slot_of(int words) { return (words + 1) / 2; }  // MMA slot of the column-index block: right behind the packed words
// k-blocks (128 operand bytes per row) of an engine's operand rows
int kblocks_of_engine(int engine, int words)
{
    if (list_epi_engine(engine)) return (int)ceil_div(idx_slot_of(words) + 1, 4);
    return (int)ceil_div(words, engine >= 3 ? Cfg<MODE_FP4>::WPK : 4);
}
int bn_of_engine(int engine) { return engine == 5 ? Cfg<MODE_FP4_2CTA_TA>::BN : (engine >= 3 ? Cfg<MODE_FP4>::BN : 256); }
int wpk_of_engine(int engine) { return engine >= 3 ? Cfg<MODE_FP4>::WPK : 4; }
bool pair_engine(int engine) { return engine >= 4; }  // items are query-tile PAIRS on SM pairs

}  // namespace

// 0 = popcount kernel, 1 = tensor cores fp8, 2 = fp8 with the panel pre-expanded in HBM (bring-up), 3 = tensor cores fp4,
// 4 = fp4 on CTA pairs (cta_group::2), 5 = 4 with the query operand in tensor memory (bring-up: only when forced with
// SNV_HAMMING_ENGINE=tc4x2ta, windows of at most 1280 sites; wider ones run engine 4)
int hamming_engine_for(const HammingSearchParams& p)
{
    int mode = -1;  // auto
    if (const char* e = getenv("SNV_HAMMING_ENGINE")) {
        if (!strcmp(e, "popc")) mode = 0;
        else if (!strcmp(e, "tc") || !strcmp(e, "tc8")) mode = 1;
        else if (!strcmp(e, "tc_hbm")) mode = 2;
        else if (!strcmp(e, "tc4")) mode = 3;
        else if (!strcmp(e, "tc4x2")) mode = 4;
        else if (!strcmp(e, "tc4x2ta")) mode = 5;
    }
    const bool can = !p.work && p.n > 0 && p.nq > 0 && p.k >= 1 && p.k <= 32 && p.d < (1 << 12) &&
                     (int64_t)p.nw * p.nq < ((int64_t)1 << 31) && p.n < ((int64_t)1 << 31);
    if (!can || mode == 0) return 0;
    if (mode == 5 && (kblocks_of_engine(5, p.words) > Cfg<MODE_FP4_2CTA_TA>::kMaxKbTmemA || p.stride > Cfg<MODE_FP4_2CTA_TA>::kMaxStrideTmemA)) mode = 4;
    if (mode > 0) return mode;
    // auto: enough queries per window to fill a useful part of the 128-lane tile, and a panel worth a tile
    if (!(p.nq >= 32 && p.n >= 512)) return 0;
    // CTA pairs need two query tiles per window to keep both SMs of a pair busy; windows of at most 1216 sites
    // (5 k-blocks of operand incl. the column-index block) keep the query operand in tensor memory (engine 5)
    if (SNV_TC_DEFAULT_ENGINE >= 4 && p.nq <= BM) return 3;
    if (SNV_TC_DEFAULT_ENGINE == 5 && (kblocks_of_engine(5, p.words) > Cfg<MODE_FP4_2CTA_TA>::kMaxKbTmemA || p.stride > Cfg<MODE_FP4_2CTA_TA>::kMaxStrideTmemA))
        return 4;
    return SNV_TC_DEFAULT_ENGINE;
}

size_t hamming_tc_plan(const HammingSearchParams& p, HammingTcPlan& plan)
{
    plan = HammingTcPlan{};
    plan.engine = hamming_engine_for(p);
    if (!plan.engine) return 0;
    const int BN = bn_of_engine(plan.engine);
    plan.kt = p.k <= 8 ? 8 : 32;
    plan.kblocks = kblocks_of_engine(plan.engine, p.words);
    plan.qtiles = (int)ceil_div(p.nq, BM);
    plan.n_tiles = (int)ceil_div(p.n, BN);
    plan.idx_bits = 32 - bit_length64((int64_t)p.d + 1);
    const int64_t max_tiles_per_split = ((int64_t)1 << plan.idx_bits) / BN;
    // Row splits: the persistent CTAs take (window, query tile, row split) items round-robin, so the step lasts
    // ceil(items / SMs) rounds of one item each.  Pick the split count that minimises rounds x (tiles per item +
    // per-item overhead): a few long items leave most SMs idle in the last round, but every item re-learns its
    // top-k threshold from scratch - about kt (1 + ln(rows / kt)) insertions per query, each a lockstep round
    // of the epilogue - and a split adds partial keys for the merge kernel.  The key's id field bounds the tiles
    // per item.
    // CTA-pair engine: an item is a PAIR of query tiles and runs on a pair of SMs
    const bool pair = pair_engine(plan.engine);
    const int64_t units = pair ? kNumSMs / 2 : kNumSMs;
    const int64_t base = (int64_t)p.nw * (pair ? ceil_div(plan.qtiles, 2) : plan.qtiles);
    const int64_t min_split = ceil_div(plan.n_tiles, max_tiles_per_split);
    int64_t nsplit = min_split;
    {
        double best_cost = -1.0;
        const int64_t max_split = std::max<int64_t>(min_split, std::min<int64_t>(32, plan.n_tiles / 2));
        // tiles of time per candidate insertion round, and per partial key merged afterwards (fitted on B200: cfg 5 at
        // 1-8 GPUs, profiles/r1_bench_cfg5_*; k = 32 re-fitted for the round-2 epilogue - with 0.3 the 200,000-row panel
        // was cut into two row splits, 7.16 ms per 8 windows; unsplit items + tail split: 6.51 ms)
        double ins = plan.kt == 8 ? 0.04 : 0.6;
        const double merge_per_key = 3.5e-5;
        if (const char* e = getenv("SNV_TC_INS")) ins = atof(e) > 0 ? atof(e) : ins;  // tuning override
        for (int64_t s = min_split; s <= max_split; ++s) {
            const int64_t per = ceil_div(plan.n_tiles, s);
            const int64_t real = ceil_div(plan.n_tiles, per);
            const int64_t rounds = ceil_div(base * real, units);
            const double rows_item = (double)per * BN;
            const double cand = plan.kt * (1.0 + std::log(std::max(1.0, rows_item / plan.kt)));
            const double cost = (double)rounds * ((double)per + 0.25 + ins * cand) +
                                (real > 1 ? merge_per_key * (double)p.nw * p.nq * real * plan.kt : 0.0);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; nsplit = real; }
        }
    }
    plan.tiles_per_split = (int)ceil_div(plan.n_tiles, nsplit);
    plan.nsplit = (int)ceil_div(plan.n_tiles, plan.tiles_per_split);
    if (base * plan.nsplit > 0x7fffffffLL) {
        set_error("hamming search: grid too large");
        return (size_t)-1;
    }
    // Tail split: with unsplit items, `base mod units` items are left for a last round that keeps only that many
    // SMs (pairs) busy for a whole item.  Cutting just those trailing items into floor(units / rem) row ranges makes
    // the last round short instead (cfg 5: 320 pair items on 74 pairs = 4 rounds + 24 items -> 4.33 rounds, not 5).
    plan.tail_items = 0;
    plan.tail_split = 0;
    plan.tail_tiles = 0;
    const char* no_tail = getenv("SNV_TC_NO_TAIL_SPLIT");
    if (plan.nsplit == 1 && base > units && base % units != 0 && !(no_tail && no_tail[0] == '1')) {
        const int64_t R = base / units, rem = base % units;
        int64_t s2 = std::min<int64_t>(std::min<int64_t>(units / rem, plan.n_tiles / 2), 32);
        if (s2 >= 2) {
            const int64_t per = ceil_div(plan.n_tiles, s2);
            s2 = ceil_div(plan.n_tiles, per);
            const double ins = plan.kt == 8 ? 0.04 : 0.6;
            auto item_cost = [&](double tiles) { return tiles + 0.25 + ins * plan.kt * (1.0 + std::log(std::max(1.0, tiles * BN / plan.kt))); };
            const double now = (double)(R + 1) * item_cost((double)plan.n_tiles);
            const double then = (double)R * item_cost((double)plan.n_tiles) + item_cost((double)per) +
                                3.5e-5 * (double)rem * (pair ? 2 * BM : BM) * s2 * plan.kt;
            if (s2 >= 2 && then < 0.97 * now) {
                plan.tail_items = (int)rem;
                plan.tail_split = (int)s2;
                plan.tail_tiles = (int)per;
            }
        }
    }
    const int64_t rows = (int64_t)p.nw * p.nq;
    // (engine 5 builds its query operand in the scan kernel: no operand rows, no biases in the workspace)
    const bool ws_ops = mode_of_engine(plan.engine) != MODE_FP4_2CTA_TA;
    plan.off_bias = ws_ops ? round_up(rows * plan.kblocks * kRowBytes, 256) : 0;
    plan.off_partial = plan.off_bias + (ws_ops ? round_up(rows * 4, 256) : 0);
    plan.off_panel = plan.off_partial + (plan.nsplit > 1 ? round_up(rows * plan.nsplit * plan.kt * 8, 256)
                                                          : round_up((int64_t)plan.tail_items * (pair ? 2 * BM : BM) * plan.tail_split * plan.kt * 8, 256));
    size_t total = plan.off_panel;
    if (plan.engine == 2) total += (size_t)p.nw * p.n * plan.kblocks * kRowBytes;
    return total + 1024;
}

// The item bookkeeping of a launch (everything decode_item reads); returns the query rows covered by tail items.
static int64_t fill_item_fields(const HammingSearchParams& p, const HammingTcPlan& plan, TcParams& tp)
{
    const bool pair = pair_engine(plan.engine);
    const int64_t rows = (int64_t)p.nw * p.nq;
    tp.nw = p.nw; tp.nq = p.nq; tp.qtiles = pair ? (int)ceil_div(plan.qtiles, 2) : plan.qtiles;
    tp.n = p.n;
    tp.words = p.words; tp.kblocks = plan.kblocks;
    tp.n_tiles = plan.n_tiles; tp.nsplit = plan.nsplit; tp.tiles_per_split = plan.tiles_per_split;
    tp.items = p.nw * tp.qtiles * plan.nsplit;
    tp.items_a = tp.items;
    tp.tail_split = 0;
    tp.tail_tiles = 0;
    tp.tail_row0 = 0;
    int64_t tail_rows = 0;
    if (plan.tail_split > 1) {
        tp.items_a = tp.items - plan.tail_items;
        tp.tail_split = plan.tail_split;
        tp.tail_tiles = plan.tail_tiles;
        tp.items = tp.items_a + plan.tail_items * plan.tail_split;
        const int64_t w_a = tp.items_a / tp.qtiles, qt_a = tp.items_a % tp.qtiles;
        tp.tail_row0 = w_a * p.nq + qt_a * (pair ? 2 * BM : BM);
        tail_rows = rows - tp.tail_row0;
    }
    return tail_rows;
}

// Host-side run of the operand code functions the kernels use (tests/test_tc_codes.py decodes the narrow floats and
// checks that the contraction is the Hamming distance).  One packed word each of query, observed mask and panel row:
// fp4: q_out / r_out 16 bytes (one chunk = 32 sites); fp8: 32 bytes (chunk pair).
void hamming_tc_debug_codes(int fp4, uint32_t q, uint32_t m, uint32_t r, uint8_t* q_out, uint8_t* r_out)
{
    const uint32_t wq = q & m;
    if (fp4) {
        const uint4 a = expand_query_chunk<true>(wq, m, 0);
        const uint4 b = expand_panel_word_fp4(r);
        memcpy(q_out, &a, 16);
        memcpy(r_out, &b, 16);
    } else {
        const uint4 a0 = expand_query_chunk<false>(wq, m, 0), a1 = expand_query_chunk<false>(wq, m, 1);
        uint4 b0, b1;
        expand_panel_word_fp8(r, b0, b1);
        memcpy(q_out, &a0, 16);
        memcpy(q_out + 16, &a1, 16);
        memcpy(r_out, &b0, 16);
        memcpy(r_out + 16, &b1, 16);
    }
}

// Host-side enumeration of the work items a launch with this plan would run, through the same decode_item the
// kernel uses (tests/test_planner.py checks that they tile every (window, query tile, panel tile) exactly once).
// out [cap][8] = (window, query tile, first panel tile, tiles, piece, pieces, partial-key row base, CTA slot).
int64_t hamming_tc_debug_items(const HammingSearchParams& p, const HammingTcPlan& plan, int64_t* out, int64_t cap)
{
    if (!plan.engine) return 0;
    TcParams tp{};
    fill_item_fields(p, plan, tp);
    const bool pair = pair_engine(plan.engine);
    const int ctas = pair ? std::min(tp.items, kNumSMs / 2) : std::min(tp.items, kNumSMs);
    int64_t n = 0;
    for (int item = 0; item < tp.items; ++item)
        for (int r = 0; r < (pair ? 2 : 1); ++r, ++n) {
            if (n >= cap) continue;
            const Item it = decode_item(tp, item, pair ? 2 : 1, r);
            const bool tail = tp.tail_split > 1 && item >= tp.items_a;
            const int64_t v[8] = {it.w, it.qt, it.t0, it.ntiles, it.split, it.out_split, tail ? tp.tail_row0 : 0, item % ctas};
            for (int j = 0; j < 8; ++j) out[n * 8 + j] = v[j];
        }
    return n;
}

int hamming_tc_launch(const HammingSearchParams& p, const HammingTcPlan& plan, void* ws, cudaStream_t stream)
{
    if (p.nw <= 0 || p.nq <= 0) return SNV_OK;
    const int64_t rows = (int64_t)p.nw * p.nq;
    uint8_t* q_ops = static_cast<uint8_t*>(ws);
    int32_t* q_bias = reinterpret_cast<int32_t*>(q_ops + plan.off_bias);
    uint64_t* partial = reinterpret_cast<uint64_t*>(q_ops + plan.off_partial);
    uint8_t* panel_ops = q_ops + plan.off_panel;
    const int kbytes = plan.kblocks * kRowBytes;
    const int BN = bn_of_engine(plan.engine);
    const bool tmem_a = mode_of_engine(plan.engine) == MODE_FP4_2CTA_TA;  // builds its query operand in the scan kernel
    if (!tmem_a) {
        const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(rows, 4), (int64_t)kNumSMs * 64);  // 4 warps = 4 rows per block
        if (plan.engine >= 3)
            tc_expand_queries_kernel<true><<<grid, 128, 0, stream>>>(p.q, p.mask, p.mask_win_stride, p.mask_q_stride, p.nq, rows, p.stride,
                                                                    p.words, p.d, plan.kblocks, list_epi_engine(plan.engine) ? 2 * idx_slot_of(p.words) : -2,
                                                                    q_ops, q_bias);
        else
            tc_expand_queries_kernel<false><<<grid, 128, 0, stream>>>(p.q, p.mask, p.mask_win_stride, p.mask_q_stride, p.nq, rows, p.stride,
                                                                     p.words, p.d, plan.kblocks, -2, q_ops, q_bias);
        SNV_LAUNCH_CHECK();
    }
    CUtensorMap map_q, map_r;
    memset(&map_q, 0, sizeof(map_q));
    if (!tmem_a) {
        const cuuint64_t gdim[2] = {(cuuint64_t)kbytes, (cuuint64_t)rows};
        const cuuint64_t gstride[1] = {(cuuint64_t)kbytes};
        const cuuint32_t box[2] = {kRowBytes, BM};
        int rc = encode_map(&map_q, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, q_ops, gdim, gstride, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    if (plan.engine != 2) {
        // raw packed panel [nw][n][stride] words: box = one k-block of words x one tile of rows of one window
        const cuuint64_t gdim[3] = {(cuuint64_t)p.stride, (cuuint64_t)p.n, (cuuint64_t)p.nw};
        const cuuint64_t gstride[2] = {(cuuint64_t)p.stride * 4, (cuuint64_t)p.panel_win_stride * 4};
        // (TMEM-A mode: whole packed rows, one box per tile)
        const cuuint32_t box[3] = {(cuuint32_t)(tmem_a ? p.stride : wpk_of_engine(plan.engine)), (cuuint32_t)(pair_engine(plan.engine) ? BN / 2 : BN), 1};
        int rc = encode_map(&map_r, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, p.panel, gdim, gstride, box, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc) return rc;
    } else {
        const int64_t total = (int64_t)p.nw * p.n * plan.kblocks * 8;
        const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(total, 256), (int64_t)kNumSMs * 32);
        tc_expand_panel_kernel<<<grid, 256, 0, stream>>>(p.panel, p.panel_win_stride, p.nw, p.n, p.stride, p.words, plan.kblocks, panel_ops);
        SNV_LAUNCH_CHECK();
        const cuuint64_t gdim[3] = {(cuuint64_t)kbytes, (cuuint64_t)p.n, (cuuint64_t)p.nw};
        const cuuint64_t gstride[2] = {(cuuint64_t)kbytes, (cuuint64_t)p.n * kbytes};
        const cuuint32_t box[3] = {kRowBytes, (cuuint32_t)BN, 1};
        int rc = encode_map(&map_r, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, panel_ops, gdim, gstride, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    TcParams tp{};
    const bool pair = pair_engine(plan.engine);
    const int64_t tail_rows = fill_item_fields(p, plan, tp);
    tp.idx_bits = plan.idx_bits; tp.k = p.k; tp.one = 1;
    tp.idx_slot = list_epi_engine(plan.engine) ? idx_slot_of(p.words) : -1;
    tp.id_offset = p.id_offset;
    tp.q_bias = q_bias;
    tp.D_i32 = p.D_i32; tp.D_f32 = p.D_f32; tp.I = p.I;
    tp.partial = partial;
    tp.q_ops = q_ops;
    tp.q = p.q; tp.mask = p.mask; tp.mask_win_stride = p.mask_win_stride; tp.mask_q_stride = p.mask_q_stride;
    tp.stride = p.stride; tp.d = p.d;
    const int grid = pair ? 2 * std::min(tp.items, kNumSMs / 2) : std::min(tp.items, kNumSMs);
    int rc;
    const bool k8 = plan.kt == 8;
    switch (mode_of_engine(plan.engine)) {
        case MODE_FP4_2CTA_TA: rc = k8 ? launch_kernel<8, MODE_FP4_2CTA_TA>(map_q, map_r, tp, grid, stream) : launch_kernel<32, MODE_FP4_2CTA_TA>(map_q, map_r, tp, grid, stream); break;
        case MODE_FP4_2CTA: rc = k8 ? launch_kernel<8, MODE_FP4_2CTA>(map_q, map_r, tp, grid, stream) : launch_kernel<32, MODE_FP4_2CTA>(map_q, map_r, tp, grid, stream); break;
        case MODE_FP4: rc = k8 ? launch_kernel<8, MODE_FP4>(map_q, map_r, tp, grid, stream) : launch_kernel<32, MODE_FP4>(map_q, map_r, tp, grid, stream); break;
        case MODE_FP8_HBM: rc = k8 ? launch_kernel<8, MODE_FP8_HBM>(map_q, map_r, tp, grid, stream) : launch_kernel<32, MODE_FP8_HBM>(map_q, map_r, tp, grid, stream); break;
        default: rc = k8 ? launch_kernel<8, MODE_FP8>(map_q, map_r, tp, grid, stream) : launch_kernel<32, MODE_FP8>(map_q, map_r, tp, grid, stream); break;
    }
    if (rc) return rc;
    if (plan.nsplit > 1)
        return merge_keys_launch(partial, plan.nsplit, plan.kt, rows, p.k, p.id_offset, false, p.D_i32, p.D_f32, p.I, stream);
    if (plan.tail_split > 1 && tail_rows > 0) {
        const int64_t o = tp.tail_row0 * p.k;
        return merge_keys_launch(partial, plan.tail_split, plan.kt, tail_rows, p.k, p.id_offset, false, p.D_i32 ? p.D_i32 + o : nullptr,
                                 p.D_f32 ? p.D_f32 + o : nullptr, p.I + o, stream);
    }
    return SNV_OK;
}

}  // namespace snv
