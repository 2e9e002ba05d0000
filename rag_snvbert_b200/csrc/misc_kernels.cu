// Pack, top-k merge and gather kernels (sm_100a).  All HBM-bound streaming work: coalesced
// warp-wide accesses, no shared-memory reuse to exploit.
#include "common.cuh"
#include "kernels.cuh"
#include "topk.cuh"

namespace snv {

namespace {

// ------------------------------------------------------------------------------ merge
__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v)
{
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const uint64_t o = __shfl_xor_sync(0xffffffffu, v, s);
        v = o < v ? o : v;
    }
    return v;
}

// Warp-wide sorting network on one 64-bit key per lane (ascending over the lanes).  The key merges below are one warp
// per row: 32 keys load as one coalesced 256-byte line, sort in 15 shuffle stages, and fold into the running best 32
// with the 6-stage bitonic merge - about 200 instructions per 32 keys instead of a serial register insertion per key
// plus k rounds of warp-minimum + pop (round 1: 1 ms of the 8.2 ms cfg 5 step at 1 GPU went into that).
__device__ __forceinline__ uint64_t warp_cmpx_u64(uint64_t v, int s, bool take_min)
{
    const uint64_t o = __shfl_xor_sync(0xffffffffu, v, s);
    const bool lt = o < v;
    return (lt == take_min) ? o : v;
}

__device__ __forceinline__ uint64_t warp_sort32_u64(uint64_t v, int lane)
{
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int s = size >> 1; s > 0; s >>= 1) v = warp_cmpx_u64(v, s, ((lane & size) == 0) == ((lane & s) == 0));
    }
    return v;
}

// cur, other: ascending over the lanes -> the 32 smallest of their union, ascending
__device__ __forceinline__ uint64_t warp_merge32_u64(uint64_t cur, uint64_t other, int lane)
{
    const uint64_t rev = __shfl_sync(0xffffffffu, other, 31 - lane);
    uint64_t v = rev < cur ? rev : cur;  // bitonic: holds the 32 smallest
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v = warp_cmpx_u64(v, s, (lane & s) == 0);
    return v;
}

// One WARP per query: the best k of parts * kin candidate keys (any order; empty = kSent64), ascending.  Keys are
// unique (they embed the row id) except the empty sentinel.  KT only names the instantiation (k <= 32 either way).
template <int KT>
__global__ void __launch_bounds__(256)
merge_keys_kernel(const uint64_t* __restrict__ keys, int parts, int kin, int64_t nq_total, int k,
                  int64_t id_offset, bool float_dist, int32_t* __restrict__ D_i32,
                  float* __restrict__ D_f32, int64_t* __restrict__ I)
{
    const int lane = threadIdx.x & 31;
    const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= nq_total) return;  // warp-uniform
    const uint64_t* src = keys + q * parts * kin;
    const int total = parts * kin;
    uint64_t cur = kSent64;
    for (int j0 = 0; j0 < total; j0 += 32) {
        const uint64_t v = warp_sort32_u64(j0 + lane < total ? src[j0 + lane] : kSent64, lane);
        cur = j0 ? warp_merge32_u64(cur, v, lane) : v;
    }
    if (lane < k) {
        const bool empty = cur == kSent64;
        const uint32_t hi = (uint32_t)(cur >> 32);
        const int64_t o = q * k + lane;
        I[o] = empty ? -1 : (int64_t)(uint32_t)cur + id_offset;
        if (float_dist) {
            if (D_f32) D_f32[o] = empty ? 3.4028234663852886e38f : __uint_as_float(hi);
        } else {
            if (D_i32) D_i32[o] = empty ? 0x7FFFFFFF : (int32_t)hi;
            if (D_f32) D_f32[o] = empty ? 3.4028234663852886e38f : (float)hi;
        }
    }
}

// (D, I) [parts][nq][kin] with global 63-bit ids -> best kout by (D, I).  Ids do not fit the
// 32-bit key half in general, so the comparison is on the (dist, id) pair explicitly.
template <int KT, typename DT>
__global__ void __launch_bounds__(128)
merge_results_kernel(const DT* __restrict__ D, const int64_t* __restrict__ I, int parts, int64_t nq,
                     int kin, int kout, DT pad, DT* __restrict__ Do, int64_t* __restrict__ Io)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    DT bd[KT];
    int64_t bi[KT];
    const int64_t kIdMax = 0x7FFFFFFFFFFFFFFFLL;
#pragma unroll
    for (int i = 0; i < KT; ++i) { bd[i] = pad; bi[i] = kIdMax; }
    for (int part = 0; part < parts; ++part) {
        const int64_t base = ((int64_t)part * nq + q) * kin;
        for (int j = 0; j < kin; ++j) {
            const int64_t id = I[base + j];
            if (id < 0) continue;
            const DT dv = D[base + j];
            if (dv < bd[KT - 1] || (dv == bd[KT - 1] && id < bi[KT - 1])) {
                // sorted insertion, branch-free over the register array
                DT cd = dv;
                int64_t ci = id;
#pragma unroll
                for (int i = 0; i < KT; ++i) {
                    const bool lt = cd < bd[i] || (cd == bd[i] && ci < bi[i]);
                    const DT td = bd[i];
                    const int64_t ti = bi[i];
                    bd[i] = lt ? cd : td;
                    bi[i] = lt ? ci : ti;
                    cd = lt ? td : cd;
                    ci = lt ? ti : ci;
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < KT; ++i) {
        if (i < kout) {
            const bool empty = bi[i] == kIdMax;
            Do[q * kout + i] = empty ? pad : bd[i];
            Io[q * kout + i] = empty ? -1 : bi[i];
        }
    }
}

// Warp-cooperative variant (parts * kin <= 32 * SHARE candidates per query): each lane keeps a sorted share of the
// candidates in registers, then kout rounds of a lexicographic warp minimum over the lane heads (three 32-bit
// reductions: distance, id high word, id low word) pop the winners in (D, I) order.  One warp per query instead
// of one thread per query: coalescing-friendly and ~50x fewer serial steps for k = 32 over 8 shards.
template <typename DT>
__device__ __forceinline__ uint32_t dist_key(DT d);
template <>
__device__ __forceinline__ uint32_t dist_key<int32_t>(int32_t d) { return (uint32_t)d ^ 0x80000000u; }
template <>
__device__ __forceinline__ uint32_t dist_key<float>(float d)
{
    const uint32_t u = __float_as_uint(d);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
template <typename DT>
__device__ __forceinline__ DT dist_unkey(uint32_t k);
template <>
__device__ __forceinline__ int32_t dist_unkey<int32_t>(uint32_t k) { return (int32_t)(k ^ 0x80000000u); }
template <>
__device__ __forceinline__ float dist_unkey<float>(uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

template <int SHARE, typename DT>
__global__ void __launch_bounds__(256)
merge_results_warp_kernel(const DT* __restrict__ D, const int64_t* __restrict__ I, int parts, int64_t nq, int kin,
                          int kout, DT pad, DT* __restrict__ Do, int64_t* __restrict__ Io)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int total = parts * kin;
    const int64_t kIdMax = 0x7FFFFFFFFFFFFFFFLL;
    for (int64_t q = warp0; q < nq; q += nwarps) {
        uint32_t kd[SHARE];
        int64_t ki[SHARE];
#pragma unroll
        for (int i = 0; i < SHARE; ++i) { kd[i] = 0xFFFFFFFFu; ki[i] = kIdMax; }
#pragma unroll
        for (int s = 0; s < SHARE; ++s) {
            const int c = lane + 32 * s;
            if (c < total) {
                const int part = c / kin, j = c - part * kin;
                const int64_t o = ((int64_t)part * nq + q) * kin + j;
                int64_t ci = I[o];
                if (ci >= 0) {
                    uint32_t cd = dist_key<DT>(D[o]);
#pragma unroll
                    for (int i = 0; i < SHARE; ++i) {  // sorted insertion, branch-free over the register array
                        const bool lt = cd < kd[i] || (cd == kd[i] && ci < ki[i]);
                        const uint32_t td = kd[i];
                        const int64_t ti = ki[i];
                        kd[i] = lt ? cd : td;
                        ki[i] = lt ? ci : ti;
                        cd = lt ? td : cd;
                        ci = lt ? ti : ci;
                    }
                }
            }
        }
        for (int r = 0; r < kout; ++r) {
            const uint32_t hd = kd[0];
            const uint32_t hh = (uint32_t)((uint64_t)ki[0] >> 32), hl = (uint32_t)ki[0];
            const uint32_t md = __reduce_min_sync(0xffffffffu, hd);
            const uint32_t mh = __reduce_min_sync(0xffffffffu, hd == md ? hh : 0xFFFFFFFFu);
            const uint32_t ml = __reduce_min_sync(0xffffffffu, (hd == md && hh == mh) ? hl : 0xFFFFFFFFu);
            const bool win = hd == md && hh == mh && hl == ml;
            const uint32_t winners = __ballot_sync(0xffffffffu, win);
            const int64_t mi = (int64_t)(((uint64_t)mh << 32) | ml);
            if (lane == 0) {
                const bool empty = mi == kIdMax;
                Do[q * kout + r] = empty ? pad : dist_unkey<DT>(md);
                Io[q * kout + r] = empty ? -1 : mi;
            }
            if (win && lane == __ffs((int)winners) - 1) {
#pragma unroll
                for (int i = 0; i + 1 < SHARE; ++i) { kd[i] = kd[i + 1]; ki[i] = ki[i + 1]; }
                kd[SHARE - 1] = 0xFFFFFFFFu;
                ki[SHARE - 1] = kIdMax;
            }
        }
    }
}

// ------------------------------------------------------------------------------ position intersection
// Observed-site masks of a ref / target position intersection (test_faiss_intersect.py:128-140 keeps only the
// shared sites of a window): thread = one packed mask word of one window; every bit's ref site is
// looked up in the ascending target positions by binary search.  `ploidy` 1: bit s <-> site s (haplotype rows);
// 2: bits 2s, 2s+1 <-> site s (the offline scripts' sample rows s0h0, s0h1, ...).
__global__ void __launch_bounds__(256)
intersect_mask_kernel(const int64_t* __restrict__ ref_pos, int64_t n_ref, const int64_t* __restrict__ tgt_pos, int64_t n_tgt,
                      const int64_t* __restrict__ window_info, int n_windows, int64_t d, int ploidy, int stride,
                      uint32_t* __restrict__ out)
{
    const int64_t total = (int64_t)n_windows * stride;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int w = (int)(i / stride);
        const int word = (int)(i % stride);
        const int64_t start = window_info[2 * w], end = window_info[2 * w + 1];
        uint32_t bits = 0;
        for (int b = 0; b < 32; ++b) {
            const int64_t col = (int64_t)word * 32 + b;
            if (col >= d) break;
            const int64_t site = start + col / ploidy;
            if (site >= end || site >= n_ref || site < 0) continue;
            if (ploidy == 2 && (b & 1)) {  // second haplotype of the same site: copy the bit just computed
                bits |= ((bits >> (b - 1)) & 1u) << b;
                continue;
            }
            const int64_t key = ref_pos[site];
            int64_t lo = 0, hi = n_tgt;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (tgt_pos[mid] < key) lo = mid + 1;
                else hi = mid;
            }
            if (lo < n_tgt && tgt_pos[lo] == key) bits |= 1u << b;
        }
        out[i] = bits;
    }
}

// ------------------------------------------------------------------------------ pack
// One warp per (row, 32-site group): coalesced 32-element read, ballot -> one packed word
// (lane l <-> site 32*w + l <-> bit l: LSB-first).
template <typename T>
__global__ void __launch_bounds__(256)
pack_sites_kernel(const T* __restrict__ x, int64_t rows, int64_t d, int words, int stride, bool invert,
                  uint32_t* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t total = rows * stride;
    for (int64_t item = warp; item < total; item += nwarps) {
        const int64_t r = item / stride;
        const int w = (int)(item % stride);
        uint32_t word = 0;
        if (w < words) {
            const int64_t s = (int64_t)w * 32 + lane;
            bool bit = false;
            if (s < d) {
                bit = x[r * d + s] != (T)0;
                if (invert) bit = !bit;
            }
            word = __ballot_sync(0xffffffffu, bit);
        }
        if (lane == 0) out[item] = word;
    }
}

// tokens: allele plane = (tok == 6), observed plane = (tok == 5 || tok == 6)
__global__ void __launch_bounds__(256)
pack_tokens_kernel(const int64_t* __restrict__ x, int64_t rows, int64_t d, int words, int stride,
                   uint32_t* __restrict__ out, uint32_t* __restrict__ out_obs)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t total = rows * stride;
    for (int64_t item = warp; item < total; item += nwarps) {
        const int64_t r = item / stride;
        const int w = (int)(item % stride);
        uint32_t a = 0, o = 0;
        if (w < words) {
            const int64_t s = (int64_t)w * 32 + lane;
            int64_t tok = 0;
            if (s < d) tok = x[r * d + s];
            a = __ballot_sync(0xffffffffu, tok == 6);
            o = __ballot_sync(0xffffffffu, tok == 5 || tok == 6);
        }
        if (lane == 0) {
            out[item] = a;
            if (out_obs) out_obs[item] = o;
        }
    }
}

// np.packbits bytes -> zero padded word rows (bit order inside a byte is irrelevant to the
// Hamming distance as long as panel and queries agree).
__global__ void __launch_bounds__(256)
pack_bytes_kernel(const uint8_t* __restrict__ x, int64_t rows, int64_t row_bytes, int stride,
                  uint32_t* __restrict__ out)
{
    const int64_t total = rows * stride;
    for (int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; item < total;
         item += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = item / stride;
        const int w = (int)(item % stride);
        uint32_t word = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int64_t byte = (int64_t)w * 4 + b;
            if (byte < row_bytes) word |= (uint32_t)x[r * row_bytes + byte] << (8 * b);
        }
        out[item] = word;
    }
}

// ------------------------------------------------------------------------------ gather
// out[w][q][j][0..seq_len): [SOS] + alleles(5|6) + [EOS] + PAD; one warp per retrieved row,
// lanes stride over the sequence (coalesced int64 stores; 8 KB per row).
__global__ void __launch_bounds__(256)
gather_tokens_kernel(const uint32_t* __restrict__ panel, int64_t panel_win_stride, int stride, int64_t n,
                     const int64_t* __restrict__ I, int64_t id_offset, int64_t rows_total,
                     int64_t rows_per_window, const int32_t* __restrict__ n_sites, int d, int seq_len,
                     int64_t* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t item = warp; item < rows_total; item += nwarps) {
        const int64_t w = item / rows_per_window;
        const int64_t id = I[item] - id_offset;
        int64_t* o = out + item * seq_len;
        if (id < 0 || id >= n) {
            for (int c = lane; c < seq_len; c += 32) o[c] = 0;
            continue;
        }
        const int ns = n_sites ? n_sites[w] : d;
        const uint32_t* row = panel + w * panel_win_stride + id * stride;
        for (int c = lane; c < seq_len; c += 32) {
            int64_t tok;
            if (c == 0) tok = 2;
            else if (c <= ns) {
                const int s = c - 1;
                tok = 5 + ((row[s >> 5] >> (s & 31)) & 1u);
            } else if (c == ns + 1) tok = 3;
            else tok = 0;
            o[c] = tok;
        }
    }
}

__global__ void __launch_bounds__(256)
gather_tokens_grouped_kernel(const uint32_t* __restrict__ panel, int64_t panel_win_stride, int stride, int64_t n,
                             const int64_t* __restrict__ I, const int32_t* __restrict__ meta, int64_t rows_total,
                             int k, int seq_len, int64_t* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t item = warp; item < rows_total; item += nwarps) {
        const int64_t qi = item / k;
        const int w = meta[2 * qi], ns = meta[2 * qi + 1];
        const int64_t id = I[item];
        int64_t* o = out + item * seq_len;
        if (id < 0 || id >= n) {
            for (int c = lane; c < seq_len; c += 32) o[c] = 0;
            continue;
        }
        const uint32_t* row = panel + (int64_t)w * panel_win_stride + id * stride;
        for (int c = lane; c < seq_len; c += 32) {
            int64_t tok;
            if (c == 0) tok = 2;
            else if (c <= ns) {
                const int s = c - 1;
                tok = 5 + ((row[s >> 5] >> (s & 31)) & 1u);
            } else if (c == ns + 1) tok = 3;
            else tok = 0;
            o[c] = tok;
        }
    }
}

__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ panel, int64_t panel_win_stride, int64_t d, int64_t n,
                   const int64_t* __restrict__ I, int64_t rows_total, int64_t rows_per_window,
                   float* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t item = warp; item < rows_total; item += nwarps) {
        const int64_t w = item / rows_per_window;
        const int64_t id = I[item];
        float* o = out + item * d;
        if (id < 0 || id >= n) {
            for (int64_t c = lane; c < d; c += 32) o[c] = 0.f;
        } else {
            const float* row = panel + w * panel_win_stride + id * d;
            for (int64_t c = lane; c < d; c += 32) o[c] = row[c];
        }
    }
}

int grid_for_warps(int64_t warps, int block)
{
    const int64_t wpb = block / 32;
    int64_t g = ceil_div(warps, wpb);
    const int64_t cap = (int64_t)kNumSMs * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace

int merge_keys_launch(const uint64_t* keys, int parts, int kin, int64_t nq_total, int k,
                      int64_t id_offset, bool float_dist, int32_t* D_i32, float* D_f32, int64_t* I,
                      cudaStream_t stream)
{
    if (nq_total <= 0) return SNV_OK;
    const int block = 256;
    const unsigned grid = (unsigned)ceil_div(nq_total, block / 32);
    if (k <= 8)
        merge_keys_kernel<8><<<grid, block, 0, stream>>>(keys, parts, kin, nq_total, k, id_offset, float_dist, D_i32, D_f32, I);
    else if (k <= 32)
        merge_keys_kernel<32><<<grid, block, 0, stream>>>(keys, parts, kin, nq_total, k, id_offset, float_dist, D_i32, D_f32, I);
    else {
        set_error("merge: k > 32 unsupported");
        return SNV_ERR_UNSUPPORTED;
    }
    SNV_LAUNCH_CHECK();
    return SNV_OK;
}

int merge_results_launch(const int32_t* D_i32, const float* D_f32, const int64_t* I, int parts,
                         int64_t nq, int kin, int kout, int32_t* Do_i32, float* Do_f32,
                         int64_t* Io, cudaStream_t stream)
{
    if (nq <= 0) return SNV_OK;
    if (kout > 32 || kout < 1) {
        set_error("merge: k_out must be in [1, 32]");
        return SNV_ERR_UNSUPPORTED;
    }
    if ((int64_t)parts * kin <= 32 * 32) {
        const int blockw = 256;
        const unsigned gridw = (unsigned)std::min<int64_t>(ceil_div(nq, blockw / 32), (int64_t)kNumSMs * 32);
        const bool small = (int64_t)parts * kin <= 32 * 8;
        if (D_i32) {
            if (small) merge_results_warp_kernel<8, int32_t><<<gridw, blockw, 0, stream>>>(D_i32, I, parts, nq, kin, kout, 0x7FFFFFFF, Do_i32, Io);
            else merge_results_warp_kernel<32, int32_t><<<gridw, blockw, 0, stream>>>(D_i32, I, parts, nq, kin, kout, 0x7FFFFFFF, Do_i32, Io);
        } else {
            if (small) merge_results_warp_kernel<8, float><<<gridw, blockw, 0, stream>>>(D_f32, I, parts, nq, kin, kout, 3.4028234663852886e38f, Do_f32, Io);
            else merge_results_warp_kernel<32, float><<<gridw, blockw, 0, stream>>>(D_f32, I, parts, nq, kin, kout, 3.4028234663852886e38f, Do_f32, Io);
        }
        SNV_LAUNCH_CHECK();
        return SNV_OK;
    }
    const int block = 128;
    const unsigned grid = (unsigned)ceil_div(nq, block);
    if (D_i32) {
        if (kout <= 8)
            merge_results_kernel<8, int32_t><<<grid, block, 0, stream>>>(D_i32, I, parts, nq, kin, kout, 0x7FFFFFFF, Do_i32, Io);
        else
            merge_results_kernel<32, int32_t><<<grid, block, 0, stream>>>(D_i32, I, parts, nq, kin, kout, 0x7FFFFFFF, Do_i32, Io);
    } else {
        if (kout <= 8)
            merge_results_kernel<8, float><<<grid, block, 0, stream>>>(D_f32, I, parts, nq, kin, kout, 3.4028234663852886e38f, Do_f32, Io);
        else
            merge_results_kernel<32, float><<<grid, block, 0, stream>>>(D_f32, I, parts, nq, kin, kout, 3.4028234663852886e38f, Do_f32, Io);
    }
    SNV_LAUNCH_CHECK();
    return SNV_OK;
}

// dense packed rows [rows][words] -> strided packed rows [rows][stride] (pad words zero; optional complement over the d sites)
__global__ void __launch_bounds__(256)
restride_words_kernel(const uint32_t* __restrict__ x, int64_t rows, int words, int stride, int64_t d, bool invert,
                      uint32_t* __restrict__ out)
{
    const int64_t total = rows * stride;
    for (int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; item < total; item += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = item / stride;
        const int w = (int)(item % stride);
        uint32_t v = w < words ? x[r * words + w] : 0u;
        if (invert) {
            const int64_t lo = (int64_t)w * 32;
            const uint32_t valid = lo + 32 <= d ? 0xFFFFFFFFu : (lo < d ? (1u << (d - lo)) - 1u : 0u);
            v = ~v & valid;
        }
        out[item] = v;
    }
}

int pack_launch(const void* x, int64_t rows, int64_t d, int dtype, bool invert, int stride,
                uint32_t* out, uint32_t* out_observed, cudaStream_t stream)
{
    if (rows <= 0) return SNV_OK;
    const int words = (int)ceil_div(d, 32);
    const int block = 256;
    const int grid = grid_for_warps(rows * stride, block);
    switch (dtype) {
        case SNV_DT_U8:
            pack_sites_kernel<uint8_t><<<grid, block, 0, stream>>>((const uint8_t*)x, rows, d, words, stride, invert, out);
            break;
        case SNV_DT_F32:
            pack_sites_kernel<float><<<grid, block, 0, stream>>>((const float*)x, rows, d, words, stride, invert, out);
            break;
        case SNV_DT_I64_TOKENS:
            pack_tokens_kernel<<<grid, block, 0, stream>>>((const int64_t*)x, rows, d, words, stride, out, out_observed);
            break;
        case SNV_DT_PACKED_U8: {
            const int g2 = (int)std::min<int64_t>(ceil_div(rows * stride, block), (int64_t)kNumSMs * 16);
            pack_bytes_kernel<<<g2, block, 0, stream>>>((const uint8_t*)x, rows, ceil_div(d, 8), stride, out);
            break;
        }
        case SNV_DT_PACKED_U32_DENSE: {
            // dense packed rows (words uint32 each) -> strided rows: the byte re-pack with 4 * words bytes per row is
            // exactly a little-endian word copy with zero padding; `invert` (masks given as MISSING sites) flips the
            // observed bits afterwards
            const int g2 = (int)std::min<int64_t>(ceil_div(rows * stride, block), (int64_t)kNumSMs * 16);
            restride_words_kernel<<<g2, block, 0, stream>>>((const uint32_t*)x, rows, words, stride, d, invert, out);
            break;
        }
        default:
            set_error("pack: unsupported dtype");
            return SNV_ERR_INVALID;
    }
    SNV_LAUNCH_CHECK();
    return SNV_OK;
}

int gather_tokens_launch(const uint32_t* panel, int64_t panel_win_stride, int stride, int64_t n,
                         const int64_t* I, int64_t id_offset, int nw, int64_t nq, int k,
                         const int32_t* n_sites_dev, int d, int seq_len, int64_t* out,
                         cudaStream_t stream)
{
    const int64_t rows_total = (int64_t)nw * nq * k;
    if (rows_total <= 0) return SNV_OK;
    const int block = 256;
    gather_tokens_kernel<<<grid_for_warps(rows_total, block), block, 0, stream>>>(
        panel, panel_win_stride, stride, n, I, id_offset, rows_total, nq * k, n_sites_dev, d, seq_len, out);
    SNV_LAUNCH_CHECK();
    return SNV_OK;
}

int gather_tokens_grouped_launch(const uint32_t* panel, int64_t panel_win_stride, int stride, int64_t n,
                                 const int64_t* I, const int32_t* meta, int64_t nq, int k, int seq_len,
                                 int64_t* out, cudaStream_t stream)
{
    const int64_t rows_total = nq * k;
    if (rows_total <= 0) return SNV_OK;
    const int block = 256;
    gather_tokens_grouped_kernel<<<grid_for_warps(rows_total, block), block, 0, stream>>>(
        panel, panel_win_stride, stride, n, I, meta, rows_total, k, seq_len, out);
    SNV_LAUNCH_CHECK();
    return SNV_OK;
}

int gather_rows_launch(const float* panel, int64_t panel_win_stride, int64_t d, int64_t n,
                       const int64_t* I, int nw, int64_t nq, int k, float* out, cudaStream_t stream)
{
    const int64_t rows_total = (int64_t)nw * nq * k;
    if (rows_total <= 0) return SNV_OK;
    const int block = 256;
    gather_rows_kernel<<<grid_for_warps(rows_total, block), block, 0, stream>>>(
        panel, panel_win_stride, d, n, I, rows_total, nq * k, out);
    SNV_LAUNCH_CHECK();
    return SNV_OK;
}

int intersect_masks_launch(const int64_t* ref_pos, int64_t n_ref, const int64_t* tgt_pos, int64_t n_tgt,
                           const int64_t* window_info, int n_windows, int64_t d, int ploidy, int stride, uint32_t* out,
                           cudaStream_t stream)
{
    if (n_windows <= 0) return SNV_OK;
    const int block = 256;
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div((int64_t)n_windows * stride, block), (int64_t)kNumSMs * 16);
    intersect_mask_kernel<<<grid, block, 0, stream>>>(ref_pos, n_ref, tgt_pos, n_tgt, window_info, n_windows, d, ploidy, stride, out);
    SNV_LAUNCH_CHECK();
    return SNV_OK;
}

// ---- compact result rows: (int32 distance, int64 id) -> (uint16 distance, int32 id) ------------------------------
// Host-buffer searches move 12 bytes per neighbour back over PCIe; ids below 2^31 and Hamming distances below 65535
// fit 6.  Missing entries (id -1, distance INT32_MAX) become (-1, 0xFFFF).
__global__ void __launch_bounds__(256)
narrow_results_kernel(const int32_t* __restrict__ D, const int64_t* __restrict__ I, int64_t n, uint16_t* __restrict__ D16,
                      int32_t* __restrict__ I32)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t id = I[i];
        const int32_t d = D[i];
        I32[i] = (int32_t)id;
        D16[i] = id < 0 ? (uint16_t)0xFFFF : (uint16_t)d;
    }
}

int narrow_results_launch(const int32_t* D, const int64_t* I, int64_t n, uint16_t* D16, int32_t* I32, cudaStream_t stream)
{
    if (n <= 0) return SNV_OK;
    const int grid = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)kNumSMs * 8);
    narrow_results_kernel<<<grid, 256, 0, stream>>>(D, I, n, D16, I32);
    SNV_LAUNCH_CHECK();
    return SNV_OK;
}

// ---- row-sharded panel: exchange keys ------------------------------------------------------------------------------
// One int64 key per neighbour, distance << 40 | global id (missing -> INT64_MAX), laid out destination-major so that ONE
// all_to_all_single moves every rank's candidates for the queries another rank owns: 8 instead of 12 bytes on the wire,
// one collective instead of two, one pack launch instead of a chain of elementwise kernels.
constexpr int kXchgIdBits = 40;
__global__ void __launch_bounds__(256)
exchange_pack_kernel(const int32_t* __restrict__ D, const int64_t* __restrict__ I, int nw, int64_t nq, int k, int parts,
                     int64_t* __restrict__ keys)
{
    const int64_t qg = nq / parts;
    const int64_t total = (int64_t)nw * nq * k;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(i % k);
        const int64_t row = i / k;
        const int64_t q = row % nq, w = row / nq;
        const int64_t g = q / qg;
        const int64_t id = I[i];
        const int64_t key = id < 0 ? 0x7FFFFFFFFFFFFFFFLL : (((int64_t)D[i] << kXchgIdBits) | (id & ((1LL << kXchgIdBits) - 1)));
        keys[(((g * nw + w) * qg) + (q - g * qg)) * k + j] = key;
    }
}

// keys [parts][n][kin] (what the all-to-all delivered: one list per source rank and row) -> the k_out best per row,
// unpacked to (int32 distance, int64 id).  One warp per row, the same sorting network as merge_keys_kernel: each part's
// list is one coalesced line.  Keys are unique (they embed the id), INT64_MAX = empty.
template <int KT>
__global__ void __launch_bounds__(256)
exchange_merge_kernel(const int64_t* __restrict__ keys, int parts, int64_t n, int kin, int kout, int32_t* __restrict__ Do,
                      int64_t* __restrict__ Io)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;  // warp-uniform
    constexpr uint64_t kEmpty = 0x7FFFFFFFFFFFFFFFull;
    uint64_t cur = kSent64;
    bool first = true;
    // short lists (kin <= 16): one chunk takes 32 / kin parts at once
    const int ppc = kin <= 16 ? 32 / kin : 1;
    const int lp = kin <= 16 ? lane / kin : 0, lj = kin <= 16 ? lane - lp * kin : lane;
    for (int part0 = 0; part0 < parts; part0 += ppc) {
        for (int j0 = 0; j0 < kin; j0 += 32) {
            const int part = part0 + lp, j = j0 + lj;
            uint64_t v = (lp < ppc && part < parts && j < kin) ? (uint64_t)keys[((int64_t)part * n + row) * kin + j] : kSent64;
            v = warp_sort32_u64(v >= kEmpty ? kSent64 : v, lane);
            cur = first ? v : warp_merge32_u64(cur, v, lane);
            first = false;
        }
    }
    if (lane < kout) {
        const bool empty = cur == kSent64;
        Do[row * kout + lane] = empty ? 0x7FFFFFFF : (int32_t)(cur >> kXchgIdBits);
        Io[row * kout + lane] = empty ? -1 : (int64_t)(cur & ((1ull << kXchgIdBits) - 1));
    }
}

// ---- the exchange fused with its merge, over NVLink peer memory (snv_peer_*): ONE launch per row-sharded batch.
// Phase 1 (push): every rank turns its scan result into keys and STORES each one straight into the receive area of the
// rank that owns the query (peer memory mapped through CUDA IPC; posted 256-byte writes through the NVSwitch), fences,
// and the last block to finish raises this rank's flag on every peer (release, system scope) to the batch's epoch.
// Phase 2 (merge): wait until every source's flag has reached the epoch (acquire), then the warp-per-row network above
// reduces the [parts][n][k] lists that arrived to the k_out best.  No block waits for another block of its own launch, and
// the push depends on nothing remote, so the launch cannot deadlock however few of its blocks are resident (blocks that
// are not resident yet only delay their rank's flag).  Measured (tools/coexist_check.py): a 128-thread block does NOT
// become resident next to a scan CTA - 15 warps x 128 registers round up to the whole register file - so the exchange
// runs between two scans and is sized to be short (whole GPU) rather than hidden.  While its blocks wait for a late peer
// they hold their SMs, so a PIPELINED caller splits the phases (snv_peer_push right after scan i, snv_peer_merge after
// scan i + 1: by then every peer's push of batch i is a whole scan old and nobody spins); the fused form is for a caller
// that needs the result of this batch now.
struct PeerXchgParams {
    const int32_t* D;
    const int64_t* I;
    int nw, k, parts, rank, kout;
    uint32_t nq, qg;
    int64_t* const* peer_recv;  // [parts] each rank's receive area of this slot: keys [src][nw][qg][k]
    uint64_t* const* peer_flags;  // [parts] each rank's flag words (one per source, 16 bytes apart)
    const int64_t* my_recv;
    const uint64_t* my_flags;
    unsigned* counter;
    uint64_t epoch;
    int32_t* Do;
    int64_t* Io;
    int phases;  // bit 0: pack + push + flag, bit 1: wait + merge (3 = the fused exchange)
    uint64_t timeout_ns;  // a source that has not raised its flag after this long traps the kernel
};

__global__ void __launch_bounds__(128, 16)
peer_exchange_kernel(const PeerXchgParams p)
{
    const int lane = threadIdx.x & 31;
    // ---- phase 1: pack + push
    const uint32_t total = (p.phases & 1) ? (uint32_t)p.nw * p.nq * (uint32_t)p.k : 0u;
    const uint32_t uk = (uint32_t)p.k;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t row = i / uk, j = i - row * uk;
        const uint32_t w = row / p.nq, q = row - w * p.nq;
        const uint32_t g = q / p.qg, ql = q - g * p.qg;
        const int64_t id = p.I[i];
        const int64_t key = id < 0 ? 0x7FFFFFFFFFFFFFFFLL : (((int64_t)p.D[i] << kXchgIdBits) | (id & ((1LL << kXchgIdBits) - 1)));
        p.peer_recv[g][(((size_t)p.rank * p.nw + w) * p.qg + ql) * uk + j] = key;
    }
    if (p.phases & 1) __threadfence_system();
    __syncthreads();
    if ((p.phases & 1) && threadIdx.x == 0) {
        const unsigned prev = atomicAdd(p.counter, 1u);
        if (prev == gridDim.x - 1) {
            *p.counter = 0u;  // for the next launch (stream-ordered)
            __threadfence_system();
            for (int g = 0; g < p.parts; ++g)
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.peer_flags[g] + 2 * p.rank), "l"(p.epoch) : "memory");
        }
    }
    if (!(p.phases & 2)) return;
    // ---- phase 2: wait for every source, merge
    if ((int)threadIdx.x < p.parts) {
        const uint64_t* f = p.my_flags + 2 * threadIdx.x;
        uint64_t t0 = 0, v;
        for (uint32_t spin = 0;; ++spin) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (v >= p.epoch) break;
            if ((spin & 1023u) == 1023u) {
                uint64_t now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > p.timeout_ns) {  // a peer never arrived
                    printf("snv peer exchange: rank %d timed out waiting for rank %d (epoch %llu, flag %llu)\n", p.rank,
                           (int)threadIdx.x, (unsigned long long)p.epoch, (unsigned long long)v);
                    __trap();
                }
            }
            __nanosleep(spin < 64 ? 100 : 1000);  // the peers are usually a scan away: do not hammer L2 while waiting
        }
    }
    __syncthreads();
    constexpr uint64_t kEmpty = 0x7FFFFFFFFFFFFFFFull;
    const uint32_t n = (uint32_t)p.nw * p.qg;
    const int kin = p.k;
    const int ppc = kin <= 16 ? 32 / kin : 1;
    const int lp = kin <= 16 ? lane / kin : 0, lj = kin <= 16 ? lane - lp * kin : lane;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n; row += warps) {
        uint64_t cur = kSent64;
        for (int part0 = 0; part0 < p.parts; part0 += ppc) {
            const int part = part0 + lp;
            uint64_t v = kSent64;
            if (lp < ppc && part < p.parts && lj < kin) v = (uint64_t)__ldcg(p.my_recv + ((size_t)part * n + row) * kin + lj);
            v = warp_sort32_u64(v >= kEmpty ? kSent64 : v, lane);
            cur = part0 ? warp_merge32_u64(cur, v, lane) : v;
        }
        if (lane < p.kout) {
            const bool empty = cur == kSent64;
            p.Do[(size_t)row * p.kout + lane] = empty ? 0x7FFFFFFF : (int32_t)(cur >> kXchgIdBits);
            p.Io[(size_t)row * p.kout + lane] = empty ? -1 : (int64_t)(cur & ((1ull << kXchgIdBits) - 1));
        }
    }
}

int peer_exchange_launch(const int32_t* D, const int64_t* I, int nw, int64_t nq, int k, int parts, int rank, int64_t* const* peer_recv,
                         uint64_t* const* peer_flags, const int64_t* my_recv, const uint64_t* my_flags, unsigned* counter,
                         uint64_t epoch, int kout, int32_t* Do, int64_t* Io, int phases, cudaStream_t stream)
{
    if (parts < 1 || parts > 64 || nq % parts != 0) { set_error("peer exchange: the queries of a window must split evenly over <= 64 ranks"); return SNV_ERR_INVALID; }
    if (k < 1 || k > 32 || kout < 1 || kout > k) { set_error("peer exchange: k must be in [1, 32], k_out <= k"); return SNV_ERR_UNSUPPORTED; }
    if ((int64_t)nw * nq * k >= (int64_t)1 << 31) { set_error("peer exchange: batch too large (nw * nq * k must be < 2^31)"); return SNV_ERR_UNSUPPORTED; }
    PeerXchgParams p{};
    p.D = D; p.I = I; p.nw = nw; p.k = k; p.parts = parts; p.rank = rank; p.kout = kout;
    p.nq = (uint32_t)nq; p.qg = (uint32_t)(nq / parts);
    p.peer_recv = peer_recv; p.peer_flags = peer_flags; p.my_recv = my_recv; p.my_flags = my_flags;
    p.counter = counter; p.epoch = epoch; p.Do = Do; p.Io = Io; p.phases = phases;
    {
        // SNV_PEER_TIMEOUT_S: how long the merge waits for a rank that is behind (default 60 s) before it traps instead of hanging
        static const uint64_t timeout_s = [] { const char* e = getenv("SNV_PEER_TIMEOUT_S"); const long v = e ? atol(e) : 60; return (uint64_t)(v > 0 ? v : 60); }();
        p.timeout_ns = timeout_s * 1000000000ull;
    }
    // 8 blocks of 128 threads per SM (half the thread slots): the whole grid is resident on an idle GPU, so nobody spins
    // while a sibling waits for a slot, and the latency-bound push / merge get enough warps in flight
    const int64_t want = std::max<int64_t>((phases & 1) ? ceil_div((int64_t)nw * nq * k, 128 * 8) : 1,
                                           (phases & 2) ? ceil_div((int64_t)nw * (nq / parts), 4 * 2) : 1);
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)kNumSMs * 8));
    peer_exchange_kernel<<<grid, 128, 0, stream>>>(p);
    SNV_LAUNCH_CHECK();
    return SNV_OK;
}

int exchange_pack_launch(const int32_t* D, const int64_t* I, int nw, int64_t nq, int k, int parts, int64_t* keys, cudaStream_t stream)
{
    const int64_t total = (int64_t)nw * nq * k;
    if (total <= 0) return SNV_OK;
    if (parts < 1 || nq % parts != 0) { set_error("exchange_pack: the queries of a window must split evenly over the ranks"); return SNV_ERR_INVALID; }
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(total, 256), (int64_t)kNumSMs * 16);
    exchange_pack_kernel<<<grid, 256, 0, stream>>>(D, I, nw, nq, k, parts, keys);
    SNV_LAUNCH_CHECK();
    return SNV_OK;
}

int exchange_merge_launch(const int64_t* keys, int parts, int64_t n, int kin, int kout, int32_t* Do, int64_t* Io, cudaStream_t stream)
{
    if (n <= 0) return SNV_OK;
    if (kout < 1 || kout > 32) { set_error("exchange_merge: k_out must be in [1, 32]"); return SNV_ERR_UNSUPPORTED; }
    const unsigned grid = (unsigned)ceil_div(n, 256 / 32);
    if (kout <= 8) exchange_merge_kernel<8><<<grid, 256, 0, stream>>>(keys, parts, n, kin, kout, Do, Io);
    else exchange_merge_kernel<32><<<grid, 256, 0, stream>>>(keys, parts, n, kin, kout, Do, Io);
    SNV_LAUNCH_CHECK();
    return SNV_OK;
}

}  // namespace snv
