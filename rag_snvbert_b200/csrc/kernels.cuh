// Host-callable launchers of the sm_100a kernels (internal; the public surface is include/snvknn.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/snvknn.h"

namespace snv {

// ---------------------------------------------------------------- Hamming scan + top-k
struct HammingSearchParams {
    // inputs (device)
    const uint32_t* panel;     // [n_windows][cap][stride]; already offset to window w0
    int64_t panel_win_stride;  // words between consecutive windows
    const uint32_t* q;         // [nw][nq][stride]
    const uint32_t* mask;      // nullptr or observed-site rows
    int64_t mask_win_stride;   // words
    int64_t mask_q_stride;     // words; 0 = one mask row shared by the window
    int words;                 // ceil(d/32)
    int stride;                // words per packed row (bucketed, multiple of 4)
    int d;                     // sites (bits)
    int64_t n;                 // rows per window
    int nq;                    // queries per window
    int nw;                    // windows in this call
    int k;                     // neighbours requested
    int64_t id_offset;         // added to every id written to I
    // outputs (device) [nw][nq][k]
    int32_t* D_i32;
    float* D_f32;
    int64_t* I;
    uint64_t* partial;  // workspace, required when nsplit > 1: [nw*nq][nsplit][kt] keys
    // grouped (ragged) mode: work items (window, start, count) over the window-sorted permutation `order`
    const int32_t* work;   // device [n_work][3]; nullptr = uniform [nw][nq] mode
    const int32_t* order;  // device [nq_total]
    int n_work, work_block;
    int64_t nq_total;
    // plan (filled by hamming_plan)
    int block, qtiles, nsplit, rows_per_split, idx_bits, kt, nw_templ, tile_rows, stages;
    uint32_t wt1, wt2, wt4;  // 1, 2, 4 as runtime values (keeps the popcount adds on the FMA pipe)
    size_t smem_bytes;
};

// Fills the plan fields; returns the partial-key workspace bytes needed (0 when nsplit == 1),
// or (size_t)-1 with the error set when the shape is unsupported.
size_t hamming_plan(HammingSearchParams& p);
int hamming_launch(const HammingSearchParams& p, cudaStream_t stream);

// Tensor-core engine for the same search (hamming_tc.cu): chosen per call by shape.
struct HammingTcPlan {
    int engine;  // 0 = popcount kernel (hamming_launch), 1 = tcgen05 fp8 with in-SM bit expansion, 2 = its bring-up variant,
                 // 3 = tcgen05 fp4 (block-scaled), 4 = fp4 on CTA pairs
    int kt, kblocks, qtiles, n_tiles, nsplit, tiles_per_split, idx_bits;
    int tail_items, tail_split, tail_tiles;     // trailing items cut into row ranges so that the last round fills the machine
    int64_t off_bias, off_partial, off_panel;  // byte offsets inside the workspace
};
// Fills `plan` (p needs its shape fields and mask null-ness only); returns the workspace bytes the
// launch needs (0 when plan.engine == 0), or (size_t)-1 with the error set.
size_t hamming_tc_plan(const HammingSearchParams& p, HammingTcPlan& plan);
int hamming_tc_launch(const HammingSearchParams& p, const HammingTcPlan& plan, void* ws, cudaStream_t stream);
// host only: the work items of that launch, out [cap][8]; returns how many there are (see hamming_tc.cu)
int64_t hamming_tc_debug_items(const HammingSearchParams& p, const HammingTcPlan& plan, int64_t* out, int64_t cap);
// host only: operand codes of one packed word (query & mask -> A chunk, panel -> B chunk), fp4: 16 bytes, fp8: 32
void hamming_tc_debug_codes(int fp4, uint32_t q, uint32_t m, uint32_t r, uint8_t* q_out, uint8_t* r_out);

// ---------------------------------------------------------------- key merge / finalize
// keys [nq_total][parts][kin] (uint64: hi = distance bits, lo = id, ~0 = empty) -> top k.
// float_dist: hi word holds float bits (L2) instead of an integer distance.
int merge_keys_launch(const uint64_t* keys, int parts, int kin, int64_t nq_total, int k,
                      int64_t id_offset, bool float_dist, int32_t* D_i32, float* D_f32, int64_t* I,
                      cudaStream_t stream);
// (D, I) [parts][nq][kin] arrays -> top k_out (row-sharded panel merge)
int merge_results_launch(const int32_t* D_i32, const float* D_f32, const int64_t* I, int parts,
                         int64_t nq, int kin, int kout, int32_t* Do_i32, float* Do_f32,
                         int64_t* Io, cudaStream_t stream);

// (int32 D, int64 I) -> (uint16 D, int32 I): the compact wire format of host-buffer searches
int narrow_results_launch(const int32_t* D, const int64_t* I, int64_t n, uint16_t* D16, int32_t* I32, cudaStream_t stream);

// row-sharded panel: per-rank (D, I) -> destination-major int64 exchange keys; received keys -> merged (D, I)
int exchange_pack_launch(const int32_t* D, const int64_t* I, int nw, int64_t nq, int k, int parts, int64_t* keys, cudaStream_t stream);
int exchange_merge_launch(const int64_t* keys, int parts, int64_t n, int kin, int kout, int32_t* Do, int64_t* Io, cudaStream_t stream);
// pack + NVLink push + flag + wait + merge in one launch (snv_peer_exchange)
int peer_exchange_launch(const int32_t* D, const int64_t* I, int nw, int64_t nq, int k, int parts, int rank, int64_t* const* peer_recv,
                         uint64_t* const* peer_flags, const int64_t* my_recv, const uint64_t* my_flags, unsigned* counter,
                         uint64_t epoch, int kout, int32_t* Do, int64_t* Io, int phases, cudaStream_t stream);

// ---------------------------------------------------------------- pack
int pack_launch(const void* x, int64_t rows, int64_t d, int dtype, bool invert, int stride,
                uint32_t* out, uint32_t* out_observed, cudaStream_t stream);

// ---------------------------------------------------------------- position intersection -> observed-site masks
int intersect_masks_launch(const int64_t* ref_pos, int64_t n_ref, const int64_t* tgt_pos, int64_t n_tgt,
                           const int64_t* window_info, int n_windows, int64_t d, int ploidy, int stride, uint32_t* out,
                           cudaStream_t stream);

// ---------------------------------------------------------------- gather
int gather_tokens_launch(const uint32_t* panel, int64_t panel_win_stride, int stride, int64_t n,
                         const int64_t* I, int64_t id_offset, int nw, int64_t nq, int k,
                         const int32_t* n_sites_dev, int d, int seq_len, int64_t* out,
                         cudaStream_t stream);
// ragged batch: meta [nq][2] = (window, n_sites) per query
int gather_tokens_grouped_launch(const uint32_t* panel, int64_t panel_win_stride, int stride, int64_t n,
                                 const int64_t* I, const int32_t* meta, int64_t nq, int k, int seq_len,
                                 int64_t* out, cudaStream_t stream);
int gather_rows_launch(const float* panel, int64_t panel_win_stride, int64_t d, int64_t n,
                       const int64_t* I, int nw, int64_t nq, int k, float* out, cudaStream_t stream);

// ---------------------------------------------------------------- float L2 (tcgen05)
struct L2SearchParams {
    const float* ref_ops;   // [N][kp] tf32 operand rows of the panel (B operand, K-major)
    const float* ref_norm;  // [N] |r|^2 (fp32, from the original fp32 rows)
    const float* q_ops;     // [nq][kp] tf32 operand rows of the queries (A operand)
    const float* q_norm;    // [nq]
    int64_t n;              // panel rows
    int64_t nq;
    int kp;                 // padded operand depth (multiple of 32)
    int k;
    int64_t id_offset;
    float* D_f32;           // [nq][k]
    int64_t* I;             // [nq][k]
    uint64_t* partial;      // [nq][nsplit][kt]
    int nsplit, kt, tiles_per_split;
    // split-K mode (skinny problems): partial dot products accumulate into dot[nq][dot_ld]
    float* dot;
    int64_t dot_ld;
    int ksplit, kb_per_split;
    bool pair;  // CTA pairs (cta_group::2): two query tiles per cluster, each CTA streams half of every panel tile
    int one;    // 1 as a runtime value (keeps the epilogue's mask adds on the FMA pipe)
};
size_t l2_plan(L2SearchParams& p);
int l2_launch(const L2SearchParams& p, cudaStream_t stream);
// fp32 rows [rows][d] -> operand rows [rows][kp] (mode TF32: x | 0-pad; TF32X3 with
// is_query: hi|lo|hi, panel: hi|hi|lo) and squared norms.
// norm_parts: scratch [rows][l2_prep_chunks(d)] floats, needed when l2_prep_chunks(d) > 1 (|x|^2 is then summed from
// per-chunk partial sums in a fixed order: deterministic)
int l2_prep_launch(const float* x, const float* mean, int64_t rows, int64_t d, int mode, bool is_query, int kp,
                   float* ops, float* norms, float* norm_parts, cudaStream_t stream);
int l2_prep_chunks(int64_t d);
// 1 when every value of x [count] is integral and |x| <= 2^11 (token / genotype vectors: products exact in tf32), else 0;
// flag is a device int the kernel ANDs into (initialise to 1)
int l2_integral_check_launch(const float* x, int64_t count, int* flag, cudaStream_t stream);
// column means of x [rows][d] (centering; squared L2 is translation invariant)
// (scratch: l2_colmean_scratch_bytes(rows, d) bytes of device memory; the result is deterministic)
int l2_colmean_launch(const float* x, int64_t rows, int64_t d, float* mean, float* scratch, cudaStream_t stream);
size_t l2_colmean_scratch_bytes(int64_t rows, int64_t d);
int l2_operand_depth(int64_t d, int mode);

}  // namespace snv
