/*
 * snvknn — C ABI of the B200-native per-window exact k-NN engine (libsnvknn.so).
 *
 * This is the drop-in boundary for the one hot path of wangbaonan/RAG-SNVBERT: the faiss
 * `IndexFlatL2` / `IndexBinaryFlat` build + search (+ the gather that follows) which the
 * reference reaches through faiss's SWIG Python API.  The reference has no C ABI of its own
 * (it is pure Python); every entry point below cites the reference call site it replaces
 * (paths relative to the reference tree).  The Python mirror of the faiss surface lives in
 * rag_snvbert_b200/faiss_compat.py and binds these symbols with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types.
 *   - every function returns SNV_OK (0) or an snv_status; snv_last_error() gives the
 *     thread-local message of the last failure.
 *   - `*_on_device` / SNV_LOC_* flags say whether a pointer is host or device memory
 *     (device pointers must belong to the index's device).  Host buffers are copied on
 *     `stream` and the call returns after the stream has been synchronised; device buffers
 *     are used in place and the call is asynchronous on `stream`.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry point
 *     fails with SNV_ERR_CUDA.
 *
 * Packed haplotype rows (the native layout): uint32 words, site s in word s/32, bit s%32,
 * pad bits zero, row stride snv_packed_stride(d) words (a multiple of 4 => 16-byte rows for
 * bulk async copies).
 */
#ifndef SNVKNN_H
#define SNVKNN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNVKNN_VERSION 100

typedef enum snv_status {
    SNV_OK = 0,
    SNV_ERR_INVALID = 1,     /* bad argument (faiss: AssertionError from the SWIG wrapper) */
    SNV_ERR_CUDA = 2,        /* CUDA runtime / driver failure, or no device                */
    SNV_ERR_UNSUPPORTED = 3, /* valid request outside what the kernels implement           */
    SNV_ERR_NOMEM = 4
} snv_status;

typedef enum snv_kind {
    SNV_KIND_HAMMING = 0, /* bit-packed rows, integer Hamming / observed-site masked Hamming  */
    SNV_KIND_L2 = 1       /* float rows, squared L2 = |q|^2 + |r|^2 - 2 q.r on tcgen05       */
} snv_kind;

typedef enum snv_dtype {
    SNV_DT_U8 = 0,         /* one byte per site, non-zero = alt allele        [rows][d]       */
    SNV_DT_F32 = 1,        /* HAMMING: non-zero = alt allele; L2: the vector  [rows][d]       */
    SNV_DT_PACKED_U32 = 2, /* native packed rows            [rows][snv_packed_stride(d)]      */
    SNV_DT_PACKED_U8 = 3,  /* np.packbits rows (faiss binary codes) [rows][(d+7)/8] bytes     */
    SNV_DT_I64_TOKENS = 4, /* model tokens (5/6 alleles, 4 = MASK, 0/2/3 specials) [rows][d]  */
    SNV_DT_PACKED_U32_DENSE = 5 /* packed rows WITHOUT stride padding [rows][snv_packed_words(d)]: the wire format of
                                 * host-buffer sweeps (132 instead of 144 bytes per 1030-site row); re-strided on the device */
} snv_dtype;

typedef enum snv_mask_mode {
    SNV_MASK_NONE = 0,
    SNV_MASK_PER_WINDOW = 1, /* one mask row per window   [nw][row]      */
    SNV_MASK_PER_QUERY = 2   /* one mask row per query    [nw][nq][row]  */
} snv_mask_mode;

/* bit flags for `flags` arguments */
#define SNV_Q_ON_DEVICE 0x1u    /* queries (and mask) are device pointers                  */
#define SNV_OUT_ON_DEVICE 0x2u  /* D / I (or gather output) are device pointers            */
#define SNV_MASK_IS_MISSING 0x4u /* mask marks MISSING sites (reference convention,        */
                                 /* partial_faiss_intersect.py:91) instead of observed ones */
#define SNV_X_ON_DEVICE 0x8u    /* rows passed to add() are a device pointer               */

/* L2 cross-term precision */
#define SNV_L2_TF32 0   /* one tf32 pass: exact for small-integer inputs (tokens), ~1e-3 rel. otherwise */
#define SNV_L2_TF32X3 1 /* hi/lo split, three tf32 products: fp32-faithful                               */
/* OR-ed into l2_mode: subtract the column means of the first rows added from panel and queries
 * before the product (squared L2 is translation invariant).  Removes the large common component
 * of embedding vectors (position / allele-frequency terms) so that |q|^2 + |r|^2 - 2 q.r does not
 * cancel catastrophically.  Not for integer-valued rows (it would make their products inexact). */
#define SNV_L2_CENTER 0x10
/* OR-ed into l2_mode: decide SNV_L2_CENTER on the first add() - on when the rows hold anything but small integers
 * (embeddings), off for integer-valued rows (tokens, 0/1 genotypes).  The Python classes default to this. */
#define SNV_L2_CENTER_AUTO 0x20

typedef struct snv_index snv_index;

const char* snv_last_error(void);
int snv_version(void);
/* number of visible CUDA devices (0 when there is none); SNV_ERR_CUDA if the runtime fails */
int snv_device_count(int* count);

int64_t snv_packed_words(int64_t d);  /* ceil(d/32)                     */
int64_t snv_packed_stride(int64_t d); /* packed_words rounded up to 4   */

/*
 * faiss.IndexFlatL2(d) / faiss.IndexBinaryFlat(d_bits) constructors
 * (build_ref_db_l2.py:89, src/dataset/rag_train_dataset.py:132, test_faiss_intersect.py:173,
 *  src/dataset/embedding_rag_infer_dataset.py:176) — plus the windowed form: one object
 * holding `n_windows` independent panels of the same shape so that a whole window sweep
 * (batch_test_faiss_l2.py:80-110) is one launch.  `l2_mode` is ignored for HAMMING.
 */
int snv_index_create(int kind, int64_t d, int n_windows, int device, int l2_mode, snv_index** out);
void snv_index_free(snv_index* idx);

int64_t snv_index_ntotal(const snv_index* idx); /* rows per window (faiss index.ntotal) */
int64_t snv_index_d(const snv_index* idx);      /* faiss index.d                        */
int snv_index_kind(const snv_index* idx);
int snv_index_n_windows(const snv_index* idx);
int snv_index_device(const snv_index* idx);
int snv_index_reset(snv_index* idx);            /* faiss index.reset()                  */

/*
 * index.add(x)  (build_ref_db_l2.py:90, src/dataset/rag_train_dataset.py:133-134,
 * partial_faiss_intersect.py:103-105, test_faiss_intersect.py:152-155,173-178).
 * Appends `n` rows to EVERY window: x is [n_windows][n][row] in `dtype`.  The index keeps
 * its own device copy (bit-packed for HAMMING; fp32 + tf32 operand planes + |r|^2 for L2);
 * the caller may free x on return (host x) / after the stream reaches this point (device x).
 */
int snv_index_add(snv_index* idx, const void* x, int64_t n, int dtype, unsigned flags, void* stream);

/*
 * index.search(q, k) -> (D, I)  (src/dataset/rag_train_dataset.py:281,
 * batch_test_faiss_l2.py:110, test_faiss.py:135, partial_faiss_intersect.py:109,
 * test_faiss_intersect.py:160,181, src/dataset/embedding_rag_infer_dataset.py:284-285;
 * torch.cdist+topk at src/dataset/embedding_rag_dataset.py:397-402).
 *
 * Windows [w0, w0+nw) are searched with q = [nw][nq][row] (`q_dtype`).  Outputs are
 * [nw][nq][k]: I int64 row ids (+ id_offset; -1 when fewer than k rows exist),
 * D_i32 (HAMMING only) and/or D_f32 (squared L2 == Hamming for 0/1 rows); either D pointer may
 * be NULL.  Results are in the canonical order (distance ascending, id ascending), ties at
 * the k-boundary keep the lowest ids.  Padding: D_i32 = INT32_MAX, D_f32 = FLT_MAX.
 * LIMIT of this entry point: k <= 32 (the running top-k lives in registers); larger k -> SNV_ERR_UNSUPPORTED.  faiss
 * takes any k (every call site of the reference asks for k <= 5: src/train.py:105, src/infer.py:66, --top_k default 5);
 * the Python classes serve k > 32 exactly through a block path built on this call: the panel re-added as windows of
 * 32 rows, one k = 32 search returning every distance of every block, then a top-k over the (distance, id) keys.
 *
 * HAMMING only: `mask` (same dtype family as q: U8/F32/PACKED_U32) restricts the distance to
 * observed sites, popc((q ^ r) & m) — partial_faiss_intersect.py:82-111.
 */
int snv_index_search(snv_index* idx, int w0, int nw, const void* q, int64_t nq, int q_dtype,
                     const void* mask, int mask_mode, int k, int64_t id_offset, int32_t* D_i32,
                     float* D_f32, int64_t* I, unsigned flags, void* stream);

/*
 * The same search with COMPACT results for host-buffer callers of the offline sweeps (batch_test_faiss_l2.py:109-111
 * keeps every window's (D, I)): D uint16 [nw][nq][k], I int32 [nw][nq][k] - 6 instead of 12 bytes per neighbour
 * over PCIe, identical values (ids are row numbers inside the window: no id_offset; padding: I = -1, D = 0xFFFF).
 * HAMMING indexes with ntotal < 2^31 and d < 65535; anything else -> SNV_ERR_UNSUPPORTED.
 */
int snv_index_search_compact(snv_index* idx, int w0, int nw, const void* q, int64_t nq, int q_dtype,
                             const void* mask, int mask_mode, int k, uint16_t* D_u16, int32_t* I_i32,
                             unsigned flags, void* stream);

/*
 * Ragged per-window batches: a training / inference batch regrouped by window_idx before the
 * search (src/dataset/rag_train_dataset.py:239-281, src/dataset/embedding_rag_dataset.py:318-321,
 * src/dataset/sampler.py:58-119).  q is [nq_total][row] in caller order, window_ids (HOST int32
 * [nq_total]) names the window of every query; results come back in caller order
 * ([nq_total][k]).  One launch for the whole batch whatever the group sizes.  HAMMING indexes;
 * `mask` is NULL or per query ([nq_total][row]).
 */
int snv_index_search_grouped(snv_index* idx, const void* q, const int32_t* window_ids, int64_t nq_total,
                             int q_dtype, const void* mask, int k, int32_t* D_i32, float* D_f32,
                             int64_t* I, unsigned flags, void* stream);

/*
 * The gather that follows search in the V17 collate
 * (src/dataset/rag_train_dataset.py:287-307): I -> retrieved haplotype -> tokens with an
 * all-zero mask: out[w][q][j][:] = [SOS=2] + (5|6 per site) + [EOS=3] + PAD(0) up to seq_len.
 * HAMMING indexes only.  n_sites: NULL (every window has d sites) or int32 [nw] real site
 * counts (host pointer).  I: [nw][nq][k] (location per SNV_Q_ON_DEVICE), out int64
 * [nw][nq][k][seq_len] (location per SNV_OUT_ON_DEVICE).  I == -1 -> all PAD.
 */
int snv_index_gather_tokens(snv_index* idx, int w0, int nw, const int64_t* I, int64_t nq, int k,
                            const int32_t* n_sites, int seq_len, int64_t* out, unsigned flags,
                            void* stream);
/* same for a ragged batch: I [nq_total][k] with window_ids (HOST int32 [nq_total]); n_sites is NULL
 * or HOST int32 [n_windows] */
int snv_index_gather_tokens_grouped(snv_index* idx, const int64_t* I, const int32_t* window_ids,
                                    int64_t nq_total, int k, const int32_t* n_sites, int seq_len,
                                    int64_t* out, unsigned flags, void* stream);

/*
 * Row gather for L2 indexes (src/dataset/embedding_rag_dataset.py:406-438,
 * src/dataset/embedding_rag_infer_dataset.py:287-319 once re-embedding is per row):
 * out[w][q][j][:] = panel[w][I[w][q][j]][:]  float32 [nw][nq][k][d]; I == -1 -> zeros.
 */
int snv_index_gather_rows(snv_index* idx, int w0, int nw, const int64_t* I, int64_t nq, int k,
                          float* out, unsigned flags, void* stream);

/*
 * faiss.write_index / read_index support (build_ref_db_l2.py:93, batch_test_faiss_l2.py:94):
 * copy the stored rows of window w to the HOST buffer `out`
 * (HAMMING: uint32 [ntotal][snv_packed_stride(d)]; L2: float32 [ntotal][d]).
 */
int snv_index_export(snv_index* idx, int window, void* out);

/*
 * k-way merge of per-shard results (row-sharded panel; new — the reference never shards,
 * SURVEY.md §2.2).  D/I are DEVICE arrays [parts][nq][k_in] with GLOBAL ids (what an
 * all-gather of per-rank search outputs produces); writes the best k_out per query by
 * (distance, id) to the device arrays Do/Io [nq][k_out].  Exactly one of D_i32/D_f32 (and
 * the matching output) is non-NULL.
 */
int snv_topk_merge(int device, const int32_t* D_i32, const float* D_f32, const int64_t* I, int parts,
                   int64_t nq, int k_in, int k_out, int32_t* Do_i32, float* Do_f32, int64_t* Io,
                   void* stream);

/*
 * The exchange step of a row-sharded search (one rank per GPU; the collective itself is the caller's, e.g.
 * torch.distributed.all_to_all_single over NCCL).  snv_exchange_pack turns this rank's results D / I [nw][nq][k]
 * (global ids) into int64 keys (distance << 40 | id; missing -> INT64_MAX) laid out [parts][nw][nq / parts][k], i.e.
 * grouped by the rank that owns each query, so that one all-to-all delivers [parts][n][k] (n = nw * nq / parts rows,
 * one sorted list per source rank); snv_exchange_merge reduces that to the k_out best per row, Do_i32 / Io [n][k_out].
 * All pointers are DEVICE pointers; ids must be < 2^40, nq % parts == 0, k_out <= 32.
 */
int snv_exchange_pack(int device, const int32_t* D_i32, const int64_t* I, int nw, int64_t nq, int k, int parts,
                      int64_t* keys, void* stream);
int snv_exchange_merge(int device, const int64_t* keys, int parts, int64_t n, int k_in, int k_out, int32_t* Do_i32,
                       int64_t* Io, void* stream);

/*
 * The same exchange over NVLink peer memory, fused into ONE kernel launch per batch (no collective library on the data
 * path).  Every rank creates one exchange object (snv_peer_create: a device buffer of a 4 KiB header + two receive slots
 * of `slot_bytes`, and its 64-byte CUDA IPC handle); the host side all-gathers the handles with whatever transport it
 * has (torch.distributed, MPI, a file) and snv_peer_open maps the peers' buffers.  snv_peer_exchange then, on `stream`:
 * packs this rank's D / I [nw][nq][k] into int64 keys, stores each key straight into the receive slot of the rank that
 * owns the query, raises this rank's flag on every peer (release), waits for every peer's flag (acquire), and merges the
 * lists that arrived into Do_i32 / Io [nw * nq / world][k_out] - the result of snv_exchange_pack + all-to-all +
 * snv_exchange_merge.  Rules: every rank makes the same sequence of calls (batch sizes equal on all ranks); all calls of
 * one object go to ONE stream (slot reuse is ordered by the previous batch's flags); nw * nq * k * 8 <= slot_bytes;
 * nq % world == 0; k <= 32; ids < 2^40.  A peer that never arrives traps the kernel after SNV_PEER_TIMEOUT_S seconds (default 60) instead of hanging it.
 * snv_peer_open_local wires objects that live in ONE process on one device by pointer (test hook: their exchanges must
 * then run on different streams, since each waits for the others' pushes).
 * Replaces, for the row-sharded search of a multi-GPU job, what the reference does on one GPU with a single faiss index
 * (src/dataset/rag_train_dataset.py:239-281 search + :283 result use); the NCCL route above stays for GPUs without
 * peer access.
 */
#define SNV_PEER_HANDLE_BYTES 64
typedef struct snv_peer snv_peer;
int snv_peer_create(int device, int rank, int world, size_t slot_bytes, snv_peer** out, void* handle_out);
int snv_peer_open(snv_peer* peer, const void* handles /* [world][SNV_PEER_HANDLE_BYTES], rank order */);
int snv_peer_open_local(snv_peer* peer, snv_peer* const* peers /* [world] */);
int snv_peer_exchange(snv_peer* peer, const int32_t* D_i32, const int64_t* I, int nw, int64_t nq, int k, int k_out,
                      int32_t* Do_i32, int64_t* Io, void* stream);
/* The two phases separately, for a pipelined caller: snv_peer_push (pack + push + flag; returns the batch's epoch) right
 * after scan i, snv_peer_merge (wait for every source's flag of that epoch + merge) after scan i + 1 has been queued - by
 * then the peers' pushes are a whole scan old and no block of the merge waits holding an SM.  Same stream, same call
 * sequence on every rank; a push may run at most one batch ahead of the merge of the previous one (two slots). */
int snv_peer_push(snv_peer* peer, const int32_t* D_i32, const int64_t* I, int nw, int64_t nq, int k, uint64_t* epoch_out,
                  void* stream);
int snv_peer_merge(snv_peer* peer, uint64_t epoch, int nw, int64_t nq, int k, int k_out, int32_t* Do_i32, int64_t* Io,
                   void* stream);
int snv_peer_destroy(snv_peer* peer);

/* device-side pack helper: rows in `dtype` (U8 / F32 / PACKED_U8 / I64_TOKENS) ->
 * packed uint32 [rows][snv_packed_stride(d)] (device pointers; all on `stream`).
 * For I64_TOKENS `out_observed` (nullable) receives the observed-site plane (token in {5,6}). */
int snv_pack_rows(int device, const void* x, int64_t rows, int64_t d, int dtype, int invert,
                  uint32_t* out, uint32_t* out_observed, void* stream);

/*
 * Observed-site masks from a position intersection (replaces the host-side
 * np.intersect1d of test_faiss_intersect.py:128-140 and the searchsorted of
 * src/dataset/embedding_rag_dataset.py:117-128).  All pointers are DEVICE pointers:
 * ref_pos int64 [n_ref] (the panel's site positions), tgt_pos int64 [n_tgt] ASCENDING,
 * window_info int64 [n_windows][2] = (start, end) site indices into ref_pos.  Writes packed
 * uint32 [n_windows][snv_packed_stride(d)]: bit c of window w is set when site start + c / ploidy
 * (< end) is also a target position.  ploidy 1 = haplotype rows, 2 = the offline scripts' sample rows
 * (s0h0, s0h1, ...).  The result is what snv_index_search takes as a PER_WINDOW observed mask.
 */
int snv_intersect_masks(int device, const int64_t* ref_pos, int64_t n_ref, const int64_t* tgt_pos, int64_t n_tgt,
                        const int64_t* window_info, int n_windows, int64_t d, int ploidy, uint32_t* out, void* stream);

/* number of kernels this library has launched in the calling process (bench "gpu_launches") */
int64_t snv_launch_count(void);

/* Measurement hook (bench.py roofline): when enabled, every launch of a dominant kernel
 * (hamming_topk_kernel / hamming_tc_kernel / l2_topk_kernel) is bracketed by CUDA events on its own stream;
 * snv_profile_last_ms() waits for the most recent such launch and returns its duration. */
int snv_profile_enable(int on);
int snv_profile_last_ms(float* ms);

/* Which engine the most recent Hamming search of the calling process ran on:
 * 0 = popcount scan (hamming_topk_kernel); 1 = tensor cores, fp8 operands (hamming_tc_kernel: tcgen05.mma over the
 * bit-packed panel expanded in shared memory); 2 = its bring-up variant (panel expanded in HBM); 3 = tensor cores,
 * fp4 block-scaled operands; 4 = fp4 on CTA pairs (cta_group::2); -1 = none yet.  The choice is made per call from
 * the shape (many queries per window -> tensor cores); SNV_HAMMING_ENGINE=popc|tc|tc_hbm|tc4|tc4x2 forces one. */
int snv_last_hamming_engine(void);

/* Host-only view of the tensor-core engine's planner (no device work; test hook for tests/test_planner.py): for a
 * uniform search of n_windows x nq queries against n rows of d sites at top-k, writes
 * plan_out[12] = (engine, k capacity, k-blocks, query tiles, panel tiles, row splits, tiles per split, id bits,
 * tail items, tail split, tail tiles, workspace KiB) and up to `cap` work items
 * items_out[i][8] = (window, query tile, first panel tile, tiles, piece, pieces, partial-key row base, CTA slot),
 * enumerated with the decoder the kernel itself uses; *n_items = how many there are (may exceed cap). */
int snv_debug_hamming_plan(int n_windows, int nq, int64_t n, int d, int k, int32_t* plan_out, int64_t* items_out,
                           int64_t cap, int64_t* n_items);

/* Host-only: the narrow-float operand codes the tensor-core engine derives from one packed word of a query (with its
 * observed-site mask word) and of a panel row, computed by the same functions the kernels run.  fp4 != 0: E2M1, 16
 * bytes each (2 codes per byte, 32 sites); fp4 == 0: E4M3, 32 bytes each (32 sites).  Byte i of q_codes multiplies
 * byte i of panel_codes in the contraction; the sum over the word is popc((q ^ r) & m) - popc(q & m). */
int snv_debug_tc_codes(int fp4, uint32_t q_word, uint32_t mask_word, uint32_t panel_word, uint8_t* q_codes, uint8_t* panel_codes);

/* Host-only: the window-chunk boundaries snv_index_search would pipeline that search over (host_io != 0: queries or
 * results in host memory; 0: everything device resident = one chunk).  bounds_out[0 .. *n_bounds) = 0 < ... < n_windows. */
int snv_debug_hamming_chunks(int n_windows, int nq, int64_t n, int d, int k, int host_io, int32_t* bounds_out, int cap,
                             int* n_bounds);

#ifdef __cplusplus
}
#endif
#endif /* SNVKNN_H */
