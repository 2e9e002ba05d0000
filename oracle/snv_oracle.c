/*
 * CPU restatement (plain C + OpenMP) of the reference's exact per-window k-NN on bit-packed
 * haplotypes.  TEST INFRASTRUCTURE / CPU BASELINE ONLY: linked by tests/, smoke() and the
 * cpu_baseline / --impl reference legs of bench.py; never by the product path.
 *
 * What it restates (paths relative to /root/reference):
 *   - faiss.IndexBinaryFlat.search as called at test_faiss_intersect.py:173-181: Hamming
 *     distance by 64-bit popcount over packed codes, per-query bounded heap of the k best,
 *     results ordered by (distance, id).  faiss itself (facebookresearch/faiss,
 *     utils/hamming.cpp `hammings_knn_hc`) is an un-vendored, un-pinned dependency of the
 *     reference; this file restates its published algorithm -> "parity unpinned" by any
 *     reference-owned known-answer test (Oracle).
 *   - the observed-site restriction of partial_faiss_intersect.py:82-111 (distance over
 *     columns where mask==0 only) as popc((q ^ r) & observed_mask).
 *   - faiss.IndexFlatL2.search on 0/1 vectors (batch_test_faiss_l2.py:110): identical
 *     ranking, D = the same integers as float.
 *
 * Packed layout: uint32 words, site s -> word s/32, bit s%32, pad bits zero, row stride
 * `stride` words (>= words).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SNV_ORACLE_MAXK 1024

static inline int pair_dist(const uint32_t* q, const uint32_t* r, const uint32_t* m, int words)
{
    int d = 0, w = 0;
    if (m) {
        for (; w + 2 <= words; w += 2) {
            uint64_t a, b, c;
            memcpy(&a, q + w, 8); memcpy(&b, r + w, 8); memcpy(&c, m + w, 8);
            d += __builtin_popcountll((a ^ b) & c);
        }
        for (; w < words; ++w) d += __builtin_popcount((q[w] ^ r[w]) & m[w]);
    } else {
        for (; w + 2 <= words; w += 2) {
            uint64_t a, b;
            memcpy(&a, q + w, 8); memcpy(&b, r + w, 8);
            d += __builtin_popcountll(a ^ b);
        }
        for (; w < words; ++w) d += __builtin_popcount(q[w] ^ r[w]);
    }
    return d;
}

/* one query against one panel: sorted insertion on the key (dist << 32 | id), which is the
 * (distance, id) lexicographic order. */
static void one_query(const uint32_t* panel, int64_t n, int stride, const uint32_t* q,
                      const uint32_t* m, int words, int k, int32_t* D, int64_t* I)
{
    uint64_t best[SNV_ORACLE_MAXK];
    int filled = 0;
    for (int64_t j = 0; j < n; ++j) {
        uint64_t key = ((uint64_t)(uint32_t)pair_dist(q, panel + j * stride, m, words) << 32) | (uint64_t)j;
        if (filled == k && key >= best[k - 1]) continue;
        int p = filled < k ? filled++ : k - 1;
        while (p > 0 && best[p - 1] > key) { best[p] = best[p - 1]; --p; }
        best[p] = key;
    }
    for (int i = 0; i < k; ++i) {
        if (i < filled) { D[i] = (int32_t)(best[i] >> 32); I[i] = (int64_t)(best[i] & 0xffffffffu); }
        else            { D[i] = INT32_MAX; I[i] = -1; }
    }
}

/*
 * panel   [n_windows][n][stride]      queries [n_windows][nq][stride]
 * mask    NULL | [n_windows][nq][stride] (per query, 1 = observed)
 * D       [n_windows][nq][k] int32    I [n_windows][nq][k] int64 (row id inside the window)
 * returns 0, or -1 on bad arguments.
 */
int snv_oracle_hamming_topk(const uint32_t* panel, const uint32_t* queries, const uint32_t* mask,
                            int64_t n_windows, int64_t n, int64_t nq, int words, int stride,
                            int k, int32_t* D, int64_t* I, int n_threads)
{
    if (k < 1 || k > SNV_ORACLE_MAXK || words < 1 || stride < words) return -1;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    int64_t total = n_windows * nq;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t t = 0; t < total; ++t) {
        int64_t w = t / nq;
        const uint32_t* q = queries + t * stride;
        const uint32_t* m = mask ? mask + t * stride : NULL;
        one_query(panel + w * n * stride, n, stride, q, m, words, k, D + t * k, I + t * k);
    }
    return 0;
}

int snv_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
