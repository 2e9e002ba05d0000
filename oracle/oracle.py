"""CPU oracle for the per-window exact k-NN path of RAG-SNVBERT.

TEST INFRASTRUCTURE ONLY.  Nothing under ``rag_snvbert_b200/`` may import this
module: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, as the checker (never as the
thing measured as the product, never shipped).

Parity status ("Oracle"):
  * layout / tokenisation / gather / embedding-space ``cdist + topk`` functions are
    PINNED: ``tests/golden/make_golden.py`` imports the reference's own Python from
    /root/reference (with stub modules for the absent faiss/h5py/allel/vcfpy) and
    the golden vectors it wrote are checked against this file in
    ``tests/test_oracle_golden.py``.
  * the faiss arithmetic itself (``IndexFlatL2.search`` / ``IndexBinaryFlat.search``)
    lives in an un-vendored, un-pinned third-party dependency (facebookresearch/faiss,
    no version in the reference tree) and the reference has no test that pins its
    output, so for that piece the oracle is a restatement of faiss's published
    algorithm: "parity unpinned" by any reference-owned known-answer test.

All citations are relative to /root/reference.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# Token layout (src/dataset/vocab.py:84-95 specials; :141 counter order 0 -> 5, 1 -> 6;
# :153-170 to_seq; src/dataset/utils.py:121-132 sequence_padding)
# --------------------------------------------------------------------------------------
PAD, UNK, SOS, EOS, MASK = 0, 1, 2, 3, 4
ALLELE0, ALLELE1 = 5, 6
MAX_SEQ_LEN = 1030  # src/dataset/rag_train_dataset.py:18-19

I32_MAX = np.int32(2**31 - 1)
F32_MAX = np.float32(3.4028234663852886e38)


def tokenize(seq01: np.ndarray, mask: np.ndarray | None = None, seq_len: int = MAX_SEQ_LEN) -> np.ndarray:
    """Restates TrainDataset.tokenize (src/dataset/dataset.py:597-625) with
    WordVocab.to_seq(with_sos=True, seq_len=MAX_SEQ_LEN) (src/dataset/vocab.py:153-170):
    ``[SOS] + (0->5, 1->6, other->UNK) + [EOS]`` then PAD to ``seq_len`` (or truncate),
    then positions where the (already padded) ``mask`` is non-zero become MASK.
    ``seq01``: [..., L_w] integer array.  Returns int64 [..., seq_len]."""
    seq01 = np.asarray(seq01)
    lead = seq01.shape[:-1]
    flat = seq01.reshape(-1, seq01.shape[-1])
    n, lw = flat.shape
    body = np.full((n, lw), UNK, dtype=np.int64)
    body[flat == 0] = ALLELE0
    body[flat == 1] = ALLELE1
    full = np.concatenate(
        [np.full((n, 1), SOS, np.int64), body, np.full((n, 1), EOS, np.int64)], axis=1
    )
    if full.shape[1] <= seq_len:
        out = np.zeros((n, seq_len), dtype=np.int64)
        out[:, : full.shape[1]] = full
    else:
        out = full[:, :seq_len].copy()
    if mask is not None:
        m = np.asarray(mask).astype(bool)
        if m.ndim == 1:
            m = m[None, :]
        out = np.where(m, MASK, out)
    return out.reshape(*lead, seq_len)


def sequence_padding(raw_mask: np.ndarray, seq_len: int = MAX_SEQ_LEN) -> np.ndarray:
    """VCFProcessingModule.sequence_padding (src/dataset/utils.py:121-132): one 0 in
    front (the SOS slot), zeros behind up to seq_len."""
    raw_mask = np.asarray(raw_mask)
    out = np.zeros(seq_len, dtype=raw_mask.dtype)
    out[1 : 1 + raw_mask.shape[0]] = raw_mask
    return out


def panel_rows_from_gt(ref_gt_window: np.ndarray) -> np.ndarray:
    """[L_w, S, 2] genotype cube -> [2S, L_w] haplotype rows, row = 2*sample + hap
    (src/dataset/rag_train_dataset.py:111-118: reshape(L_w, -1).T)."""
    lw = ref_gt_window.shape[0]
    return ref_gt_window.reshape(lw, -1).T


# --------------------------------------------------------------------------------------
# Bit packing
# --------------------------------------------------------------------------------------
def pack_bits_u32(x01: np.ndarray, stride_words: int | None = None) -> np.ndarray:
    """[n, d] 0/1 -> uint32 [n, stride] ; site s lives in word s//32, bit s%32 (LSB first);
    pad bits and pad words are zero.  This is the engine's native packed layout."""
    x01 = np.asarray(x01)
    n, d = x01.shape
    nw = (d + 31) // 32
    stride = nw if stride_words is None else stride_words
    bits = np.zeros((n, nw * 32), dtype=np.uint8)
    bits[:, :d] = (x01 != 0)
    by = np.packbits(bits, axis=1, bitorder="little")  # [n, nw*4] bytes, LSB-first
    words = by.view("<u4").reshape(n, nw)
    out = np.zeros((n, stride), dtype=np.uint32)
    out[:, :nw] = words
    return out


def packbits_msb(x01: np.ndarray) -> np.ndarray:
    """bitpack_2d_array (test_faiss_intersect.py:46-54): np.packbits(axis=1), MSB first,
    last byte zero padded.  Input of faiss.IndexBinaryFlat."""
    return np.packbits(np.asarray(x01).astype(np.uint8), axis=1)


# --------------------------------------------------------------------------------------
# Exact top-k with the canonical total order (distance ascending, index ascending)
# --------------------------------------------------------------------------------------
def _topk_lex(dist: np.ndarray, k: int, pad_d):
    """dist [nq, N] -> (D [nq,k], I [nq,k]) smallest k by (dist, idx); short rows padded
    with I=-1, D=pad_d (faiss HeapArray semantics: unfilled slots keep id -1)."""
    nq, n = dist.shape
    kk = min(k, n)
    # stable argsort on distance == lexicographic (distance, index)
    order = np.argsort(dist, axis=1, kind="stable")[:, :kk]
    D = np.take_along_axis(dist, order, axis=1)
    I = order.astype(np.int64)
    if kk < k:
        D = np.concatenate([D, np.full((nq, k - kk), pad_d, dtype=dist.dtype)], axis=1)
        I = np.concatenate([I, np.full((nq, k - kk), -1, dtype=np.int64)], axis=1)
    return D, I


def hamming_matrix(panel01: np.ndarray, queries01: np.ndarray, mask01: np.ndarray | None = None,
                   block: int = 256) -> np.ndarray:
    """Integer distance matrix [nq, N].  mask01: None, [d] (shared) or [nq, d] (per query),
    1 = site OBSERVED (counted).  Distance = popc((q ^ r) & m): squared L2 restricted to
    observed columns of 0/1 vectors (partial_faiss_intersect.py:82-111)."""
    P = pack_bits_u32(panel01)
    Q = pack_bits_u32(queries01)
    nq = Q.shape[0]
    if mask01 is not None:
        mask01 = np.asarray(mask01)
        if mask01.ndim == 1:
            mask01 = np.broadcast_to(mask01, (nq, mask01.shape[0]))
        M = pack_bits_u32(mask01)
    out = np.empty((nq, P.shape[0]), dtype=np.int32)
    for s in range(0, nq, block):
        x = Q[s : s + block, None, :] ^ P[None, :, :]
        if mask01 is not None:
            x &= M[s : s + block, None, :]
        out[s : s + block] = np.bitwise_count(x).sum(axis=2, dtype=np.int32)
    return out


def hamming_topk(panel01, queries01, k, mask01=None):
    """(D int32 [nq,k], I int64 [nq,k]) — what IndexBinaryFlat.search returns
    (test_faiss_intersect.py:164-181) and, as float32, what IndexFlatL2.search returns on
    0/1 vectors (batch_test_faiss_l2.py:110), under the (distance, index) order."""
    d = hamming_matrix(panel01, queries01, mask01)
    return _topk_lex(d, k, I32_MAX)


def hamming_topk_packed(P: np.ndarray, Q: np.ndarray, k: int, M: np.ndarray | None = None, block: int = 256):
    """Same on already packed uint32 rows [n, words]."""
    nq = Q.shape[0]
    out = np.empty((nq, P.shape[0]), dtype=np.int32)
    for s in range(0, nq, block):
        x = Q[s : s + block, None, :] ^ P[None, :, :]
        if M is not None:
            x &= M[s : s + block, None, :]
        out[s : s + block] = np.bitwise_count(x).sum(axis=2, dtype=np.int32)
    return _topk_lex(out, k, I32_MAX)


def token_l2_matrix(panel_tok: np.ndarray, query_tok: np.ndarray) -> np.ndarray:
    """Exact integer squared L2 between token rows (values in {0,2,3,4,5,6}); what
    IndexFlatL2(1030) computes on the V17 path (src/dataset/rag_train_dataset.py:132-134,281).
    All partial sums are integers < 2^24 so fp32 evaluation is exact in any order."""
    p = np.asarray(panel_tok, dtype=np.int64)
    q = np.asarray(query_tok, dtype=np.int64)
    qn = (q * q).sum(1)[:, None]
    pn = (p * p).sum(1)[None, :]
    return (qn + pn - 2 * (q @ p.T)).astype(np.int64)


def token_l2_topk(panel_tok, query_tok, k):
    d = token_l2_matrix(panel_tok, query_tok).astype(np.float32)
    return _topk_lex(d, k, F32_MAX)


def l2_matrix_f64(panel: np.ndarray, queries: np.ndarray) -> np.ndarray:
    """float64 "truth" squared L2 [nq, N]."""
    p = np.asarray(panel, dtype=np.float64)
    q = np.asarray(queries, dtype=np.float64)
    qn = (q * q).sum(1)[:, None]
    pn = (p * p).sum(1)[None, :]
    return np.maximum(qn + pn - 2.0 * (q @ p.T), 0.0)


def l2_topk_f64(panel, queries, k):
    return _topk_lex(l2_matrix_f64(panel, queries), k, np.float64(F32_MAX))


def l2_topk_f32_blas(panel: np.ndarray, queries: np.ndarray, k: int, bq: int = 4096, bn: int = 1024):
    """faiss's BLAS path restated (facebookresearch/faiss, utils/distances.cpp
    ``exhaustive_L2sqr_blas``; version unpinned by the reference): fp32
    ``|x|^2 + |y|^2 - 2 x.y`` in (4096 query) x (1024 database) blocks through sgemm,
    negative results clamped to 0, heap top-k; output ordered by (distance, id)."""
    p = np.ascontiguousarray(panel, dtype=np.float32)
    q = np.ascontiguousarray(queries, dtype=np.float32)
    pn = np.einsum("ij,ij->i", p, p).astype(np.float32)
    qn = np.einsum("ij,ij->i", q, q).astype(np.float32)
    nq, n = q.shape[0], p.shape[0]
    dist = np.empty((nq, n), dtype=np.float32)
    for i0 in range(0, nq, bq):
        for j0 in range(0, n, bn):
            ip = q[i0 : i0 + bq] @ p[j0 : j0 + bn].T
            blk = qn[i0 : i0 + bq, None] + pn[None, j0 : j0 + bn] - np.float32(2.0) * ip
            np.maximum(blk, np.float32(0.0), out=blk)
            dist[i0 : i0 + bq, j0 : j0 + bn] = blk
    return _topk_lex(dist, k, F32_MAX)


def cdist_topk_indices(panel: np.ndarray, queries: np.ndarray, k: int) -> np.ndarray:
    """The V18 train search (src/dataset/embedding_rag_dataset.py:392-402):
    ``torch.cdist(q, r, p=2).topk(k, largest=False)`` — non-squared L2, indices only.
    Restated in float64 with the canonical order (sqrt is monotone, so it equals the
    squared-L2 ranking up to ties)."""
    return l2_topk_f64(panel, queries, k)[1]


def assert_ids_match_within_tolerance(d64: np.ndarray, I_test: np.ndarray, I_ref: np.ndarray, tol_abs) -> int:
    """Float-L2 index parity: the two id lists must select rows whose float64
    distances agree position by position within ``tol_abs`` ([nq] or scalar) — i.e. ids may
    differ only at ties / near-ties inside the stated tolerance.  Returns #positions that differ."""
    tol = np.broadcast_to(np.asarray(tol_abs, dtype=np.float64).reshape(-1, 1) if np.ndim(tol_abs) else tol_abs,
                          I_ref.shape)
    da = np.take_along_axis(d64, np.where(I_test >= 0, I_test, 0), axis=1)
    db = np.take_along_axis(d64, np.where(I_ref >= 0, I_ref, 0), axis=1)
    diff = I_test != I_ref
    bad = diff & (np.abs(da - db) > tol)
    if bad.any():
        q, j = np.argwhere(bad)[0]
        raise AssertionError(
            f"ids differ outside tolerance at query {q} rank {j}: {I_test[q, j]} (d={da[q, j]!r}) vs "
            f"{I_ref[q, j]} (d={db[q, j]!r}), tol={tol[q, j]!r}; {int(bad.sum())} such positions")
    return int(diff.sum())


# --------------------------------------------------------------------------------------
# Gather into the model's input layout
# --------------------------------------------------------------------------------------
def gather_tokens(raw_ref_window: np.ndarray, I: np.ndarray, seq_len: int = MAX_SEQ_LEN) -> np.ndarray:
    """src/dataset/rag_train_dataset.py:287-307: for every (query, j):
    ``s = I//2, h = I%2`` -> ``raw_ref_window[:, s, h]`` -> tokenize with an all-zero mask.
    raw_ref_window [L_w, S, 2]; I [nq, k] -> int64 [nq, k, seq_len] (retrieved rows are
    UNMASKED).  I == -1 rows come back all PAD."""
    I = np.asarray(I)
    nq, k = I.shape
    rows = panel_rows_from_gt(raw_ref_window)  # [2S, L_w]; row 2s+h == raw[:, s, h]
    out = np.zeros((nq, k, seq_len), dtype=np.int64)
    valid = I >= 0
    tok = tokenize(rows[np.where(valid, I, 0).reshape(-1)], None, seq_len).reshape(nq, k, seq_len)
    out[valid] = tok[valid]
    return out


def gather_rows(panel: np.ndarray, I: np.ndarray) -> np.ndarray:
    """Row gather ``panel[I]`` -> [nq, k, d] (src/dataset/embedding_rag_dataset.py:406-438
    reduces to this once the re-embedding is a per-row function); I == -1 -> zeros."""
    I = np.asarray(I)
    out = np.asarray(panel)[np.where(I >= 0, I, 0)]
    out[I < 0] = 0
    return out


# --------------------------------------------------------------------------------------
# k-way merge of per-shard top-k (row-sharded panel; no reference counterpart: the oracle
# for it is "sharded == unsharded")
# --------------------------------------------------------------------------------------
def merge_topk(D_parts, I_parts, k, pad_d):
    """Concatenate per-shard (D, I) with GLOBAL ids along axis 1 and reselect by (D, I)."""
    D = np.concatenate(D_parts, axis=1)
    I = np.concatenate(I_parts, axis=1)
    big = np.where(I < 0, np.iinfo(np.int64).max, I)
    order = np.lexsort((big, D), axis=1)[:, :k]
    Do = np.take_along_axis(D, order, axis=1)
    Io = np.take_along_axis(I, order, axis=1)
    Do = np.where(Io < 0, pad_d, Do)
    return Do, Io


# --------------------------------------------------------------------------------------
# Synthetic haplotypes (SURVEY.md §8d hapgen): shared by tests and bench so that the CPU
# and GPU legs see identical inputs.
# --------------------------------------------------------------------------------------
def hapgen(seed: int, n: int, n_sites: int, founders: int = 64, switch: float = 1 / 200,
           flip: float = 1e-3, iid: bool = False, founder_seed: int | None = None) -> np.ndarray:
    """Mosaic-of-founders haplotypes, uint8 [n, n_sites] in {0,1}.  Per-site alt-allele
    frequency ~ Beta(0.25, 0.75); each haplotype copies a founder, switching founder with
    probability ``switch`` per site and flipping alleles with probability ``flip``.
    ``founder_seed`` lets panel and queries share founders (same population)."""
    frng = np.random.default_rng(seed if founder_seed is None else founder_seed)
    p = frng.beta(0.25, 0.75, size=n_sites)
    F = (frng.random((founders, n_sites)) < p[None, :]).astype(np.uint8)
    rng = np.random.default_rng(seed)
    if iid:
        return (rng.random((n, n_sites)) < p[None, :]).astype(np.uint8)
    sw = rng.random((n, n_sites)) < switch
    sw[:, 0] = True
    seg = np.cumsum(sw, axis=1) - 1  # segment id per site
    nseg = int(seg.max()) + 1
    choice = rng.integers(0, founders, size=(n, nseg))
    fid = np.take_along_axis(choice, seg, axis=1)
    hap = F[fid, np.arange(n_sites)[None, :]]
    hap ^= (rng.random((n, n_sites)) < flip).astype(np.uint8)
    return hap
