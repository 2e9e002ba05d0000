"""Device-resident mirror of the V17 retrieval in the reference's collate.

Replaces, for one training / inference batch,
  * the per-window faiss indexes built by RAGTrainDataset._build_faiss_indexes
    (src/dataset/rag_train_dataset.py:41-139; infer twin rag_infer_dataset.py:38-114), and
  * the search + gather + re-tokenise loop of rag_collate_fn_with_dataset (:232-307),
with one grouped Hamming search and one gather launch on the GPU (no Python loop per query, no
`vocab.to_seq` per retrieved row).  Call it from the process that owns the CUDA context
(`num_workers=0` or the trainer process, the V18 pattern) — never from forked workers.

Semantics: panel rows are tokenised with the window's mask and queries with the same mask (static
mask, the reference default `use_dynamic_mask=False`), so squared L2 over tokens == Hamming over
the observed sites (pinned by tests/golden/g2).  With per-query masks that DIFFER from the panel's
(dynamic masks) the reference's token-space costs are (MASK,0)->1, (MASK,1)->4; use
faiss_compat.IndexFlatL2 (exact on token vectors) for that case.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from .index import WindowedHammingIndex, _is_torch

SOS, EOS, MASK, ALLELE0, ALLELE1 = 2, 3, 4, 5, 6
MAX_SEQ_LEN = 1030


class RagRetriever:
    def __init__(self, raw_ref_windows: Sequence[np.ndarray], device: Optional[int] = None,
                 seq_len: int = MAX_SEQ_LEN):
        """raw_ref_windows[w]: genotype cube [L_w, S, 2] (dataset.raw_ref_data_windows,
        rag_train_dataset.py:111-112).  Panel row ids follow the reference: 2*sample + hap."""
        if not len(raw_ref_windows):
            raise ValueError("no windows")
        self.seq_len = int(seq_len)
        self.n_sites = np.array([w.shape[0] for w in raw_ref_windows], dtype=np.int32)
        if int(self.n_sites.max()) > self.seq_len - 2:
            raise ValueError("a window has more sites than seq_len - 2")
        n_rows = {w.shape[1] * w.shape[2] for w in raw_ref_windows}
        if len(n_rows) != 1:
            raise ValueError("all windows must hold the same reference samples")
        self.d = self.seq_len - 2
        W, N = len(raw_ref_windows), n_rows.pop()
        rows = np.zeros((W, N, self.d), dtype=np.uint8)
        for w, cube in enumerate(raw_ref_windows):
            lw = cube.shape[0]
            rows[w, :, :lw] = (np.asarray(cube).reshape(lw, -1).T != 0)  # :115-118
        self.index = WindowedHammingIndex(self.d, W, device)
        self.index.add(rows)

    def _queries(self, hap_1, hap_2):
        """tokens [B, L] x2 -> interleaved [2B, d] allele / observed planes (queries.extend([h1, h2]), :262-276)."""
        if _is_torch(hap_1):
            import torch

            tok = torch.stack([hap_1, hap_2], dim=1).reshape(-1, hap_1.shape[-1])[:, 1:1 + self.d]
            return (tok == ALLELE1).to(torch.uint8).contiguous(), ((tok == ALLELE0) | (tok == ALLELE1)).to(torch.uint8).contiguous()
        tok = np.stack([np.asarray(hap_1), np.asarray(hap_2)], axis=1).reshape(-1, np.asarray(hap_1).shape[-1])[:, 1:1 + self.d]
        return (tok == ALLELE1).astype(np.uint8), ((tok == ALLELE0) | (tok == ALLELE1)).astype(np.uint8)

    def search(self, window_idx, hap_1, hap_2, k: int):
        """-> D float32 [B, 2, k] (== the reference's faiss D), I int64 [B, 2, k]."""
        q, obs = self._queries(hap_1, hap_2)
        wid = np.repeat(np.asarray(window_idx.cpu() if _is_torch(window_idx) else window_idx, dtype=np.int32), 2)
        D, I = self.index.search_grouped(q, wid, k, observed=obs, dist_dtype=np.float32)
        B = wid.shape[0] // 2
        return D.reshape(B, 2, k), I.reshape(B, 2, k)

    def retrieve(self, window_idx, hap_1, hap_2, k: int):
        """-> rag_seg_h1, rag_seg_h2 int64 [B, k, seq_len] in CALLER order (the reference returns them
        regrouped by window; tokens per row are identical): retrieved haplotypes tokenised UNMASKED
        (rag_train_dataset.py:254-255,303)."""
        D, I = self.search(window_idx, hap_1, hap_2, k)
        B = I.shape[0]
        wid = np.repeat(np.asarray(window_idx.cpu() if _is_torch(window_idx) else window_idx, dtype=np.int32), 2)
        seg = self.index.gather_tokens_grouped(I.reshape(2 * B, k), wid, n_sites=self.n_sites, seq_len=self.seq_len)
        seg = seg.reshape(B, 2, k, self.seq_len)
        return seg[:, 0], seg[:, 1]
