"""Device-resident mirror of the V18 embedding-space retrieval used inside the training / inference loop.

Replaces `EmbeddingRAGDataset.process_batch_retrieval` (src/dataset/embedding_rag_dataset.py:285-444) and its
inference twin (src/dataset/embedding_rag_infer_dataset.py:226-330).  Same inputs, same outputs
(`rag_emb_h1`, `rag_emb_h2` float32 [B, k, L, D] carrying gradient into the embedding layer), same order of
embedding-layer calls (so dropout consumes the random stream exactly as in the reference) - but

  * the search is the tcgen05 squared-L2 engine (one `WindowedL2Index` per cached window, `search` on device
    tensors: no .cpu().numpy() round trip, no faiss-gpu, no dense [B, N] distance matrix),
  * unique(I1 u I2) -> re-encode -> scatter is done with static-shape tensor ops (sort, first-occurrence flags,
    prefix sum, one index_select) instead of a Python dict and B*k `int()` reads: the call makes NO host
    synchronisation (tests run it under torch.cuda.set_sync_debug_mode("error")),
  * window search-side panels are cached per window (LRU, `max_cached_windows`), where the reference keeps one.

Semantics kept: the search-side panel of a window is the embedding of its MASKED complete tokens in eval mode
without gradient (:334-377); queries are embedded in the layer's current mode (:385-386) and only their ids are
used; the retrieved rows are the COMPLETE tokens re-encoded WITH gradient, once per distinct row (:406-421), and
duplicates share that one encoding (:427-438).  torch.cdist + topk and this engine agree on the neighbours except
for ties inside the stated L2 tolerance.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional, Sequence

import numpy as np
import torch

from .index import WindowedL2Index


class EmbeddingRagRetriever:
    def __init__(self, ref_tokens_complete: Sequence, ref_af_windows: Sequence, window_masks: Sequence, embed_dim: int,
                 mask_index: int = 4, device: Optional[int] = None, precision: str = "tf32x3", max_cached_windows: int = 1):
        """ref_tokens_complete[w]: int64 [N, L] complete (unmasked) reference tokens of window w; ref_af_windows[w]:
        float32 [L]; window_masks[w]: int [L], 1 = masked position (EmbeddingRAGDataset attributes of the same names)."""
        if not (len(ref_tokens_complete) == len(ref_af_windows) == len(window_masks)) or not len(window_masks):
            raise ValueError("one token matrix, AF vector and mask per window")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", int(device))
        self.embed_dim = int(embed_dim)
        self.mask_index = int(mask_index)
        self.precision = precision
        self.max_cached = max(1, int(max_cached_windows))
        self._tok = [torch.as_tensor(np.asarray(t), dtype=torch.int64) for t in ref_tokens_complete]   # host until first use
        self._af = [torch.as_tensor(np.asarray(a), dtype=torch.float32) for a in ref_af_windows]
        self._mask = [torch.as_tensor(np.asarray(m), dtype=torch.int64) for m in window_masks]
        self._cache: "OrderedDict[int, tuple]" = OrderedDict()  # window -> (index, complete tokens on device, af on device)
        # per-call index lists (which batch rows belong to which window) reach the device from a small ring of PINNED
        # staging rows with asynchronous copies: a pageable copy would block the host
        self._pin = None
        self._pin_slot = 0

    def set_window_mask(self, w: int, mask) -> None:
        """regenerate_masks (embedding_rag_dataset.py:228-283): a new mask invalidates the window's cached panel"""
        self._mask[w] = torch.as_tensor(np.asarray(mask), dtype=torch.int64)
        self._cache.pop(int(w), None)

    def invalidate(self) -> None:
        """drop every cached panel (e.g. after an optimiser step, if the search side should follow the weights)"""
        self._cache.clear()

    # ---- index side (embedding_rag_dataset.py:334-377)
    def _window(self, w: int, embedding_layer):
        hit = self._cache.get(w)
        if hit is not None:
            self._cache.move_to_end(w)
            return hit
        tok = self._tok[w].to(self.device, non_blocking=True)
        af = self._af[w].to(self.device, non_blocking=True)
        mask = self._mask[w].to(self.device, non_blocking=True)
        masked = torch.where(mask.unsqueeze(0) == 1, torch.full_like(tok, self.mask_index), tok)   # :446-461, no boolean indexing
        n, L = tok.shape
        was_training = embedding_layer.training
        embedding_layer.eval()
        with torch.no_grad():
            ref_emb = embedding_layer(masked, af=af.unsqueeze(0).expand(n, -1), pos=True)            # [N, L, D]
        embedding_layer.train(was_training)
        index = WindowedL2Index(L * self.embed_dim, 1, self.device.index, self.precision)
        index.add(ref_emb.reshape(n, L * self.embed_dim).contiguous())
        del ref_emb
        entry = (index, tok, af)
        self._cache[w] = entry
        while len(self._cache) > self.max_cached:
            self._cache.popitem(last=False)
        return entry

    # ---- the batch call (embedding_rag_dataset.py:285-444)
    def process_batch_retrieval(self, batch: dict, embedding_layer, k_retrieve: int = 1) -> dict:
        dev = self.device
        h1_tokens = batch["hap_1"].to(dev, non_blocking=True)
        h2_tokens = batch["hap_2"].to(dev, non_blocking=True)
        af_batch = batch["af"].to(dev, non_blocking=True)
        wins = batch["window_idx"]
        wins = wins.tolist() if hasattr(wins, "tolist") else list(wins)   # host metadata, as in the reference's loop
        groups: "OrderedDict[int, list]" = OrderedDict()
        for i, w in enumerate(wins):
            groups.setdefault(int(w), []).append(i)
        B, L = h1_tokens.shape
        D, k = self.embed_dim, int(k_retrieve)
        # [rows grouped by window | inverse permutation] through one pinned staging row (ring of 4 calls in flight)
        if self._pin is None or self._pin.shape[1] < 2 * B:
            self._pin = torch.empty((4, max(2 * B, 256)), dtype=torch.int64).pin_memory()
        stage = self._pin[self._pin_slot]
        self._pin_slot = (self._pin_slot + 1) % 4
        order = [i for idxs in groups.values() for i in idxs]
        stage[:B] = torch.tensor(order)
        inv = torch.empty(B, dtype=torch.int64)
        inv[stage[:B]] = torch.arange(B)
        stage[B:2 * B] = inv
        staged = stage[:2 * B].to(dev, non_blocking=True)
        order_dev, inv_order = staged[:B], staged[B:]
        outs1, outs2 = [], []
        at = 0
        for w, idxs in groups.items():
            index, ref_tok, ref_af = self._window(w, embedding_layer)
            sel = order_dev[at:at + len(idxs)]
            at += len(idxs)
            h1_win, h2_win, af_win = h1_tokens.index_select(0, sel), h2_tokens.index_select(0, sel), af_batch.index_select(0, sel)
            bw = len(idxs)
            # queries: embedded in the layer's current mode (the random stream advances as in the reference); only ids are used
            with torch.no_grad():
                h1_emb = embedding_layer(h1_win, af=af_win, pos=True)
                h2_emb = embedding_layer(h2_win, af=af_win, pos=True)
                q = torch.cat([h1_emb.reshape(bw, L * D), h2_emb.reshape(bw, L * D)], dim=0).contiguous()
                _, I = index.search(q, k)                                   # [2 bw, k] int64, on the device, no sync
            ids = I.reshape(-1)                                            # h1 rows first, then h2 (torch.cat order, :407)
            m = ids.numel()
            # unique + inverse with static shapes: sort, flag first occurrences, prefix-sum -> slot of every id
            srt, perm = torch.sort(ids, stable=True)
            first = torch.ones(m, dtype=torch.int64, device=dev)
            first[1:] = (srt[1:] != srt[:-1]).to(torch.int64)
            slot_sorted = torch.cumsum(first, 0) - 1                        # slot of each sorted id among the distinct ids
            uniq = torch.zeros(m, dtype=torch.int64, device=dev).scatter_(0, slot_sorted, srt)   # slots >= n_unique keep row 0 (unused)
            inverse = torch.empty(m, dtype=torch.int64, device=dev).scatter_(0, perm, slot_sorted)
            # re-encode the COMPLETE rows once per distinct id, WITH gradient (:409-421)
            retrieved_tokens = ref_tok.index_select(0, uniq)
            retrieved_emb = embedding_layer(retrieved_tokens, af=ref_af.unsqueeze(0).expand(m, -1), pos=True)   # [m, L, D]
            gathered = retrieved_emb.index_select(0, inverse).reshape(2, bw, k, L, D)
            outs1.append(gathered[0])
            outs2.append(gathered[1])
        batch["rag_emb_h1"] = torch.cat(outs1, dim=0).index_select(0, inv_order).contiguous()   # [B, k, L, D], caller order
        batch["rag_emb_h2"] = torch.cat(outs2, dim=0).index_select(0, inv_order).contiguous()
        return batch
