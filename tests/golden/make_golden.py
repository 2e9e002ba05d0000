#!/usr/bin/env python
"""Generates tests/golden/*.npz by IMPORTING THE REFERENCE'S OWN PYTHON from /root/reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference cannot be imported as-is here: faiss, h5py, allel, vcfpy and matplotlib are absent
(SURVEY.md §8c).  They are replaced by stub modules below; only `faiss` is ever *called*, and
the stub implements exactly the documented contract of the two calls the reference makes
(IndexFlatL2.add / .search: exact squared L2 in float32, results ordered by (distance, id)),
in a few lines of numpy — so what the fixtures pin is the REFERENCE-OWNED logic around it:

  g1  tokenisation          TrainDataset.tokenize + WordVocab.to_seq + sequence_padding
                            (src/dataset/dataset.py:597-625, vocab.py:153-170, utils.py:121-132)
  g2  V17 panel layout,     the index-side statements of RAGTrainDataset._build_faiss_indexes
      search + gather       (rag_train_dataset.py:111-134) followed by the real
                            rag_collate_fn_with_dataset (:232-358) -> rag_seg_h1 / rag_seg_h2
  g3  binary packing        bitpack_2d_array (test_faiss_intersect.py:46-54)
  g4  V18 embedding search  EmbeddingRAGDataset.process_batch_retrieval
                            (embedding_rag_dataset.py:285-444) with the real BERTEmbedding:
                            torch.cdist + topk ids and the gathered rag_emb tensors
  g5  observed-site search  expand_target_to_ref + build_partial_index_l2 (partial_faiss_intersect.py:46-111),
                            see make_g5()
  g6  offline DB workflow   build_ref_db_l2(args) + batch_test_faiss_l2(args) run whole, see make_g6()
  g7  V18 inference search  EmbeddingRAGInferDataset.process_batch_retrieval (faiss flat index per window), see make_g7()
  g8  intersect workflow    build_ref_db_intersect(args) + test_faiss_intersect(args) in both distance modes, see make_g8()
  g9  V18 training retrieval EmbeddingRAGDataset.process_batch_retrieval over a batch that interleaves two windows, with
                            the GRADIENTS of a loss on rag_emb_h1 / rag_emb_h2 w.r.t. every embedding parameter, see make_g9()
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


# ------------------------------------------------------------------------------- stubs
class _ShimIndexFlatL2:
    """The faiss.IndexFlatL2 contract in numpy (float32 exact squared L2, (distance, id) order)."""

    def __init__(self, d):
        self.d = int(d)
        self.xb = np.zeros((0, self.d), np.float32)

    @property
    def ntotal(self):
        return self.xb.shape[0]

    def add(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d
        self.xb = np.concatenate([self.xb, x])

    def search(self, x, k):
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d
        diff = x[:, None, :] - self.xb[None, :, :]
        dist = np.einsum("qnd,qnd->qn", diff, diff).astype(np.float32)
        order = np.argsort(dist, axis=1, kind="stable")[:, :k]
        return np.take_along_axis(dist, order, 1), order.astype(np.int64)


def install_stubs():
    faiss = types.ModuleType("faiss")
    faiss.IndexFlatL2 = _ShimIndexFlatL2
    faiss.StandardGpuResources = lambda: None
    sys.modules["faiss"] = faiss
    for name in ("h5py", "allel", "matplotlib", "matplotlib.pyplot", "vcfpy", "seaborn"):
        m = types.ModuleType(name)
        sys.modules[name] = m
    sys.modules["vcfpy"].Header = object
    sys.modules["vcfpy"].Reader = object
    sys.modules["vcfpy"].Writer = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].use = lambda *a, **k: None


def main():
    install_stubs()
    sys.path.insert(0, REF)
    os.chdir("/tmp")  # PanelData writes POP.json into the CWD; keep the repo clean
    import torch

    from src.dataset.dataset import TrainDataset
    from src.dataset.utils import VCFProcessingModule
    from src.dataset.vocab import WordVocab
    from src.dataset import rag_train_dataset as RTD
    from src.dataset.embedding_rag_dataset import EmbeddingRAGDataset
    from src.model.embedding.bert import BERTEmbedding

    rng = np.random.default_rng(20261018)
    vocab = WordVocab(["AFR", "EUR", "EAS"])
    holder = types.SimpleNamespace(vocab=vocab)
    tokenize = lambda seq, mask=None: TrainDataset.tokenize(holder, seq, mask)  # noqa: E731

    # ---- g1 tokenisation ------------------------------------------------------------------
    g1 = {}
    for name, lw in (("a", 1000), ("b", 1028), ("c", 37)):
        seq = (rng.random((6, lw)) < 0.3).astype(np.int64)
        raw_mask = (rng.random(lw) < 0.25).astype(np.int64)
        padded = VCFProcessingModule.sequence_padding(raw_mask, dtype="int")
        g1[f"seq_{name}"] = seq
        g1[f"rawmask_{name}"] = raw_mask
        g1[f"padmask_{name}"] = padded
        g1[f"tok_masked_{name}"] = tokenize(seq, padded)
        g1[f"tok_plain_{name}"] = tokenize(seq, np.zeros_like(padded))
    g1["specials"] = np.array([vocab.pad_index, vocab.unk_index, vocab.sos_index, vocab.eos_index,
                               vocab.mask_index, vocab.stoi[0], vocab.stoi[1]])
    np.savez_compressed(os.path.join(OUT, "g1_tokenize.npz"), **g1)

    # ---- g2 V17: panel layout -> index -> collate search + gather ----------------------------
    lw, S_ref, n_samples, k = 1000, 40, 6, 3
    founders = (rng.random((8, lw)) < 0.3).astype(np.int8)

    def mosaic(n):
        out = np.empty((n, lw), np.int8)
        for i in range(n):
            cuts = np.sort(rng.choice(lw, 4, replace=False))
            f = rng.integers(0, 8, 5)
            seg = np.searchsorted(cuts, np.arange(lw), side="right")
            out[i] = founders[f[seg], np.arange(lw)]
            out[i] ^= (rng.random(lw) < 0.01).astype(np.int8)
        return out

    windows = []
    ds = types.SimpleNamespace(window_indexes=[], raw_ref_data_windows=[], raw_window_masks=[],
                               ref_data_windows=[], vocab=vocab)
    ds.tokenize = tokenize
    for w in range(2):
        ref_gt = mosaic(2 * S_ref).reshape(S_ref, 2, lw).transpose(2, 0, 1).copy()  # [L_w, S, 2]
        raw_mask = (rng.random(lw) < 0.3).astype(np.int64)
        padded_mask = VCFProcessingModule.sequence_padding(raw_mask, dtype="int")
        # rag_train_dataset.py:111-134, statement for statement
        raw_ref = ref_gt[slice(0, lw), :, :]
        ds.raw_ref_data_windows.append(raw_ref)
        raw_ref2 = raw_ref.reshape(raw_ref.shape[0], -1)
        raw_ref2 = raw_ref2.T
        ref_tokenized = tokenize(raw_ref2, padded_mask)
        ds.ref_data_windows.append(ref_tokenized)
        index_data = ref_tokenized.astype(np.float32)
        index = sys.modules["faiss"].IndexFlatL2(index_data.shape[1])
        index.add(index_data)
        ds.window_indexes.append(index)
        ds.raw_window_masks.append(raw_mask)
        windows.append((ref_gt, raw_mask, padded_mask, ref_tokenized))
    batch = []
    q_h1, q_h2, q_win = [], [], []
    for i in range(n_samples):
        w = i % 2
        padded_mask = windows[w][2]
        h = mosaic(2)
        if i == 0:  # planted exact match with panel row 2*7+1 (test_R_only.py:32-55 invariant)
            h[0] = windows[w][0][:, 7, 1]
        hap1 = tokenize(h[0].astype(np.int64), padded_mask)
        hap2 = tokenize(h[1].astype(np.int64), padded_mask)
        q_h1.append(hap1)
        q_h2.append(hap2)
        q_win.append(w)
        batch.append({"window_idx": w, "hap_1": torch.from_numpy(hap1), "hap_2": torch.from_numpy(hap2)})
    # record the (D, I) the collate's index.search produced
    rec = []
    orig_search = _ShimIndexFlatL2.search

    def rec_search(self, x, k):
        D, I = orig_search(self, x, k)
        rec.append((np.array(x), D, I))
        return D, I

    _ShimIndexFlatL2.search = rec_search
    out = RTD.rag_collate_fn_with_dataset(batch, ds, k)
    _ShimIndexFlatL2.search = orig_search
    g2 = {"k": np.array(k), "lw": np.array(lw)}
    for w in range(2):
        g2[f"ref_gt_{w}"] = windows[w][0]
        g2[f"raw_mask_{w}"] = windows[w][1]
        g2[f"padded_mask_{w}"] = windows[w][2]
        g2[f"ref_tokenized_{w}"] = windows[w][3]
        g2[f"search_q_{w}"] = rec[w][0]
        g2[f"search_D_{w}"] = rec[w][1]
        g2[f"search_I_{w}"] = rec[w][2]
    g2["window_idx"] = np.array(out["window_idx"])
    g2["hap_1"] = out["hap_1"].numpy()
    g2["hap_2"] = out["hap_2"].numpy()
    g2["rag_seg_h1"] = out["rag_seg_h1"].numpy()
    g2["rag_seg_h2"] = out["rag_seg_h2"].numpy()
    np.savez_compressed(os.path.join(OUT, "g2_v17_collate.npz"), **g2)

    # ---- g3 bit packing ----------------------------------------------------------------------
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_tfi", os.path.join(REF, "test_faiss_intersect.py"))
    tfi = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tfi)
    arr = (rng.random((5, 2 * 517)) < 0.5).astype(np.int8)
    np.savez_compressed(os.path.join(OUT, "g3_bitpack.npz"), arr=arr, packed=tfi.bitpack_2d_array(arr))

    # ---- g4 V18 embedding-space retrieval ------------------------------------------------------
    torch.manual_seed(7)
    L, D, N, B, k18 = 96, 16, 48, 10, 2
    emb = BERTEmbedding(vocab_size=len(vocab), embed_size=D, dropout=0.0, use_af=True)
    emb.train()  # the trainer calls it in train mode; dropout=0 keeps it deterministic
    hap = (rng.random((N, L - 2)) < 0.3).astype(np.int64)
    ref_tokens_complete = tokenize(hap, None)[:, :L] if False else None
    # tokens of length L: [SOS] + L-2 sites + [EOS]
    ref_tokens_complete = np.concatenate(
        [np.full((N, 1), 2), np.where(hap == 0, 5, 6), np.full((N, 1), 3)], axis=1).astype(np.int64)
    win_mask = np.zeros(L, np.int64)
    win_mask[1:-1] = (rng.random(L - 2) < 0.3)
    ref_af = rng.random(L).astype(np.float32)
    qhap = hap[rng.integers(0, N, 2 * B)].copy()
    qhap ^= (rng.random(qhap.shape) < 0.05)
    q_tok = np.concatenate([np.full((2 * B, 1), 2), np.where(qhap == 0, 5, 6), np.full((2 * B, 1), 3)], axis=1).astype(np.int64)
    q_tok[:, win_mask == 1] = vocab.mask_index
    fake = types.SimpleNamespace(
        embed_dim=D, jit_cache_win_idx=-1, jit_ref_emb_search=None, jit_ref_tokens_raw=None, jit_ref_af_raw=None,
        ref_tokens_complete=[ref_tokens_complete], ref_af_windows=[ref_af], window_masks=[win_mask], vocab=vocab)
    fake._apply_mask_to_tokens_gpu = lambda t, m: EmbeddingRAGDataset._apply_mask_to_tokens_gpu(fake, t, m)
    batch = {"hap_1": torch.from_numpy(q_tok[:B]), "hap_2": torch.from_numpy(q_tok[B:]),
             "af": torch.from_numpy(np.tile(ref_af, (B, 1))), "window_idx": [0] * B}
    topk_rec = []
    orig_topk = torch.Tensor.topk

    def rec_topk(self, *a, **kw):
        r = orig_topk(self, *a, **kw)
        topk_rec.append((self.detach().numpy().copy(), r[1].numpy().copy()))
        return r

    torch.Tensor.topk = rec_topk
    out = EmbeddingRAGDataset.process_batch_retrieval(fake, batch, emb, "cpu", k18)
    torch.Tensor.topk = orig_topk
    with torch.no_grad():
        emb.eval()
        ref_masked = ref_tokens_complete.copy()
        ref_masked[:, win_mask == 1] = vocab.mask_index
        af_t = torch.from_numpy(ref_af)
        ref_emb_search = emb(torch.from_numpy(ref_masked), af=af_t.unsqueeze(0).expand(N, -1), pos=True)
        ref_emb_complete = emb(torch.from_numpy(ref_tokens_complete), af=af_t.unsqueeze(0).expand(N, -1), pos=True)
        q1 = emb(batch["hap_1"], af=batch["af"], pos=True)
        q2 = emb(batch["hap_2"], af=batch["af"], pos=True)
    np.savez_compressed(
        os.path.join(OUT, "g4_v18_embedding.npz"),
        ref_flat=ref_emb_search.reshape(N, L * D).numpy(), ref_complete=ref_emb_complete.numpy(),
        q1_flat=q1.reshape(B, L * D).numpy(), q2_flat=q2.reshape(B, L * D).numpy(),
        dists_h1=topk_rec[0][0], I1=topk_rec[0][1], dists_h2=topk_rec[1][0], I2=topk_rec[1][1],
        rag_emb_h1=out["rag_emb_h1"].detach().numpy(), rag_emb_h2=out["rag_emb_h2"].detach().numpy(),
        k=np.array(k18))
    make_g5()
    make_g6()
    make_g7()
    make_g8()
    make_g9()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


def make_g5():
    """g5  observed-site search   expand_target_to_ref + build_partial_index_l2 of partial_faiss_intersect.py
                                  (:46-80, :82-111), called per (window, target sample) as its main() does (:145-172).

    build_partial_index_l2 lays the query out as [h1 sites.., h2 sites..] (:94) but the panel rows site-major
    (s0h0, s0h1, s1h0, ..; :101) - a layout slip, so its literal output compares misaligned columns.  The fixture keeps
    BOTH: `*_literal` = the function called exactly as main() calls it, and `*_aligned` = the same function given a
    query whose [h1.., h2..] concatenation equals the site-major interleave (the evident intent, and what the engine
    implements): everything else - valid-column selection from the mask, the per-sample panel slicing loop, the
    index build and search - is the reference's own code in both."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_pfi", os.path.join(REF, "partial_faiss_intersect.py"))
    pfi = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pfi)

    rng = np.random.default_rng(55)
    var_ref, s_ref, s_tgt, k = 260, 30, 5, 4
    ref_pos = np.sort(rng.choice(np.arange(1000, 5000), var_ref, replace=False)).astype(np.int64)
    # target: a subset of the ref sites (in order) plus a few positions the ref does not have
    keep = np.sort(rng.choice(var_ref, 180, replace=False))
    extra = np.setdiff1d(rng.choice(np.arange(1000, 5000), 40, replace=False), ref_pos)
    tgt_pos = np.sort(np.concatenate([ref_pos[keep], extra])).astype(np.int64)
    founders = (rng.random((6, var_ref)) < 0.35).astype(np.uint8)
    ref_hap = founders[rng.integers(0, 6, 2 * s_ref)] ^ (rng.random((2 * s_ref, var_ref)) < 0.03)
    ref_data = ref_hap.reshape(s_ref, 2, var_ref).transpose(2, 0, 1).astype(np.uint8).copy()        # [var_ref, S, 2]
    tgt_full = founders[rng.integers(0, 6, 2 * s_tgt)] ^ (rng.random((2 * s_tgt, var_ref)) < 0.03)
    tgt_full = tgt_full.reshape(s_tgt, 2, var_ref).transpose(2, 0, 1).astype(np.uint8)
    pos_to_ref = {int(p): i for i, p in enumerate(ref_pos)}
    tgt_data = np.zeros((tgt_pos.size, s_tgt, 2), np.uint8)
    for t, p_ in enumerate(tgt_pos):
        if int(p_) in pos_to_ref:
            tgt_data[t] = tgt_full[pos_to_ref[int(p_)]]
        else:
            tgt_data[t] = rng.integers(0, 2, (s_tgt, 2))
    expanded, missing_ref = pfi.expand_target_to_ref(ref_pos, tgt_data, tgt_pos)
    missing = missing_ref.copy()  # the masks the searches below use (differs only where a window would have no site left)
    windows = np.array([[0, 70], [70, 170], [170, 260], [40, 41]], np.int64)
    I_lit = np.zeros((len(windows), s_tgt, k), np.int64)
    D_lit = np.zeros((len(windows), s_tgt, k), np.float32)
    I_al = np.zeros_like(I_lit)
    D_al = np.zeros_like(D_lit)
    for w, (a, b) in enumerate(windows):
        ref_sub = np.transpose(ref_data[a:b], (1, 0, 2))  # (samp_ref, w_len, 2), :152-153
        for s in range(s_tgt):
            sub_tgt = expanded[a:b, s, :]
            hap_1, hap_2 = sub_tgt[:, 0], sub_tgt[:, 1]
            sub_mask = missing[a:b, s].copy()
            if (sub_mask == 0).sum() == 0:
                sub_mask[0] = 0  # keep at least one observed site: IndexFlatL2(0) is not a case main() survives
                missing[a, s] = 0
            kk = min(k, ref_sub.shape[0])
            I, D, _, _ = pfi.build_partial_index_l2(hap_1, hap_2, sub_mask, ref_sub, top_k=kk)
            I_lit[w, s], D_lit[w, s] = I, D
            valid = np.where(sub_mask == 0)[0]
            z = np.stack([hap_1[valid], hap_2[valid]], axis=1).reshape(-1)  # site-major interleave
            h1p, h2p = hap_1.copy(), hap_2.copy()
            h1p[valid], h2p[valid] = z[: valid.size], z[valid.size:]
            I, D, _, _ = pfi.build_partial_index_l2(h1p, h2p, sub_mask, ref_sub, top_k=kk)
            I_al[w, s], D_al[w, s] = I, D
    np.savez_compressed(os.path.join(OUT, "g5_partial_intersect.npz"), ref_pos=ref_pos, tgt_pos=tgt_pos, tgt_data=tgt_data,
                        ref_data=ref_data, expanded=expanded, missing_ref=missing_ref, missing=missing, windows=windows, k=np.array(k),
                        I_literal=I_lit, D_literal=D_lit, I_aligned=I_al, D_aligned=D_al)


def make_g6():
    """g6  offline DB workflow   build_ref_db_l2(args) (build_ref_db_l2.py:15-99) followed by batch_test_faiss_l2(args)
                                  (batch_test_faiss_l2.py:48-136), both run WHOLE on synthetic files: a panel file and a
                                  window csv read by the reference's PanelData / Window, genotypes served through an
                                  h5py stub.  Pins the `window_{i}.npy` contents, the sample-row vector layout
                                  (s0h0, s0h1, s1h0, ..) and the (D, I) of every window's batched search."""
    import argparse
    import importlib.util
    import pickle
    import shutil
    import tempfile

    rng = np.random.default_rng(66)
    V, s_ref, s_tgt, k = 400, 36, 7, 5
    founders = (rng.random((8, V)) < 0.3)
    ref_gt = (founders[rng.integers(0, 8, 2 * s_ref)] ^ (rng.random((2 * s_ref, V)) < 0.02)).reshape(s_ref, 2, V).transpose(2, 0, 1)
    tgt_gt = (founders[rng.integers(0, 8, 2 * s_tgt)] ^ (rng.random((2 * s_tgt, V)) < 0.02)).reshape(s_tgt, 2, V).transpose(2, 0, 1)
    # multi-allelic codes > 0 are folded to 1 by the scripts (build_ref_db_l2.py:50, batch_test_faiss_l2.py:45)
    ref_raw = ref_gt.astype(np.int8) * rng.integers(1, 3, ref_gt.shape).astype(np.int8)
    tgt_raw = tgt_gt.astype(np.int8) * rng.integers(1, 3, tgt_gt.shape).astype(np.int8)
    tgt_raw[3, 2] = ref_raw[3, 11]  # a planted near-copy is not needed: founders give ties already
    pos = np.sort(rng.choice(np.arange(10_000, 90_000), V, replace=False)).astype(np.int64)
    windows = np.array([[0, 130], [130, 250], [250, 400]], np.int64)

    store = {}

    class _Dataset:
        def __init__(self, a):
            self.a = a

        def __getitem__(self, _):
            return self.a.copy()

    class _File:
        def __init__(self, path, mode="r"):
            self.d = store[str(path)]

        def __getitem__(self, key):
            return _Dataset(self.d[key])

        def close(self):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    sys.modules["h5py"].File = _File
    faiss = sys.modules["faiss"]
    faiss.write_index = lambda index, path: pickle.dump(index, open(path, "wb"))
    faiss.read_index = lambda path: pickle.load(open(path, "rb"))

    tmp = tempfile.mkdtemp(prefix="g6_")
    try:
        store[os.path.join(tmp, "ref.h5")] = {"calldata/GT": ref_raw, "variants/POS": pos}
        store[os.path.join(tmp, "tgt.h5")] = {"calldata/GT": tgt_raw, "variants/POS": pos}
        pops = ["AFR", "EUR", "EAS"]
        with open(os.path.join(tmp, "ref.panel"), "w") as f:
            f.write("sample\tpop\tsuper_pop\tgender\n")  # PanelData.from_file drops the header line (dataset.py:84-85)
            for i in range(s_ref):
                f.write(f"S{i}\tP{i % 5}\t{pops[i % 3]}\tmale\n")
        with open(os.path.join(tmp, "win.csv"), "w") as f:
            f.write("start,end\n" + "".join(f"{a},{b}\n" for a, b in windows))

        def load(name):
            spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
            m = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(m)
            return m

        db = os.path.join(tmp, "db")
        load("build_ref_db_l2").build_ref_db_l2(argparse.Namespace(
            ref_vcf=os.path.join(tmp, "ref.h5"), ref_panel=os.path.join(tmp, "ref.panel"),
            window_csv=os.path.join(tmp, "win.csv"), output_dir=db))
        rec = []
        orig_search = _ShimIndexFlatL2.search

        def rec_search(self, x, kk):
            D, I = orig_search(self, x, kk)
            rec.append((np.array(x), D, I))
            return D, I

        _ShimIndexFlatL2.search = rec_search
        load("batch_test_faiss_l2").batch_test_faiss_l2(argparse.Namespace(
            target_vcf=os.path.join(tmp, "tgt.h5"), window_csv=os.path.join(tmp, "win.csv"), ref_db=db, sample_idx=-1,
            top_k=k, print_snippet=False, show_snp_len=10))
        _ShimIndexFlatL2.search = orig_search
        out = {"ref_raw": ref_raw, "tgt_raw": tgt_raw, "pos": pos, "windows": windows, "k": np.array(k)}
        for w in range(len(windows)):
            out[f"window_{w}"] = np.load(os.path.join(db, f"window_{w}.npy"))
            out[f"pop_{w}"] = np.load(os.path.join(db, f"window_{w}_pop.npy"))
            out[f"index_xb_{w}"] = pickle.load(open(os.path.join(db, f"window_{w}.faiss"), "rb")).xb
            out[f"query_{w}"], out[f"D_{w}"], out[f"I_{w}"] = rec[w]
        np.savez_compressed(os.path.join(OUT, "g6_ref_db_workflow.npz"), **out)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def make_g7():
    """g7  V18 inference search   EmbeddingRAGInferDataset.process_batch_retrieval
                                  (src/dataset/embedding_rag_infer_dataset.py:250-324) with the real BERTEmbedding: per
                                  window group, faiss.IndexFlatL2(L*D).search of the flattened query embeddings against
                                  the window's flattened masked-panel embeddings (built as :150-181 do), then the
                                  unique-id re-embedding gather into rag_emb_h1 / rag_emb_h2.  Two windows, interleaved
                                  batch: exercises the window grouping too."""
    import torch

    from src.dataset.embedding_rag_infer_dataset import EmbeddingRAGInferDataset
    from src.dataset.vocab import WordVocab
    from src.model.embedding.bert import BERTEmbedding

    rng = np.random.default_rng(77)
    torch.manual_seed(77)
    vocab = WordVocab(["AFR", "EUR", "EAS"])
    L, D, N, B, k = 80, 16, 40, 8, 2
    emb = BERTEmbedding(vocab_size=len(vocab), embed_size=D, dropout=0.0, use_af=True)
    emb.eval()
    W = 2
    ref_tokens_complete, ref_af, masks, flat_search, recs = [], [], [], [], []
    for w in range(W):
        hap = (rng.random((N, L - 2)) < 0.3).astype(np.int64)
        tok = np.concatenate([np.full((N, 1), 2), np.where(hap == 0, 5, 6), np.full((N, 1), 3)], axis=1).astype(np.int64)
        m = np.zeros(L, np.int64)
        m[1:-1] = rng.random(L - 2) < 0.3
        af = rng.random(L).astype(np.float32)
        masked = tok.copy()
        masked[:, m == 1] = vocab.mask_index
        with torch.no_grad():
            e = emb(torch.from_numpy(masked), af=torch.from_numpy(af).unsqueeze(0).expand(N, -1), pos=True)
        ref_tokens_complete.append(tok)
        ref_af.append(af)
        masks.append(m)
        flat_search.append(e.reshape(N, L * D).numpy().astype(np.float32))
    indexes = []
    for w in range(W):
        ix = _ShimIndexFlatL2(L * D)   # :176-177
        ix.add(flat_search[w])
        indexes.append(ix)
    win_of = [0, 1, 1, 0, 1, 0, 0, 1]
    q_tok = np.zeros((2 * B, L), np.int64)
    for i in range(2 * B):
        w = win_of[i % B]
        src = ref_tokens_complete[w][rng.integers(0, N)].copy()
        flip = np.zeros(L, bool)
        flip[1:-1] = rng.random(L - 2) < 0.05
        src[flip] = np.where(src[flip] == 5, 6, 5)
        src[masks[w] == 1] = vocab.mask_index
        q_tok[i] = src
    af_batch = np.stack([ref_af[w] for w in win_of])
    fake = types.SimpleNamespace(embed_dim=D, ref_tokens_complete=ref_tokens_complete, ref_af_windows=ref_af,
                                 load_index=lambda w: indexes[w])
    orig_search = _ShimIndexFlatL2.search

    def rec_search(self, x, k):
        Dd, I = orig_search(self, x, k)
        recs.append((np.array(x), Dd, I))
        return Dd, I

    _ShimIndexFlatL2.search = rec_search
    batch = {"hap_1": torch.from_numpy(q_tok[:B]), "hap_2": torch.from_numpy(q_tok[B:]),
             "af": torch.from_numpy(af_batch), "window_idx": torch.tensor(win_of)}
    out = EmbeddingRAGInferDataset.process_batch_retrieval(fake, batch, emb, "cpu", k_retrieve=k)
    _ShimIndexFlatL2.search = orig_search
    # searches were issued per window group in first-appearance order (0 then 1), h1 then h2
    groups = {w: [i for i, x in enumerate(win_of) if x == w] for w in (0, 1)}
    g7 = {"k": np.array(k), "L": np.array(L), "D": np.array(D), "window_idx": np.array(win_of)}
    r = 0
    for w in (0, 1):
        with torch.no_grad():
            comp = emb(torch.from_numpy(ref_tokens_complete[w]),
                       af=torch.from_numpy(ref_af[w]).unsqueeze(0).expand(N, -1), pos=True)
        g7[f"ref_flat_{w}"] = flat_search[w]
        g7[f"ref_complete_{w}"] = comp.numpy()
        g7[f"members_{w}"] = np.array(groups[w])
        for h in (1, 2):
            g7[f"q{h}_flat_{w}"], g7[f"D{h}_{w}"], g7[f"I{h}_{w}"] = recs[r]
            r += 1
    g7["rag_emb_h1"] = out["rag_emb_h1"].numpy()
    g7["rag_emb_h2"] = out["rag_emb_h2"].numpy()
    np.savez_compressed(os.path.join(OUT, "g7_v18_infer.npz"), **g7)


class _ShimIndexBinaryFlat:
    """faiss.IndexBinaryFlat in numpy: Hamming distance between np.packbits codes, int32 D, (distance, id) order."""

    def __init__(self, d_bits):
        self.d = int(d_bits)
        self.codes = np.zeros((0, (self.d + 7) // 8), np.uint8)

    @property
    def ntotal(self):
        return self.codes.shape[0]

    def add(self, x):
        x = np.ascontiguousarray(x, dtype=np.uint8)
        assert x.ndim == 2 and x.shape[1] == self.codes.shape[1]
        self.codes = np.concatenate([self.codes, x])

    def search(self, x, k):
        x = np.ascontiguousarray(x, dtype=np.uint8)
        dist = np.unpackbits(x[:, None, :] ^ self.codes[None, :, :], axis=2).sum(2).astype(np.int32)
        order = np.argsort(dist, axis=1, kind="stable")[:, :k]
        return np.take_along_axis(dist, order, 1), order.astype(np.int64)


def make_g8():
    """g8  intersect workflow    build_ref_db_intersect(args) (build_ref_db_intersect.py:14-81) followed by
                                  test_faiss_intersect(args) (test_faiss_intersect.py:57-203) in BOTH distance modes
                                  ('l2': IndexFlatL2 on the shared sites; 'binary': bitpack_2d_array + IndexBinaryFlat),
                                  run whole on synthetic files.  The target carries the reference's variant list except
                                  that a quarter of its positions differ, so every window's intersection is partial
                                  (the script slices the target with the same index window as the reference).
                                  Pins window_{i}_pos.npy, the per-window shared-site sets and both modes' (D, I)."""
    import argparse
    import importlib.util
    import shutil
    import tempfile

    rng = np.random.default_rng(88)
    V, s_ref, s_tgt, k = 360, 40, 6, 4
    founders = (rng.random((8, V)) < 0.3)
    ref_gt = (founders[rng.integers(0, 8, 2 * s_ref)] ^ (rng.random((2 * s_ref, V)) < 0.02)).reshape(s_ref, 2, V).transpose(2, 0, 1).astype(np.int8)
    tgt_gt = (founders[rng.integers(0, 8, 2 * s_tgt)] ^ (rng.random((2 * s_tgt, V)) < 0.02)).reshape(s_tgt, 2, V).transpose(2, 0, 1).astype(np.int8)
    ref_pos = (1000 + 5 * np.arange(V) + rng.integers(0, 2, V)).astype(np.int64)     # gaps >= 4
    tgt_pos = ref_pos.copy()
    moved = rng.random(V) < 0.25
    tgt_pos[moved] += 2                                                                # never equals another ref position
    windows = np.array([[0, 100], [100, 232], [232, 360]], np.int64)                   # 2 * shared sites happens to vary mod 8
    store = {}

    class _Dataset:
        def __init__(self, a):
            self.a = a

        def __getitem__(self, _):
            return self.a.copy()

    class _File:
        def __init__(self, path, mode="r"):
            self.d = store[str(path)]

        def __getitem__(self, key):
            return _Dataset(self.d[key])

        def close(self):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    sys.modules["h5py"].File = _File
    sys.modules["faiss"].IndexBinaryFlat = _ShimIndexBinaryFlat
    tmp = tempfile.mkdtemp(prefix="g8_")
    try:
        store[os.path.join(tmp, "ref.h5")] = {"calldata/GT": ref_gt, "variants/POS": ref_pos}
        store[os.path.join(tmp, "tgt.h5")] = {"calldata/GT": tgt_gt, "variants/POS": tgt_pos}
        with open(os.path.join(tmp, "ref.panel"), "w") as f:
            f.write("sample\tpop\tsuper_pop\tgender\n")
            for i in range(s_ref):
                f.write(f"S{i}\tP{i % 5}\t{['AFR', 'EUR', 'EAS'][i % 3]}\tmale\n")
        with open(os.path.join(tmp, "win.csv"), "w") as f:
            f.write("start,end\n" + "".join(f"{a},{b}\n" for a, b in windows))

        def load(name):
            spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
            m = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(m)
            return m

        db = os.path.join(tmp, "db")
        load("build_ref_db_intersect").build_ref_db_intersect(argparse.Namespace(
            ref_vcf=os.path.join(tmp, "ref.h5"), ref_panel=os.path.join(tmp, "ref.panel"),
            window_csv=os.path.join(tmp, "win.csv"), output_dir=db))
        tfi = load("test_faiss_intersect")
        out = {"ref_gt": ref_gt, "tgt_gt": tgt_gt, "ref_pos": ref_pos, "tgt_pos": tgt_pos, "windows": windows, "k": np.array(k)}
        for mode, shim in (("l2", _ShimIndexFlatL2), ("binary", _ShimIndexBinaryFlat)):
            rec = []
            orig_add, orig_search = shim.add, shim.search

            def rec_add(self, x, _o=orig_add):
                rec.append(["add", np.array(x)])
                return _o(self, x)

            def rec_search(self, x, kk, _o=orig_search):
                D, I = _o(self, x, kk)
                rec.append(["search", np.array(x), D, I])
                return D, I

            shim.add, shim.search = rec_add, rec_search
            tfi.test_faiss_intersect(argparse.Namespace(
                target_vcf=os.path.join(tmp, "tgt.h5"), ref_db=db, window_csv=os.path.join(tmp, "win.csv"),
                distance_mode=mode, top_k=k, sample_idx=-1, show_snps=False, show_snp_len=10))
            shim.add, shim.search = orig_add, orig_search
            assert len(rec) == 2 * len(windows)
            for w in range(len(windows)):
                out[f"{mode}_added_{w}"] = rec[2 * w][1]
                out[f"{mode}_query_{w}"], out[f"{mode}_D_{w}"], out[f"{mode}_I_{w}"] = rec[2 * w + 1][1:]
        for w in range(len(windows)):
            out[f"window_{w}"] = np.load(os.path.join(db, f"window_{w}.npy"))
            out[f"window_pos_{w}"] = np.load(os.path.join(db, f"window_{w}_pos.npy"))
        np.savez_compressed(os.path.join(OUT, "g8_intersect_workflow.npz"), **out)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def make_g9():
    """g9  V18 training retrieval   the real EmbeddingRAGDataset.process_batch_retrieval (embedding_rag_dataset.py:285-444)
                                    + the real BERTEmbedding (src/model/embedding/bert.py) in TRAIN mode (dropout 0) on a
                                    batch whose samples alternate between two windows (so the JIT cache is rebuilt
                                    mid-batch, :334-377), k = 2.  Pins the outputs rag_emb_h1 / rag_emb_h2 [B, k, L, D]
                                    AND the gradients of  sum(rag_emb_h1 * W1) + sum(rag_emb_h2 * W2)  w.r.t. every
                                    parameter of the embedding layer - the re-encode-with-grad path (:406-438)."""
    import torch

    from src.dataset.vocab import WordVocab
    from src.dataset.embedding_rag_dataset import EmbeddingRAGDataset
    from src.model.embedding.bert import BERTEmbedding

    rng = np.random.default_rng(20261019)
    torch.manual_seed(11)
    vocab = WordVocab(["AFR", "EUR", "EAS"])
    L, D, N, B, k = 64, 16, 40, 8, 2
    emb = BERTEmbedding(vocab_size=len(vocab), embed_size=D, dropout=0.0, use_af=True)
    emb.train()
    W = 2
    ref_tokens, ref_af, masks = [], [], []
    # Tie-free by construction (float32 cdist and any other exact engine must then agree on the ids): reference rows come
    # in pairs, row 2i+1 = row 2i with 5 observed sites flipped; a query is a row 2i with 2 OTHER observed sites flipped, so
    # its nearest rows are 2i (2 mismatches) and 2i+1 (7), every other row being ~20 away
    for w in range(W):
        m = np.zeros(L, np.int64)
        m[1:-1] = rng.random(L - 2) < 0.3
        masks.append(m)
        obs = np.flatnonzero(m[1:-1] == 0)
        hap = (rng.random((N, L - 2)) < 0.35).astype(np.int64)
        for i in range(0, N, 2):
            hap[i + 1] = hap[i]
            hap[i + 1, rng.choice(obs, 5, replace=False)] ^= 1
        ref_tokens.append(np.concatenate([np.full((N, 1), 2), np.where(hap == 0, 5, 6), np.full((N, 1), 3)], axis=1).astype(np.int64))
        ref_af.append(rng.random(L).astype(np.float32))
    window_idx = [0, 1, 0, 1, 1, 0, 0, 1]

    def queries():
        rows = []
        for w in window_idx:
            src = ref_tokens[w][2 * rng.integers(0, N // 2)].copy()
            obs = np.flatnonzero(masks[w][1:-1] == 0) + 1
            pos = rng.choice(obs, 2, replace=False)
            src[pos] = 11 - src[pos]  # 5 <-> 6
            src[masks[w] == 1] = vocab.mask_index
            rows.append(src)
        return np.stack(rows)

    h1, h2 = queries(), queries()
    af = np.stack([ref_af[w] for w in window_idx]).astype(np.float32)
    fake = types.SimpleNamespace(
        embed_dim=D, jit_cache_win_idx=-1, jit_ref_emb_search=None, jit_ref_tokens_raw=None, jit_ref_af_raw=None,
        ref_tokens_complete=ref_tokens, ref_af_windows=ref_af, window_masks=masks, vocab=vocab)
    fake._apply_mask_to_tokens_gpu = lambda t, m: EmbeddingRAGDataset._apply_mask_to_tokens_gpu(fake, t, m)
    batch = {"hap_1": torch.from_numpy(h1), "hap_2": torch.from_numpy(h2), "af": torch.from_numpy(af), "window_idx": list(window_idx)}
    out = EmbeddingRAGDataset.process_batch_retrieval(fake, batch, emb, "cpu", k)
    W1 = torch.from_numpy(rng.standard_normal((B, k, L, D)).astype(np.float32))
    W2 = torch.from_numpy(rng.standard_normal((B, k, L, D)).astype(np.float32))
    loss = (out["rag_emb_h1"] * W1).sum() + (out["rag_emb_h2"] * W2).sum()
    emb.zero_grad()
    loss.backward()
    rec = {"L": np.array(L), "D": np.array(D), "N": np.array(N), "k": np.array(k), "mask_index": np.array(vocab.mask_index),
           "vocab_size": np.array(len(vocab)), "window_idx": np.array(window_idx), "hap_1": h1, "hap_2": h2, "af": af,
           "rag_emb_h1": out["rag_emb_h1"].detach().numpy(), "rag_emb_h2": out["rag_emb_h2"].detach().numpy(),
           "W1": W1.numpy(), "W2": W2.numpy(), "loss": np.array(float(loss))}
    for w in range(W):
        rec[f"ref_tokens_{w}"] = ref_tokens[w]
        rec[f"ref_af_{w}"] = ref_af[w]
        rec[f"mask_{w}"] = masks[w]
    for name, t in emb.state_dict().items():
        rec["state/" + name] = t.detach().numpy()
    for name, prm in emb.named_parameters():
        rec["grad/" + name] = (prm.grad if prm.grad is not None else torch.zeros_like(prm)).numpy()
    np.savez_compressed(os.path.join(OUT, "g9_v18_train_grad.npz"), **rec)
    print("g9: loss", float(loss), "params", [n for n, _ in emb.named_parameters()])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] in ("g5", "g6", "g7", "g8", "g9"):  # only the newer fixtures (leaves g1-g4 untouched)
        install_stubs()
        sys.path.insert(0, REF)
        os.chdir("/tmp")
        {"g5": make_g5, "g6": make_g6, "g7": make_g7, "g8": make_g8, "g9": make_g9}[sys.argv[1]]()
    else:
        main()
