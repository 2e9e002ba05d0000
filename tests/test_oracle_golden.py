"""CPU: pins the oracle against golden vectors produced by the reference's own Python
(tests/golden/make_golden.py, run in the build container against /root/reference)."""
import os

import numpy as np

from oracle import oracle as O
from oracle import cbind

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(G, name))


def test_g1_tokenize_matches_reference():
    g = load("g1_tokenize.npz")
    assert list(g["specials"]) == [O.PAD, O.UNK, O.SOS, O.EOS, O.MASK, O.ALLELE0, O.ALLELE1]
    for name in "abc":
        seq, raw = g[f"seq_{name}"], g[f"rawmask_{name}"]
        padded = O.sequence_padding(raw)
        np.testing.assert_array_equal(padded, g[f"padmask_{name}"])
        np.testing.assert_array_equal(O.tokenize(seq, padded), g[f"tok_masked_{name}"])
        np.testing.assert_array_equal(O.tokenize(seq, None), g[f"tok_plain_{name}"])


def test_g2_v17_panel_layout_search_and_gather():
    g = load("g2_v17_collate.npz")
    k = int(g["k"])
    h1_rows, h2_rows = [], []
    # the collate groups samples by window (window 0 first), two queries per sample
    for w in range(2):
        ref_gt = g[f"ref_gt_{w}"]
        rows = O.panel_rows_from_gt(ref_gt)
        ref_tok = O.tokenize(rows, g[f"padded_mask_{w}"])
        np.testing.assert_array_equal(ref_tok, g[f"ref_tokenized_{w}"])  # index-side layout
        q = g[f"search_q_{w}"]
        D, I = O.token_l2_topk(ref_tok, q.astype(np.int64), k)
        np.testing.assert_array_equal(I, g[f"search_I_{w}"])
        np.testing.assert_array_equal(D, g[f"search_D_{w}"])
        # token-space squared L2 == masked Hamming when panel and queries share the mask
        obs = 1 - g[f"raw_mask_{w}"]
        lw = int(g["lw"])
        q01 = (q[:, 1:1 + lw] == 6).astype(np.uint8)
        Dh, Ih = O.hamming_topk(rows.astype(np.uint8), q01, k, obs)
        np.testing.assert_array_equal(Ih, I)
        np.testing.assert_array_equal(Dh.astype(np.float32), D)
        seg = O.gather_tokens(ref_gt, I)
        h1_rows.append(seg[0::2])
        h2_rows.append(seg[1::2])
    np.testing.assert_array_equal(np.concatenate(h1_rows), g["rag_seg_h1"])
    np.testing.assert_array_equal(np.concatenate(h2_rows), g["rag_seg_h2"])
    # planted exact match: sample 0 / hap 1 retrieves panel row 2*7+1 at distance 0
    assert g["search_I_0"][0, 0] == 15 and g["search_D_0"][0, 0] == 0


def test_g3_bitpack_matches_reference():
    g = load("g3_bitpack.npz")
    np.testing.assert_array_equal(O.packbits_msb(g["arr"]), g["packed"])


def test_g4_v18_cdist_topk_ids_and_gather():
    g = load("g4_v18_embedding.npz")
    k = int(g["k"])
    for qn, In, dn, out in (("q1_flat", "I1", "dists_h1", "rag_emb_h1"), ("q2_flat", "I2", "dists_h2", "rag_emb_h2")):
        # torch.cdist + topk ran in fp32: ids may differ from the float64 ranking only at
        # near-ties below fp32 resolution (tolerance 1e-6 * (|q|^2 + |r|^2), stated in DESIGN.md)
        d2 = O.l2_matrix_f64(g["ref_flat"], g[qn])
        tol = 1e-6 * ((g[qn].astype(np.float64) ** 2).sum(1) + (g["ref_flat"].astype(np.float64) ** 2).sum(1).max())
        I = O.cdist_topk_indices(g["ref_flat"], g[qn], k)
        n_diff = O.assert_ids_match_within_tolerance(d2, I, g[In], tol)
        assert n_diff <= 1
        # torch.cdist returns the NON-squared distance from the fp32 expansion; compare squares
        np.testing.assert_allclose(g[dn].astype(np.float64) ** 2, d2, rtol=0, atol=float(tol.max()) * 4)
        D32, I32 = O.l2_topk_f32_blas(g["ref_flat"], g[qn], k)
        O.assert_ids_match_within_tolerance(d2, I32, g[In], tol)
        I = g[In]
        rows = O.gather_rows(g["ref_complete"].reshape(g["ref_complete"].shape[0], -1), I)
        np.testing.assert_allclose(rows.reshape(g[out].shape), g[out], rtol=1e-6, atol=1e-6)


def test_c_restatement_equals_numpy_oracle():
    rng = np.random.default_rng(1)
    for (n, nq, d, k) in ((300, 50, 1030, 8), (64, 9, 100, 32), (5, 4, 33, 8), (1000, 20, 2060, 1)):
        p = (rng.random((n, d)) < 0.4).astype(np.uint8)
        q = (rng.random((nq, d)) < 0.4).astype(np.uint8)
        m = (rng.random((nq, d)) < 0.6).astype(np.uint8)
        words = (d + 31) // 32
        stride = -(-words // 4) * 4
        P, Q, M = (O.pack_bits_u32(x, stride) for x in (p, q, m))
        for mask01, Mp in ((None, None), (m, M)):
            De, Ie = O.hamming_topk(p, q, k, mask01)
            Dc, Ic = cbind.hamming_topk_packed(P, Q, k, Mp, words=words)
            np.testing.assert_array_equal(Ic, Ie)
            np.testing.assert_array_equal(Dc, De)


def test_oracle_invariants():
    p = O.hapgen(3, 500, 200)
    q = p[[4, 99]].copy()
    D, I = O.hamming_topk(p, q, 1)
    assert list(I[:, 0]) == [4, 99] and (D == 0).all()
    # row id <-> (sample, hap) convention (rag_train_dataset.py:298-299)
    gt = np.arange(3 * 4 * 2).reshape(3, 4, 2)
    rows = O.panel_rows_from_gt(gt)
    for s in range(4):
        for h in range(2):
            np.testing.assert_array_equal(rows[2 * s + h], gt[:, s, h])
    # sharded == unsharded
    De, Ie = O.hamming_topk(p, O.hapgen(4, 20, 200, founder_seed=3), 8)
    parts = [O.hamming_topk(p[s:s + 125], O.hapgen(4, 20, 200, founder_seed=3), 8) for s in range(0, 500, 125)]
    Dm, Im = O.merge_topk([d for d, _ in parts], [i + 125 * g for g, (_, i) in enumerate(parts)], 8, O.I32_MAX)
    np.testing.assert_array_equal(Im, Ie)
    np.testing.assert_array_equal(Dm, De)


def test_g5_observed_site_search_matches_reference():
    """partial_faiss_intersect.py: expand_target_to_ref (:46-80) and build_partial_index_l2 (:82-111) run by
    tests/golden/make_golden.py; the product's host helper and the oracle's masked Hamming must reproduce them."""
    from rag_snvbert_b200 import refdb

    g = load("g5_partial_intersect.npz")
    k = int(g["k"])
    # product host code (numpy, no GPU involved) == the reference's own function
    expanded, missing = refdb.expand_target_to_ref(g["ref_pos"], g["tgt_data"], g["tgt_pos"])
    np.testing.assert_array_equal(expanded, g["expanded"])
    np.testing.assert_array_equal(missing, g["missing_ref"])
    rows_p = refdb.sample_rows(g["ref_data"], g["windows"])
    rows_q = refdb.sample_rows(g["expanded"], g["windows"])
    for w, (a, b) in enumerate(g["windows"]):
        wl = b - a
        panel = rows_p[w][:, : 2 * wl]
        np.testing.assert_array_equal(panel, np.transpose(g["ref_data"][a:b], (1, 0, 2)).reshape(panel.shape[0], -1))
        for s in range(rows_q.shape[1]):
            observed = np.repeat(1 - g["missing"][a:b, s], 2)  # both haplotypes of a site share the sample's site mask
            D, I = O.hamming_topk(panel, rows_q[w][s : s + 1, : 2 * wl], k, observed[None])
            np.testing.assert_array_equal(I[0], g["I_aligned"][w, s])
            np.testing.assert_array_equal(D[0].astype(np.float32), g["D_aligned"][w, s])
            # the script's literal call pairs misaligned columns ([h1.., h2..] against s0h0, s0h1, ..): restated here only
            # to show that the difference between `literal` and `aligned` is exactly that slip
            valid = np.where(g["missing"][a:b, s] == 0)[0]
            q_lit = np.concatenate([g["expanded"][a:b, s, 0][valid], g["expanded"][a:b, s, 1][valid]]).astype(np.int64)
            p_lit = np.transpose(g["ref_data"][a:b], (1, 0, 2))[:, valid, :].reshape(panel.shape[0], -1).astype(np.int64)
            d_lit = ((p_lit - q_lit[None]) ** 2).sum(1)
            order = np.argsort(d_lit, kind="stable")[:k]
            np.testing.assert_array_equal(order, g["I_literal"][w, s])
            np.testing.assert_array_equal(d_lit[order].astype(np.float32), g["D_literal"][w, s])


def test_g6_offline_db_workflow_matches_reference():
    """build_ref_db_l2.py + batch_test_faiss_l2.py run whole by tests/golden/make_golden.py: the sample-row vector layout
    (product host helper), the window_{i}.npy contents and the per-window batched search results."""
    from rag_snvbert_b200 import refdb

    g = load("g6_ref_db_workflow.npz")
    k = int(g["k"])
    rows_p = refdb.sample_rows(g["ref_raw"], g["windows"])  # codes > 0 count as alt (build_ref_db_l2.py:50)
    rows_q = refdb.sample_rows(g["tgt_raw"], g["windows"])
    for w, (a, b) in enumerate(g["windows"]):
        wl = b - a
        np.testing.assert_array_equal(rows_p[w][:, : 2 * wl], g[f"index_xb_{w}"].astype(np.uint8))
        assert (rows_p[w][:, 2 * wl:] == 0).all()
        np.testing.assert_array_equal(rows_q[w][:, : 2 * wl], g[f"query_{w}"].astype(np.uint8))
        np.testing.assert_array_equal(g[f"window_{w}"], np.transpose((g["ref_raw"][a:b] > 0), (1, 0, 2)).astype(np.int8))
        D, I = O.hamming_topk(rows_p[w], rows_q[w], k)  # zero padding adds nothing
        np.testing.assert_array_equal(I, g[f"I_{w}"])
        np.testing.assert_array_equal(D.astype(np.float32), g[f"D_{w}"])
        D2, I2 = O.l2_topk_f32_blas(g[f"index_xb_{w}"], g[f"query_{w}"], k)
        np.testing.assert_array_equal(I2, g[f"I_{w}"])
        np.testing.assert_array_equal(D2, g[f"D_{w}"])


def test_g7_v18_inference_search_and_gather():
    """EmbeddingRAGInferDataset.process_batch_retrieval (embedding_rag_infer_dataset.py:250-324): faiss flat L2 per
    window group + unique-id gather, run by tests/golden/make_golden.py with the real BERTEmbedding."""
    g = load("g7_v18_infer.npz")
    k = int(g["k"])
    for w in (0, 1):
        ref = g[f"ref_flat_{w}"]
        members = g[f"members_{w}"]
        np.testing.assert_array_equal(members, np.where(g["window_idx"] == w)[0])  # the window grouping
        for h in (1, 2):
            q, Ig, Dg = g[f"q{h}_flat_{w}"], g[f"I{h}_{w}"], g[f"D{h}_{w}"]
            d2 = O.l2_matrix_f64(ref, q)
            tol = 1e-6 * ((q.astype(np.float64) ** 2).sum(1) + (ref.astype(np.float64) ** 2).sum(1).max())
            D64, I64 = O.l2_topk_f64(ref, q, k)
            O.assert_ids_match_within_tolerance(d2, I64, Ig, tol)
            D32, I32 = O.l2_topk_f32_blas(ref, q, k)
            O.assert_ids_match_within_tolerance(d2, I32, Ig, tol)
            np.testing.assert_allclose(Dg, np.take_along_axis(d2, Ig, 1), rtol=0, atol=float(tol.max()) * 4)
            rows = O.gather_rows(g[f"ref_complete_{w}"].reshape(ref.shape[0], -1), Ig)
            np.testing.assert_array_equal(rows.reshape((len(members), k) + g["rag_emb_h1"].shape[2:]), g[f"rag_emb_h{h}"][members])


def test_g8_intersect_workflow_matches_reference():
    """build_ref_db_intersect.py + test_faiss_intersect.py (both distance modes) run whole by tests/golden/make_golden.py:
    the shared-site selection (product host helpers), the np.packbits codes and both modes' (D, I)."""
    from rag_snvbert_b200 import refdb

    g = load("g8_intersect_workflow.npz")
    k, win = int(g["k"]), g["windows"]
    obs = refdb.intersect_windows(g["ref_pos"], g["tgt_pos"], win)            # [W, Lmax] per site
    expanded, missing = refdb.expand_target_to_ref(g["ref_pos"], g["tgt_gt"], g["tgt_pos"])
    rows_p = refdb.sample_rows(g["ref_gt"], win)
    rows_q = refdb.sample_rows(expanded, win)
    for w, (a, b) in enumerate(win):
        wl = b - a
        np.testing.assert_array_equal(g[f"window_pos_{w}"], g["ref_pos"][a:b])
        np.testing.assert_array_equal(obs[w, :wl], 1 - missing[a:b, 0])       # the two host helpers agree on the shared sites
        cols = np.repeat(obs[w, :wl], 2).astype(bool)                          # both haplotypes of a shared site
        np.testing.assert_array_equal(rows_p[w][:, : 2 * wl][:, cols], g[f"l2_added_{w}"].astype(np.uint8))
        np.testing.assert_array_equal(rows_q[w][:, : 2 * wl][:, cols], g[f"l2_query_{w}"].astype(np.uint8))
        np.testing.assert_array_equal(O.packbits_msb(g[f"l2_added_{w}"].astype(np.uint8)), g[f"binary_added_{w}"])
        np.testing.assert_array_equal(O.packbits_msb(g[f"l2_query_{w}"].astype(np.uint8)), g[f"binary_query_{w}"])
        observed = np.zeros(rows_p.shape[2], np.uint8)
        observed[: 2 * wl] = cols
        D, I = O.hamming_topk(rows_p[w], rows_q[w], k, observed)
        for mode in ("l2", "binary"):
            np.testing.assert_array_equal(I, g[f"{mode}_I_{w}"])
            np.testing.assert_array_equal(D.astype(g[f"{mode}_D_{w}"].dtype), g[f"{mode}_D_{w}"])


from g9_helpers import _g9_check, _g9_setup  # noqa: E402


def test_g9_v18_training_retrieval_outputs_and_gradients_restated():
    """the oracle's restatement of BERTEmbedding + process_batch_retrieval reproduces what the reference's own code produced:
    rag_emb_h1 / rag_emb_h2 and the gradients w.r.t. every embedding parameter (fixture g9)"""
    from oracle.ref_embedding import v18_process_batch_retrieval

    g, emb, ref_tokens, ref_af, masks, batch, k = _g9_setup()
    out1, out2 = v18_process_batch_retrieval(ref_tokens, ref_af, masks, int(g["mask_index"]), batch, emb, k)
    _g9_check(g, emb, out1, out2)
