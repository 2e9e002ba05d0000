"""GPU parity: tcgen05 squared-L2 search (through the C ABI) vs the CPU oracle.

Tolerances (stated, DESIGN.md "Parity"):
  * integer-valued inputs (0/1 genotypes, V17 token vectors): D and I bit-exact in both modes.
  * general fp32 inputs, mode tf32x3: |D - D_f64| <= 1e-5 * (|q|^2 + |r|^2) (the tensor core's fp32
    accumulation truncates: ~1 ulp of the running dot product per MMA step, measured);
    mode tf32: |D - D_f64| <= 2e-3 * (|q|^2 + |r|^2).
  * ids equal the float64 ranking except where the float64 distances of the two candidates
    differ by less than twice that tolerance (ties inside the tolerance).
"""
import os

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = {"tf32x3": 1e-5, "tf32": 2e-3}


def _l2(d, precision="tf32x3", W=1):
    from rag_snvbert_b200 import WindowedL2Index

    return WindowedL2Index(d, W, None, precision)


def _check_float(refs, q, k, precision):
    idx = _l2(refs.shape[1], precision)
    idx.add(refs)
    assert idx.ntotal == refs.shape[0]
    D, I = idx.search(q, k)
    assert D.dtype == np.float32 and I.dtype == np.int64 and D.shape == (q.shape[0], k)
    d64 = O.l2_matrix_f64(refs, q)
    D64, I64 = O._topk_lex(d64, k, np.float64(O.F32_MAX))
    qn = (q.astype(np.float64) ** 2).sum(1)
    rn_max = (refs.astype(np.float64) ** 2).sum(1).max()
    tol = TOL[precision] * (qn + rn_max)
    valid = I >= 0
    got = np.take_along_axis(d64, np.where(valid, I, 0), axis=1)
    # distances returned vs float64 distances OF THE RETURNED ROWS
    assert (np.abs(D.astype(np.float64) - got)[valid] <= np.broadcast_to(tol[:, None], D.shape)[valid]).all(), \
        f"max |dD| = {np.abs(D - got)[valid].max()} tol {tol.max()}"
    assert (np.diff(D, axis=1) >= 0).all(), "D not ascending"
    n_diff = O.assert_ids_match_within_tolerance(d64, I, I64, 2 * tol)
    return D, I, n_diff


@pytest.mark.parametrize("precision", ["tf32x3", "tf32"])
def test_cfg4_shape_gaussian(precision):
    rng = np.random.default_rng(4001)
    refs = rng.standard_normal((5008, 256)).astype(np.float32)
    q = np.random.default_rng(4002).standard_normal((4096, 256)).astype(np.float32)
    D, I, n_diff = _check_float(refs, q, 8, precision)
    if precision == "tf32x3":
        assert n_diff <= 40  # essentially the float64 ranking (4096 x 8 positions)


def test_cfg4_clustered_near_ties():
    rng = np.random.default_rng(4003)
    cent = rng.standard_normal((64, 256)).astype(np.float32)
    refs = (cent[rng.integers(0, 64, 5008)] + 0.05 * rng.standard_normal((5008, 256))).astype(np.float32)
    q = (cent[rng.integers(0, 64, 1000)] + 0.05 * rng.standard_normal((1000, 256))).astype(np.float32)
    _check_float(refs, q, 8, "tf32x3")


@pytest.mark.parametrize("n,nq,d,k", [(1000, 200, 256, 8), (300, 5, 100, 3), (257, 129, 33, 8), (256, 128, 32, 1),
                                       (77, 300, 64, 32), (3, 10, 16, 8), (2500, 64, 1000, 16)])
def test_ragged_shapes(n, nq, d, k):
    rng = np.random.default_rng(n + nq + d)
    refs = rng.standard_normal((n, d)).astype(np.float32)
    q = rng.standard_normal((nq, d)).astype(np.float32)
    D, I, _ = _check_float(refs, q, k, "tf32x3")
    if n < k:
        assert (I[:, n:] == -1).all() and (D[:, n:] == O.F32_MAX).all()


@pytest.mark.parametrize("precision", ["tf32x3", "tf32"])
def test_cfg1_binary_float_rows_exact(precision):
    """BASELINE cfg 1 (batch_test_faiss_l2.py path): float 0/1 panel 5008 x 1030, 1000 queries, k=1."""
    panel = O.hapgen(1001, 5008, 1030).astype(np.float32)
    q = O.hapgen(1002, 1000, 1030, founder_seed=1001).astype(np.float32)
    idx = _l2(1030, precision)
    idx.add(panel)
    D, I = idx.search(q, 1)
    De, Ie = O.hamming_topk(panel, q, 1)
    np.testing.assert_array_equal(I, Ie)
    np.testing.assert_array_equal(D, De.astype(np.float32))
    D8, I8 = idx.search(q[:300], 8)
    De, Ie = O.hamming_topk(panel, q[:300], 8)
    np.testing.assert_array_equal(I8, Ie)
    np.testing.assert_array_equal(D8, De.astype(np.float32))


@pytest.mark.parametrize("precision", ["tf32x3", "tf32"])
def test_v17_token_vectors_exact_with_different_masks(precision):
    """IndexFlatL2(1030) on tokenised rows (rag_train_dataset.py:132-134,281) where query and panel
    masks DIFFER (dynamic mask / epoch 0): pair costs (4,5)->1, (4,6)->4, (5,6)->1 — still exact."""
    rng = np.random.default_rng(17)
    lw = 1000
    pm = O.sequence_padding((rng.random(lw) < 0.3).astype(np.int64))
    qm = O.sequence_padding((rng.random(lw) < 0.3).astype(np.int64))
    ptok = O.tokenize(O.hapgen(171, 2008, lw), pm)
    qtok = O.tokenize(O.hapgen(172, 64, lw, founder_seed=171), qm)
    idx = _l2(O.MAX_SEQ_LEN, precision)
    idx.add(ptok.astype(np.float32))
    D, I = idx.search(qtok.astype(np.float32), 5)
    De, Ie = O.token_l2_topk(ptok, qtok, 5)
    np.testing.assert_array_equal(I, Ie)
    np.testing.assert_array_equal(D, De)


def test_golden_g2_v17_collate_through_faiss_compat():
    """The reference's own collate outputs (tests/golden/g2): index side via faiss_compat.IndexFlatL2,
    gather via the Hamming index's token gather."""
    import rag_snvbert_b200.faiss_compat as faiss
    from rag_snvbert_b200 import IndexHamming

    g = np.load(os.path.join(G, "g2_v17_collate.npz"))
    k, lw = int(g["k"]), int(g["lw"])
    h1, h2 = [], []
    for w in range(2):
        index = faiss.IndexFlatL2(O.MAX_SEQ_LEN)
        index.add(g[f"ref_tokenized_{w}"].astype(np.float32))
        D, I = index.search(g[f"search_q_{w}"], k)
        np.testing.assert_array_equal(I, g[f"search_I_{w}"])
        np.testing.assert_array_equal(D, g[f"search_D_{w}"])
        rows = O.panel_rows_from_gt(g[f"ref_gt_{w}"]).astype(np.uint8)
        hidx = IndexHamming(lw)
        hidx.add(rows)
        # token rows straight into the bit-packed index give the same neighbours
        Dh, Ih = hidx.search((g[f"search_q_{w}"][:, 1:1 + lw] == 6).astype(np.uint8), k,
                             missing=g[f"raw_mask_{w}"].astype(np.uint8), dist_dtype=np.float32)
        np.testing.assert_array_equal(Ih, I)
        np.testing.assert_array_equal(Dh, D)
        seg = hidx.gather_tokens(I, n_sites=lw)
        h1.append(seg[0::2])
        h2.append(seg[1::2])
    np.testing.assert_array_equal(np.concatenate(h1), g["rag_seg_h1"])
    np.testing.assert_array_equal(np.concatenate(h2), g["rag_seg_h2"])


def test_golden_g4_v18_embedding_search_and_gather():
    g = np.load(os.path.join(G, "g4_v18_embedding.npz"))
    k = int(g["k"])
    ref = np.ascontiguousarray(g["ref_flat"])
    idx = _l2(ref.shape[1])
    idx.add(ref)
    comp = np.ascontiguousarray(g["ref_complete"].reshape(ref.shape[0], -1))
    cidx = _l2(comp.shape[1])
    cidx.add(comp)
    for qn, In, out in (("q1_flat", "I1", "rag_emb_h1"), ("q2_flat", "I2", "rag_emb_h2")):
        q = np.ascontiguousarray(g[qn])
        D, I = idx.search(q, k)
        d64 = O.l2_matrix_f64(ref, q)
        tol = 1e-5 * ((q.astype(np.float64) ** 2).sum(1) + (ref.astype(np.float64) ** 2).sum(1).max())
        O.assert_ids_match_within_tolerance(d64, I, g[In], 2 * tol)
        rows = cidx.gather_rows(np.ascontiguousarray(g[In]))
        np.testing.assert_array_equal(rows.reshape(g[out].shape), g[out])


def test_faiss_compat_write_read_roundtrip(tmp_path):
    import rag_snvbert_b200.faiss_compat as faiss

    rng = np.random.default_rng(5)
    x = rng.standard_normal((500, 96)).astype(np.float32)
    q = rng.standard_normal((20, 96)).astype(np.float32)
    index = faiss.IndexFlatL2(96)
    index.add(x)
    D, I = index.search(q, 4)
    faiss.write_index(index, str(tmp_path / "w.faiss"))
    again = faiss.read_index(str(tmp_path / "w.faiss"))
    assert again.ntotal == 500 and again.d == 96
    D2, I2 = again.search(q, 4)
    np.testing.assert_array_equal(I, I2)
    np.testing.assert_array_equal(D, D2)
    res = faiss.StandardGpuResources()
    assert faiss.index_cpu_to_gpu(res, 0, again) is again
    with pytest.raises(AssertionError):
        index.add(np.zeros((3, 95), np.float32))
    b = faiss.IndexBinaryFlat(64)
    b.add(rng.integers(0, 256, (50, 8), dtype=np.uint8))
    faiss.write_index(b, str(tmp_path / "b.faiss"))
    b2 = faiss.read_index(str(tmp_path / "b.faiss"))
    qq = rng.integers(0, 256, (5, 8), dtype=np.uint8)
    np.testing.assert_array_equal(b.search(qq, 3)[1], b2.search(qq, 3)[1])


def test_torch_cuda_tensors_and_cdist_agreement():
    """zero-copy device path, and agreement with the reference's live GPU path (torch.cdist+topk)."""
    import torch

    torch.manual_seed(0)
    refs = torch.randn(2008, 512, device="cuda")
    q = torch.randn(48, 512, device="cuda")
    idx = _l2(512)
    idx.add(refs)
    D, I = idx.search(q, 4)
    assert D.is_cuda and I.is_cuda
    d64 = O.l2_matrix_f64(refs.cpu().numpy(), q.cpu().numpy())
    _, Iref = torch.cdist(q, refs, p=2).topk(4, largest=False, dim=1)
    tol = 1e-5 * (float((q ** 2).sum(1).max()) + float((refs ** 2).sum(1).max()))
    O.assert_ids_match_within_tolerance(d64, I.cpu().numpy(), Iref.cpu().numpy(), 4 * tol)


def test_multi_window_l2():
    rng = np.random.default_rng(8)
    refs = rng.standard_normal((3, 400, 64)).astype(np.float32)
    q = rng.standard_normal((3, 50, 64)).astype(np.float32)
    idx = _l2(64, W=3)
    idx.add(refs)
    D, I = idx.search(q, 5)
    for w in range(3):
        D64, I64 = O.l2_topk_f64(refs[w], q[w], 5)
        np.testing.assert_array_equal(I[w], I64)


def _embedding_like(rng, n, nq, L, D):
    """rows = per-position token embedding + a large position/AF term shared by every row
    (src/model/embedding/bert.py:53-75): the V18 search vectors."""
    T = rng.standard_normal((7, D)).astype(np.float32)
    P = (3.0 * rng.standard_normal((L, D))).astype(np.float32)
    founders = rng.random((12, L)) < 0.25
    mask = rng.random(L) < 0.3

    def emb(m):
        h = founders[rng.integers(0, 12, m)] ^ (rng.random((m, L)) < 0.03)
        tok = np.where(h, 6, 5)
        tok[:, mask] = 4
        return (T[tok] + P[None]).reshape(m, L * D).astype(np.float32)

    return emb(n), emb(nq)


@pytest.mark.parametrize("center", [False, True])
def test_skinny_deep_problem_split_k(center):
    """nq << 128, deep vectors (the reference's real V18 shape, scaled): split-K path."""
    from rag_snvbert_b200 import WindowedL2Index

    rng = np.random.default_rng(18)
    refs, q = _embedding_like(rng, 2008, 24, 160, 48)  # d = 7680
    idx = WindowedL2Index(refs.shape[1], 1, None, "tf32x3", center=center)
    idx.add(refs)
    D, I = idx.search(q, 4)
    d64 = O.l2_matrix_f64(refs, q)
    qn = (q.astype(np.float64) ** 2).sum(1)
    rn = (refs.astype(np.float64) ** 2).sum(1).max()
    tol = 1e-5 * (qn + rn)
    got = np.take_along_axis(d64, I, axis=1)
    err = np.abs(D.astype(np.float64) - got)
    assert (err <= tol[:, None]).all(), f"max err {err.max()} tol {tol.min()}"
    O.assert_ids_match_within_tolerance(d64, I, O._topk_lex(d64, 4, np.float64(O.F32_MAX))[1], 2 * tol)
    if center:
        # centring removes the shared component: errors drop by orders of magnitude
        assert err.max() <= 1e-6 * float(qn.max() + rn) + 1e-3


def test_l2_cta_pair_kernel_matches_default():
    """the optional CTA-pair (cta_group::2) L2 kernel returns what the default kernel returns"""
    import os

    import torch

    from rag_snvbert_b200 import WindowedL2Index

    torch.manual_seed(3)
    refs = torch.randn(3000, 200, device="cuda")
    q = torch.randn(700, 200, device="cuda")  # 6 query tiles: 3 pairs
    idx = WindowedL2Index(200, 1, 0)
    idx.add(refs)
    D0, I0 = idx.search(q, 8)
    old = os.environ.get("SNV_L2_PAIR")
    os.environ["SNV_L2_PAIR"] = "1"
    try:
        D1, I1 = idx.search(q, 8)
    finally:
        if old is None:
            del os.environ["SNV_L2_PAIR"]
        else:
            os.environ["SNV_L2_PAIR"] = old
    assert torch.equal(I0, I1) and torch.equal(D0, D1)


def test_indexflatl2_binary_fast_path_equals_float_kernel():
    """BASELINE cfg 1 through the faiss-shaped surface: 0/1 float rows take the bit-packed Hamming engine and
    return exactly what the float L2 kernel returns (same D, same I, ties included); anything non-binary falls
    back to the float kernel."""
    import rag_snvbert_b200.faiss_compat as faiss

    panel = O.hapgen(1001, 5008, 1030).astype(np.float32)
    q = O.hapgen(1002, 1000, 1030, founder_seed=1001).astype(np.float32)
    fast = faiss.IndexFlatL2(1030)
    fast.binary_min_work = 0  # force the Hamming engine at this (small) size
    slow = faiss.IndexFlatL2(1030, binary_fast_path=False)
    fast.add(panel[:3000])
    fast.add(panel[3000:])
    slow.add(panel)
    for k in (1, 8):
        Df, If = fast.search(q, k)
        Ds, Is = slow.search(q, k)
        assert fast.last_search_path == "hamming" and slow.last_search_path == "l2"
        assert Df.dtype == np.float32 and If.dtype == np.int64
        np.testing.assert_array_equal(If, Is)
        np.testing.assert_array_equal(Df, Ds)
    De, Ie = O.hamming_topk(panel.astype(np.uint8), q.astype(np.uint8), 8)
    np.testing.assert_array_equal(If, Ie)
    np.testing.assert_array_equal(Df, De.astype(np.float32))
    # non-binary queries: float kernel
    q2 = q.copy()
    q2[0, 0] = 0.5
    D2, I2 = fast.search(q2, 8)
    assert fast.last_search_path == "l2"
    np.testing.assert_array_equal(I2[1:], Is[1:])
    # a non-binary row switches the index to the float kernel for good
    fast.add(np.full((1, 1030), 0.25, dtype=np.float32))
    fast.search(q, 8)
    assert fast.last_search_path == "l2"


def test_reference_real_depth_v18_shape_default_mode():
    """The reference's REAL inference / training retrieval shape (embedding_rag_infer_dataset.py:173-181,281-285;
    embedding_rag_dataset.py:392-402): N = 2008 reference haplotypes, 48 queries, vectors of L*D = 1030*192 = 197,760
    floats, built like BERTEmbedding output (token + position/AF components shared by all rows).  DEFAULT index
    settings (tf32x3, centring decided automatically) must meet the stated 1e-5 (|q|^2 + |r|^2) against float64,
    return the float64 nearest neighbour except inside that tolerance, be bit-reproducible run to run (split-K
    reduces in a fixed order) and agree with the reference's own GPU path, a warmed torch.cdist + topk."""
    import torch

    from rag_snvbert_b200 import WindowedL2Index

    L, Dm, N, nq, k = 1030, 192, 2008, 48, 1
    dev = "cuda"
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    T = torch.randn(7, Dm, device=dev, generator=g)
    P = torch.randn(L, Dm, device=dev, generator=g)
    founders = torch.rand(16, L, device=dev, generator=g) < 0.25
    mask = torch.rand(L, device=dev, generator=g) < 0.3

    def embed(n):
        h = founders[torch.randint(0, 16, (n,), device=dev, generator=g)].clone()
        h ^= torch.rand(n, L, device=dev, generator=g) < 0.02
        tok = torch.where(h, 6, 5)
        tok[:, mask] = 4
        return (T[tok] + P[None]).reshape(n, L * Dm).contiguous()

    refs, q = embed(N), embed(nq)
    # float64 reference in column blocks (the full double copies would be 3.2 GB + ...: fine, but blocks keep it light)
    d64 = torch.zeros((nq, N), dtype=torch.float64, device=dev)
    for c0 in range(0, L * Dm, 16384):
        a, b = q[:, c0:c0 + 16384].double(), refs[:, c0:c0 + 16384].double()
        d64 += (a * a).sum(1)[:, None] + (b * b).sum(1)[None, :] - 2.0 * (a @ b.T)
    idx = WindowedL2Index(L * Dm, 1, 0)          # defaults: tf32x3, centring decided on the first add
    idx.add(refs)
    D1, I1 = idx.search(q, k)
    D2, I2 = idx.search(q, k)
    assert torch.equal(D1, D2) and torch.equal(I1, I2), "split-K search is not reproducible"
    scale = ((q.double() ** 2).sum(1) + (refs.double() ** 2).sum(1).max())
    tol = 1e-5 * scale
    got = torch.gather(d64, 1, I1)
    err = (D1.double() - got).abs()
    assert bool((err <= tol[:, None]).all()), f"max |dD| {float(err.max())} vs tol {float(tol.min())}"
    best = d64.min(1).values
    assert bool(((got[:, 0] - best) <= 2 * tol).all()), "a returned neighbour is outside the tolerance of the float64 nearest"
    # the reference's GPU path on the same tensors (warmed): same neighbours except inside the tolerance
    for _ in range(2):
        It = torch.cdist(q, refs, p=2).topk(k, dim=1, largest=False).indices
    agree = (It == I1)
    gap = (torch.gather(d64, 1, It) - got).abs()
    assert bool((agree | (gap <= 2 * tol[:, None])).all())
