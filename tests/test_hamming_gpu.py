"""GPU parity: CUDA Hamming / masked-Hamming search (through the C ABI) vs the CPU oracle.
Bit-exact on D and I, ties included."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import cbind

pytestmark = pytest.mark.gpu


def _idx(d, W=1):
    from rag_snvbert_b200 import WindowedHammingIndex

    return WindowedHammingIndex(d, W)


def _check(panel, queries, k, observed=None, dtype="u8"):
    """panel [W,N,d], queries [W,Q,d] uint8 0/1 -> compare with oracle per window."""
    from rag_snvbert_b200 import _lib

    W, N, d = panel.shape
    idx = _idx(d, W)
    stride = _lib.packed_stride(d)

    def conv(x):
        if dtype == "u8":
            return x
        if dtype == "f32":
            return x.astype(np.float32)
        if dtype == "bool":
            return x.astype(bool)
        if dtype == "packed":
            return O.pack_bits_u32(x.reshape(-1, d), stride).reshape(x.shape[:-1] + (stride,))
        raise AssertionError(dtype)

    idx.add(conv(panel))
    assert idx.ntotal == N
    D, I = idx.search(conv(queries), k, observed=None if observed is None else conv(observed))
    assert D.dtype == np.int32 and I.dtype == np.int64
    for w in range(W):
        De, Ie = O.hamming_topk(panel[w], queries[w], k, None if observed is None else observed[w])
        np.testing.assert_array_equal(I[w], Ie, err_msg=f"window {w} ids")
        np.testing.assert_array_equal(D[w], De, err_msg=f"window {w} distances")
    return D, I


@pytest.mark.parametrize("dtype", ["u8", "f32", "bool", "packed"])
def test_small_all_dtypes(dtype):
    rng = np.random.default_rng(0)
    panel = (rng.random((1, 100, 50)) < 0.3).astype(np.uint8)
    q = (rng.random((1, 7, 50)) < 0.3).astype(np.uint8)
    _check(panel, q, 3, dtype=dtype)


@pytest.mark.parametrize("d", [1, 31, 32, 33, 100, 128, 257, 500, 777, 1024, 1030, 1056, 1100, 1500, 2060, 2176])
def test_site_counts(d):
    """every register-resident word bucket plus the generic fallback (d > 68 words)."""
    rng = np.random.default_rng(d)
    panel = (rng.random((2, 300, d)) < 0.4).astype(np.uint8)
    q = (rng.random((2, 37, d)) < 0.4).astype(np.uint8)
    _check(panel, q, 8)


def test_generic_wide_rows():
    rng = np.random.default_rng(5)
    d = 5000
    panel = (rng.random((1, 200, d)) < 0.5).astype(np.uint8)
    q = (rng.random((1, 40, d)) < 0.5).astype(np.uint8)
    obs = (rng.random((1, 40, d)) < 0.5).astype(np.uint8)
    _check(panel, q, 8)
    _check(panel, q, 16, observed=obs)


def test_chr21_window_shape_k8():
    """one BASELINE cfg-2 window: 5008 x 1030, mosaic haplotypes (many exact ties)."""
    panel = O.hapgen(2000, 5008, 1030)[None]
    q = O.hapgen(5000, 300, 1030, founder_seed=2000)[None]
    D, I = _check(panel, q, 8, dtype="packed")
    # ties must actually occur for this to be a tie test
    assert (D[0][:, 1:] == D[0][:, :-1]).any()


@pytest.mark.parametrize("k", [1, 3, 5, 8, 9, 16, 32])
def test_k_values(k):
    panel = O.hapgen(11, 1000, 1030)[None]
    q = O.hapgen(12, 130, 1030, founder_seed=11)[None]
    _check(panel, q, k)


def test_duplicates_and_ties_keep_lowest_ids():
    rng = np.random.default_rng(3)
    base = (rng.random((10, 200)) < 0.5).astype(np.uint8)
    panel = np.tile(base, (40, 1))[None]  # every row appears 40 times
    q = base[:5][None]
    D, I = _check(panel, q, 32)
    assert (D[0] == 0).all()
    np.testing.assert_array_equal(I[0][0], np.arange(0, 320, 10))


def test_planted_exact_match_returns_row_at_distance_zero():
    panel = O.hapgen(21, 3000, 1030)[None]
    q = panel[:, [5, 77, 2999]].copy()
    D, I = _check(panel, q, 1)
    np.testing.assert_array_equal(I[0][:, 0], [5, 77, 2999])
    assert (D[0] == 0).all()


def test_fewer_rows_than_k_pads_with_minus_one():
    rng = np.random.default_rng(4)
    panel = (rng.random((1, 5, 64)) < 0.5).astype(np.uint8)
    q = (rng.random((1, 9, 64)) < 0.5).astype(np.uint8)
    D, I = _check(panel, q, 8)
    assert (I[0][:, 5:] == -1).all() and (D[0][:, 5:] == np.iinfo(np.int32).max).all()


def test_per_query_observed_mask_cfg3():
    """BASELINE cfg 3: per-query missing rate U[0.1, 0.9], masked Hamming, k=8."""
    rng = np.random.default_rng(8000)
    panel = O.hapgen(2001, 5008, 1030)[None]
    q = O.hapgen(5001, 200, 1030, founder_seed=2001)[None]
    rate = rng.uniform(0.1, 0.9, size=(1, 200, 1))
    missing = (rng.random((1, 200, 1030)) < rate).astype(np.uint8)
    _check(panel, q, 8, observed=1 - missing)
    _check(panel, q, 8, observed=1 - missing, dtype="packed")
    # reference convention: mask marks MISSING sites (partial_faiss_intersect.py:91)
    idx = _idx(1030)
    idx.add(panel[0])
    D, I = idx.search(q[0], 8, missing=missing[0])
    De, Ie = O.hamming_topk(panel[0], q[0], 8, 1 - missing[0])
    np.testing.assert_array_equal(I, Ie)
    np.testing.assert_array_equal(D, De)
    # packed + missing flag exercises the device-side inversion with zeroed pad bits
    from rag_snvbert_b200 import _lib

    s = _lib.packed_stride(1030)
    D, I = idx.search(O.pack_bits_u32(q[0], s), 8, missing=O.pack_bits_u32(missing[0], s))
    np.testing.assert_array_equal(I, Ie)
    np.testing.assert_array_equal(D, De)


def test_per_window_shared_mask():
    rng = np.random.default_rng(9)
    panel = (rng.random((3, 400, 300)) < 0.5).astype(np.uint8)
    q = (rng.random((3, 50, 300)) < 0.5).astype(np.uint8)
    obs = (rng.random((3, 300)) < 0.7).astype(np.uint8)
    idx = _idx(300, 3)
    idx.add(panel)
    D, I = idx.search(q, 5, observed=obs)
    for w in range(3):
        De, Ie = O.hamming_topk(panel[w], q[w], 5, obs[w])
        np.testing.assert_array_equal(I[w], Ie)
        np.testing.assert_array_equal(D[w], De)


def test_row_split_path_small_query_batches():
    """nq = 2 (the reference's training-time call, rag_train_dataset.py:281) and a large panel:
    the planner splits rows across CTAs and merges partial top-k."""
    panel = O.hapgen(31, 20000, 1030)[None]
    q = O.hapgen(32, 2, 1030, founder_seed=31)[None]
    _check(panel, q, 5)
    q = O.hapgen(33, 48, 1030, founder_seed=31)[None]
    _check(panel, q, 32)


def test_many_windows_and_window_offset():
    W = 12
    panel = np.stack([O.hapgen(100 + w, 600, 1030) for w in range(W)])
    q = np.stack([O.hapgen(200 + w, 150, 1030, founder_seed=100 + w) for w in range(W)])
    _check(panel, q, 8)
    idx = _idx(1030, W)
    idx.add(panel)
    D, I = idx.search(q[4:9], 8, w0=4, id_offset=1000)
    for j, w in enumerate(range(4, 9)):
        De, Ie = O.hamming_topk(panel[w], q[w], 8)
        np.testing.assert_array_equal(I[j], Ie + 1000)
        np.testing.assert_array_equal(D[j], De)


def test_incremental_add_grows_panel():
    rng = np.random.default_rng(6)
    panel = (rng.random((2, 900, 130)) < 0.5).astype(np.uint8)
    q = (rng.random((2, 33, 130)) < 0.5).astype(np.uint8)
    idx = _idx(130, 2)
    for s in (0, 100, 350):
        e = {0: 100, 100: 350, 350: 900}[s]
        idx.add(panel[:, s:e])
    assert idx.ntotal == 900
    D, I = idx.search(q, 8)
    for w in range(2):
        De, Ie = O.hamming_topk(panel[w], q[w], 8)
        np.testing.assert_array_equal(I[w], Ie)
        np.testing.assert_array_equal(D[w], De)
    np.testing.assert_array_equal(idx.export_packed(1)[:, : idx.words], O.pack_bits_u32(panel[1]))


def test_float_distance_output_equals_faiss_l2_on_binary_rows():
    panel = O.hapgen(41, 500, 1030)[None]
    q = O.hapgen(42, 20, 1030, founder_seed=41)[None]
    idx = _idx(1030)
    idx.add(panel[0].astype(np.float32))
    D, I = idx.search(q[0].astype(np.float32), 4, dist_dtype=np.float32)
    De, Ie = O.l2_topk_f32_blas(panel[0].astype(np.float32), q[0].astype(np.float32), 4)
    np.testing.assert_array_equal(I, Ie)
    np.testing.assert_array_equal(D, De)


def test_torch_cuda_tensors_zero_copy():
    import torch

    panel = O.hapgen(51, 2000, 1030)
    q = O.hapgen(52, 257, 1030, founder_seed=51)
    idx = _idx(1030)
    idx.add(torch.from_numpy(panel).cuda())
    D, I = idx.search(torch.from_numpy(q).cuda(), 8)
    assert D.is_cuda and I.is_cuda and D.dtype == torch.int32 and I.dtype == torch.int64
    De, Ie = O.hamming_topk(panel, q, 8)
    np.testing.assert_array_equal(I.cpu().numpy(), Ie)
    np.testing.assert_array_equal(D.cpu().numpy(), De)


def test_token_rows_shared_mask_equals_token_space_l2():
    """V17 layout (rag_train_dataset.py:111-134, 262-281): tokenised panel and queries with the
    window's mask; squared L2 over tokens == masked Hamming."""
    rng = np.random.default_rng(7)
    lw = 1000
    raw_mask = (rng.random(lw) < 0.3).astype(np.int64)
    pm = O.sequence_padding(raw_mask)
    panel01 = O.hapgen(61, 400, lw)
    q01 = O.hapgen(62, 30, lw, founder_seed=61)
    ptok = O.tokenize(panel01, pm)
    qtok = O.tokenize(q01, pm)
    idx = _idx(O.MAX_SEQ_LEN)
    idx.add(ptok)
    D, I = idx.search(qtok, 3, dist_dtype=np.float32)
    De, Ie = O.token_l2_topk(ptok, qtok, 3)
    np.testing.assert_array_equal(I, Ie)
    np.testing.assert_array_equal(D, De)


def test_binary_codes_indexbinaryflat():
    import rag_snvbert_b200.faiss_compat as faiss

    rng = np.random.default_rng(10)
    d_bits = 2 * 516
    ref = (rng.random((700, d_bits)) < 0.5).astype(np.uint8)
    t = (rng.random((60, d_bits)) < 0.5).astype(np.uint8)
    index = faiss.IndexBinaryFlat(d_bits)
    index.add(O.packbits_msb(ref))
    D, I = index.search(O.packbits_msb(t), 5)
    De, Ie = O.hamming_topk(ref, t, 5)
    np.testing.assert_array_equal(I, Ie)
    np.testing.assert_array_equal(D, De)
    assert index.ntotal == 700 and index.d == d_bits


def test_gather_tokens_matches_reference_layout():
    rng = np.random.default_rng(12)
    lw, S = 1000, 60
    gt = (rng.random((lw, S, 2)) < 0.3).astype(np.int8)  # raw_ref_window [L_w, S, 2]
    rows = O.panel_rows_from_gt(gt).astype(np.uint8)
    d = 1028
    padded = np.zeros((rows.shape[0], d), np.uint8)
    padded[:, :lw] = rows
    idx = _idx(d)
    idx.add(padded)
    I = rng.integers(0, 2 * S, size=(9, 3)).astype(np.int64)
    I[0, 1] = -1
    out = idx.gather_tokens(I, n_sites=lw, seq_len=O.MAX_SEQ_LEN)
    exp = O.gather_tokens(gt, I)
    np.testing.assert_array_equal(out, exp)


def test_row_sharded_merge_equals_unsharded():
    import torch
    from rag_snvbert_b200 import topk_merge

    panel = O.hapgen(71, 4000, 1030)
    q = O.hapgen(72, 100, 1030, founder_seed=71)
    De, Ie = O.hamming_topk(panel, q, 32)
    parts_D, parts_I = [], []
    for g in range(4):
        idx = _idx(1030)
        idx.add(torch.from_numpy(panel[g * 1000:(g + 1) * 1000]).cuda())
        D, I = idx.search(torch.from_numpy(q).cuda(), 32, id_offset=g * 1000)
        parts_D.append(D)
        parts_I.append(I)
    D, I = topk_merge(torch.stack(parts_D), torch.stack(parts_I), 32)
    np.testing.assert_array_equal(I.cpu().numpy(), Ie)
    np.testing.assert_array_equal(D.cpu().numpy(), De)


def test_c_oracle_agrees_on_gpu_sized_case():
    """the C restatement (bench's CPU baseline) against the CUDA path at 2 full cfg-2 windows."""
    from rag_snvbert_b200 import _lib

    s = _lib.packed_stride(1030)
    P = np.stack([O.pack_bits_u32(O.hapgen(2000 + w, 5008, 1030), s) for w in range(2)])
    Q = np.stack([O.pack_bits_u32(O.hapgen(5000 + w, 2000, 1030, founder_seed=2000 + w), s) for w in range(2)])
    idx = _idx(1030, 2)
    idx.add(P)
    D, I = idx.search(Q, 8)
    Dc, Ic = cbind.hamming_topk_packed(P, Q, 8, words=33)
    np.testing.assert_array_equal(I, Ic)
    np.testing.assert_array_equal(D, Dc)


def test_error_behaviour():
    idx = _idx(64)
    with pytest.raises(ValueError):
        idx.add(np.zeros((3, 65), np.uint8))
    with pytest.raises(ValueError):
        idx.search(np.zeros((3, 64), np.uint8), 0)
    idx.add(np.zeros((3, 64), np.uint8))
    # k > 32: the Python classes take the block path (faiss pads with -1 when k > ntotal); caller-owned results are a
    # stated limit there, and the C-ABI itself refuses k > 32
    D, I = idx.search(np.zeros((3, 64), np.uint8), 100)
    assert D.shape == (3, 100) and (np.sort(I[:, :3], axis=1) == np.arange(3)).all() and (I[:, 3:] == -1).all()
    with pytest.raises(ValueError):
        idx.search(np.zeros((3, 64), np.uint8), 100, out=(np.empty((3, 100), np.int32), np.empty((3, 100), np.int64)))
    from rag_snvbert_b200 import _lib as L

    q = np.zeros((3, 64), np.uint8)
    Dc, Ic = np.empty((3, 40), np.int32), np.empty((3, 40), np.int64)
    rc = L.lib().snv_index_search(idx._h, 0, 1, q.ctypes.data, 3, L.DT_U8, None, L.MASK_NONE, 40, 0, Dc.ctypes.data, None, Ic.ctypes.data, 0, None)
    assert rc != 0 and "32" in L.lib().snv_last_error().decode()


def test_host_pipeline_many_windows_all_mask_modes():
    """>= 8 windows with host buffers: the library pipelines window chunks over internal streams."""
    rng = np.random.default_rng(77)
    W, N, Q, d = 21, 700, 90, 1030
    panel = (rng.random((W, N, d)) < 0.4).astype(np.uint8)
    q = (rng.random((W, Q, d)) < 0.4).astype(np.uint8)
    obs_q = (rng.random((W, Q, d)) < 0.6).astype(np.uint8)
    obs_w = (rng.random((W, d)) < 0.6).astype(np.uint8)
    idx = _idx(d, W)
    idx.add(panel)
    outD = np.empty((W, Q, 8), np.int32)
    outI = np.empty((W, Q, 8), np.int64)
    for obs, kw in ((None, {}), (obs_q, {"observed": obs_q}), (obs_w, {"observed": obs_w}),
                    (obs_q, {"missing": 1 - obs_q})):
        D, I = idx.search(q, 8, out=(outD, outI), **kw)
        assert D.base is outD or D is outD
        for w in range(W):
            m = None if obs is None else (obs[w] if obs.ndim == 3 else obs[w])
            De, Ie = O.hamming_topk(panel[w], q[w], 8, m)
            np.testing.assert_array_equal(I[w], Ie)
            np.testing.assert_array_equal(D[w], De)
    # token queries (own observed plane) through the pipelined host path
    pm = O.sequence_padding((rng.random(1000) < 0.3).astype(np.int64))
    ptok = np.stack([O.tokenize(O.hapgen(900 + w, 300, 1000), pm) for w in range(9)])
    qtok = np.stack([O.tokenize(O.hapgen(950 + w, 17, 1000, founder_seed=900 + w), pm) for w in range(9)])
    tidx = _idx(O.MAX_SEQ_LEN, 9)
    tidx.add(ptok)
    D, I = tidx.search(qtok, 3, dist_dtype=np.float32)
    for w in range(9):
        De, Ie = O.token_l2_topk(ptok[w], qtok[w], 3)
        np.testing.assert_array_equal(I[w], Ie)
        np.testing.assert_array_equal(D[w], De)


def test_grouped_ragged_batches_one_launch():
    """a training batch regrouped by window (rag_train_dataset.py:239-281): ragged group sizes."""
    rng = np.random.default_rng(91)
    W, N, d = 7, 2008, 1030
    panel = np.stack([O.hapgen(300 + w, N, d) for w in range(W)])
    wid = rng.integers(0, W, size=137).astype(np.int32)
    wid[wid == 3] = 2  # leave one window without queries
    q = np.stack([O.hapgen(400 + i, 1, d, founder_seed=300 + int(w))[0] for i, w in enumerate(wid)])
    obs = (rng.random(q.shape) < 0.7).astype(np.uint8)
    idx = _idx(d, W)
    idx.add(panel)
    from rag_snvbert_b200 import _lib

    n0 = _lib.launch_count()
    D, I = idx.search_grouped(q, wid, 5)
    assert _lib.launch_count() - n0 <= 3  # pack + scan (+ merge): not one launch per window
    Dm, Im = idx.search_grouped(q, wid, 5, observed=obs)
    for i, w in enumerate(wid):
        De, Ie = O.hamming_topk(panel[w], q[i:i + 1], 5)
        np.testing.assert_array_equal(I[i], Ie[0])
        np.testing.assert_array_equal(D[i], De[0])
        De, Ie = O.hamming_topk(panel[w], q[i:i + 1], 5, obs[i:i + 1])
        np.testing.assert_array_equal(Im[i], Ie[0])
        np.testing.assert_array_equal(Dm[i], De[0])
    # big groups take 128-query blocks
    wid2 = np.repeat(np.arange(W, dtype=np.int32), 150)
    q2 = np.concatenate([O.hapgen(500 + w, 150, d, founder_seed=300 + w) for w in range(W)])
    perm = rng.permutation(len(wid2))
    D2, I2 = idx.search_grouped(q2[perm], wid2[perm], 8)
    for w in range(W):
        sel = np.where(wid2[perm] == w)[0]
        De, Ie = O.hamming_topk(panel[w], q2[perm][sel], 8)
        np.testing.assert_array_equal(I2[sel], Ie)
        np.testing.assert_array_equal(D2[sel], De)


def test_rag_retriever_reproduces_reference_collate_golden():
    """tests/golden/g2: the reference's own rag_collate_fn_with_dataset outputs."""
    import os

    from rag_snvbert_b200.collate import RagRetriever

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "g2_v17_collate.npz"))
    k = int(g["k"])
    r = RagRetriever([g["ref_gt_0"], g["ref_gt_1"]])
    # golden rows are in the collate's regrouped order (window 0 samples, then window 1)
    wid = g["window_idx"]
    h1, h2 = r.retrieve(wid, g["hap_1"], g["hap_2"], k)
    np.testing.assert_array_equal(h1, g["rag_seg_h1"])
    np.testing.assert_array_equal(h2, g["rag_seg_h2"])
    D, I = r.search(wid, g["hap_1"], g["hap_2"], k)
    n0 = int((wid == 0).sum())
    np.testing.assert_array_equal(I[:n0].reshape(-1, k), g["search_I_0"])
    np.testing.assert_array_equal(D[:n0].reshape(-1, k), g["search_D_0"])
    np.testing.assert_array_equal(I[n0:].reshape(-1, k), g["search_I_1"])
    np.testing.assert_array_equal(D[n0:].reshape(-1, k), g["search_D_1"])
    # torch CUDA inputs (the trainer-process pattern): same result, device tensors out
    import torch

    t1, t2 = r.retrieve(torch.from_numpy(wid), torch.from_numpy(g["hap_1"]).cuda(), torch.from_numpy(g["hap_2"]).cuda(), k)
    assert t1.is_cuda
    np.testing.assert_array_equal(t1.cpu().numpy(), g["rag_seg_h1"])
    np.testing.assert_array_equal(t2.cpu().numpy(), g["rag_seg_h2"])


def test_empty_index_and_empty_queries():
    idx = _idx(1030)
    q = np.zeros((5, 1030), np.uint8)
    D, I = idx.search(q, 8)
    assert (I == -1).all() and (D == np.iinfo(np.int32).max).all()
    idx.add(np.zeros((3, 1030), np.uint8))
    D, I = idx.search(np.zeros((0, 1030), np.uint8), 8)
    assert D.shape == (0, 8) and I.shape == (0, 8)
    idx.reset()
    assert idx.ntotal == 0
    from rag_snvbert_b200 import WindowedL2Index

    l2 = WindowedL2Index(64)
    D, I = l2.search(np.zeros((4, 64), np.float32), 3)
    assert (I == -1).all() and (D == O.F32_MAX).all()


def test_k32_large_panel_list_selection():
    """k = 32 selects through shared-memory candidate lists (BASELINE cfg 5 shard shape, scaled)."""
    panel = O.hapgen(81, 25000, 1030)[None]
    q = O.hapgen(82, 130, 1030, founder_seed=81)[None]
    _check(panel, q, 32, dtype="packed")
    rng = np.random.default_rng(83)
    obs = (rng.random((1, 130, 1030)) < 0.5).astype(np.uint8)
    _check(panel[:, :6000], q, 17, observed=obs)
    # adversarial order for the lagging threshold: distances strictly decreasing along the scan
    d = 1030
    rows = np.zeros((1, 900, d), np.uint8)
    for i in range(900):
        rows[0, i, : 900 - i] = 1
    _check(rows, np.zeros((1, 40, d), np.uint8), 32)


@pytest.mark.parametrize("parts,kin,kout,dt", [(2, 8, 8, "i32"), (8, 32, 32, "i32"), (8, 32, 32, "f32"), (3, 5, 4, "f32"),
                                             (20, 32, 32, "i32"), (40, 32, 16, "i32")])
def test_topk_merge_arbitrary_order_and_padding(parts, kin, kout, dt):
    """snv_topk_merge against a numpy lexsort: unsorted per-part lists, -1 padding, ties on D broken by id;
    covers the warp-cooperative kernels (<= 256 and <= 1024 candidates per query) and the generic one."""
    import torch

    from rag_snvbert_b200 import topk_merge

    rng = np.random.default_rng(parts * 1000 + kin)
    nq = 777
    D = rng.integers(0, 50, size=(parts, nq, kin)).astype(np.int32)
    I = rng.permuted(np.tile(np.arange(parts * kin, dtype=np.int64) * 7 + (1 << 33), (nq, 1)), axis=1).reshape(nq, parts, kin).transpose(1, 0, 2).copy()
    pad = rng.random((parts, nq, kin)) < 0.15
    I[pad] = -1
    Dx = D.astype(np.float32) * 0.5 - 3.0 if dt == "f32" else D
    padv = np.float32(3.4028234663852886e38) if dt == "f32" else np.int32(2**31 - 1)
    Dx = np.where(pad, padv, Dx).astype(Dx.dtype)
    Dm, Im = topk_merge(torch.from_numpy(Dx).cuda(), torch.from_numpy(I).cuda(), kout)
    Dm, Im = Dm.cpu().numpy(), Im.cpu().numpy()
    for q in range(0, nq, 37):
        d = Dx[:, q, :].reshape(-1)
        i = I[:, q, :].reshape(-1)
        keep = i >= 0
        d, i = d[keep], i[keep]
        order = np.lexsort((i, d))[:kout]
        exp_d = np.full(kout, padv, dtype=Dx.dtype)
        exp_i = np.full(kout, -1, dtype=np.int64)
        exp_d[: len(order)] = d[order]
        exp_i[: len(order)] = i[order]
        np.testing.assert_array_equal(Im[q], exp_i)
        np.testing.assert_array_equal(Dm[q], exp_d)


@pytest.mark.parametrize("shape", [(3, 700, 300, 1030, 8), (10, 5008, 260, 1030, 8), (1, 40, 9, 77, 5), (2, 3, 4, 33, 8)])
@pytest.mark.parametrize("where", ["host", "device"])
def test_compact_results_equal_the_wide_ones(shape, where):
    """search_compact: (uint16 D, int32 I) must carry exactly the values of (int32 D, int64 I), padding included
    (k > rows), from host buffers (pipelined chunks) and from device tensors, with and without masks."""
    import torch

    W, N, Q, d, k = shape
    rng = np.random.default_rng(sum(shape))
    panel = (rng.random((W, N, d)) < 0.4).astype(np.uint8)
    q = (rng.random((W, Q, d)) < 0.4).astype(np.uint8)
    obs = (rng.random((W, Q, d)) < 0.7).astype(np.uint8)
    idx = _idx(d, W)
    idx.add(panel)
    for m in (None, obs):
        if where == "host":
            D, I = idx.search(q, k, observed=m)
            Dc, Ic = idx.search_compact(q, k, observed=m)
            assert Dc.dtype == np.uint16 and Ic.dtype == np.int32
        else:
            qd = torch.from_numpy(q).cuda()
            md = None if m is None else torch.from_numpy(m).cuda()
            D, I = idx.search(qd, k, observed=md)
            Dc, Ic = idx.search_compact(qd, k, observed=md)
            D, I = D.cpu().numpy(), I.cpu().numpy()
            Dc, Ic = Dc.cpu().numpy().view(np.uint16), Ic.cpu().numpy()
        np.testing.assert_array_equal(Ic.astype(np.int64), I)
        np.testing.assert_array_equal(np.where(I < 0, 0x7FFFFFFF, Dc.astype(np.int32)), D)
        if k > N:
            assert (Ic[..., N:] == -1).all() and (Dc[..., N:] == 0xFFFF).all()


def test_row_sharded_search_object_single_rank_and_key_packing():
    """RowShardedSearch on one rank is the plain search; pack/unpack of the exchange keys on the device"""
    import torch

    from rag_snvbert_b200.sharding import RowShardedSearch

    rng = np.random.default_rng(5)
    panel = (rng.random((2, 600, 200)) < 0.4).astype(np.uint8)
    q = torch.from_numpy((rng.random((2, 50, 200)) < 0.4).astype(np.uint8)).cuda()
    idx = _idx(200, 2)
    idx.add(panel)
    s = RowShardedSearch(idx, 1000, world=1)
    lo, hi, D, I = s.search(q, 8)
    D0, I0 = idx.search(q, 8, id_offset=1000)
    assert (lo, hi) == (0, 50) and torch.equal(D, D0) and torch.equal(I, I0)
    key = RowShardedSearch.pack_keys(D, I)
    D2, I2 = RowShardedSearch.unpack_keys(key)
    assert torch.equal(D2, D) and torch.equal(I2, I)


def test_exchange_pack_and_merge_kernels_match_torch_reference():
    """snv_exchange_pack / snv_exchange_merge (the row-sharded exchange on the device) against the torch statement of the
    same steps: destination-major int64 keys, then the k best of [parts] sorted lists per row, padding included."""
    import torch

    from rag_snvbert_b200 import _lib as L
    from rag_snvbert_b200.sharding import RowShardedSearch

    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    for (nw, nq, k, G) in [(3, 40, 8, 4), (2, 64, 32, 8), (1, 6, 5, 2)]:
        D = torch.randint(0, 1031, (nw, nq, k), device="cuda", generator=g, dtype=torch.int32).sort(dim=2).values
        I = torch.randint(0, 200000, (nw, nq, k), device="cuda", generator=g, dtype=torch.int64)
        I[:, :, -1] = torch.where(torch.rand((nw, nq), device="cuda", generator=g) < 0.2, -1, I[:, :, -1])
        D[:, :, -1] = torch.where(I[:, :, -1] < 0, 0x7FFFFFFF, D[:, :, -1])
        keys = torch.empty((G, nw, nq // G, k), dtype=torch.int64, device="cuda")
        L.check(L.lib().snv_exchange_pack(0, D.data_ptr(), I.data_ptr(), nw, nq, k, G, keys.data_ptr(), 0), "pack")
        torch.cuda.synchronize()
        ref = RowShardedSearch.pack_keys(D, I).reshape(nw, G, nq // G, k).permute(1, 0, 2, 3).contiguous()
        assert torch.equal(keys, ref)
        # merge: treat the G destination blocks as G source ranks' lists for n rows
        n = nw * (nq // G)
        lists = keys.reshape(G, n, k).sort(dim=2).values.contiguous()
        Do = torch.empty((n, k), dtype=torch.int32, device="cuda")
        Io = torch.empty((n, k), dtype=torch.int64, device="cuda")
        L.check(L.lib().snv_exchange_merge(0, lists.data_ptr(), G, n, k, k, Do.data_ptr(), Io.data_ptr(), 0), "merge")
        torch.cuda.synchronize()
        allk = lists.permute(1, 0, 2).reshape(n, G * k).sort(dim=1).values[:, :k]
        De, Ie = RowShardedSearch.unpack_keys(allk)
        assert torch.equal(Do, De) and torch.equal(Io, Ie)


def test_merge_kernels_take_unsorted_lists_and_short_k():
    """the warp sorting-network merges do not rely on sorted inputs: random key order, k_in 8 / 20 / 32, 1-9 parts"""
    import torch

    from rag_snvbert_b200 import _lib as L
    from rag_snvbert_b200.sharding import RowShardedSearch

    g = torch.Generator(device="cuda")
    g.manual_seed(11)
    for (G, n, kin, kout) in [(1, 50, 32, 32), (9, 333, 32, 32), (5, 1000, 8, 8), (3, 77, 20, 7), (8, 64, 1, 1), (2, 10, 16, 16)]:
        D = torch.randint(0, 1031, (G, n, kin), device="cuda", generator=g, dtype=torch.int32)
        I = torch.randperm(G * n * kin, device="cuda", generator=g).reshape(G, n, kin).to(torch.int64)  # unique ids
        I = torch.where(torch.rand((G, n, kin), device="cuda", generator=g) < 0.1, -1, I)
        keys = RowShardedSearch.pack_keys(D, I).contiguous()
        Do = torch.empty((n, kout), dtype=torch.int32, device="cuda")
        Io = torch.empty((n, kout), dtype=torch.int64, device="cuda")
        L.check(L.lib().snv_exchange_merge(0, keys.data_ptr(), G, n, kin, kout, Do.data_ptr(), Io.data_ptr(), 0), "merge")
        torch.cuda.synchronize()
        allk = keys.permute(1, 0, 2).reshape(n, G * kin).sort(dim=1).values[:, :kout]
        De, Ie = RowShardedSearch.unpack_keys(allk)
        assert torch.equal(Do, De) and torch.equal(Io, Ie), (G, n, kin, kout)


@pytest.mark.parametrize("world,nw,nq,k", [(2, 3, 40, 8), (4, 2, 64, 32), (8, 1, 80, 32), (3, 2, 30, 5)])
def test_peer_exchange_fused_kernel_single_process(world, nw, nq, k):
    """snv_peer_exchange (pack + NVLink push + flags + merge in one launch) with `world` exchange objects wired by pointer
    inside this process, one stream each: every rank's output equals the k best of all ranks' candidates for its query
    slice; three batches in a row exercise both receive slots and the epoch flags."""
    import torch

    from rag_snvbert_b200.sharding import PeerExchange, RowShardedSearch

    g = torch.Generator(device="cuda")
    g.manual_seed(world * 100 + k)
    peers = PeerExchange.connect_local(0, world, nw * nq * k * 8)
    streams = [torch.cuda.Stream() for _ in range(world)]
    qg = nq // world
    try:
        for batch in range(3):
            Ds, Is = [], []
            for r in range(world):
                D = torch.randint(0, 1031, (nw, nq, k), device="cuda", generator=g, dtype=torch.int32).sort(dim=2).values
                I = torch.randint(0, 1 << 30, (nw, nq, k), device="cuda", generator=g, dtype=torch.int64) * world + r  # unique over ranks
                I[:, :, -1] = torch.where(torch.rand((nw, nq), device="cuda", generator=g) < 0.2, -1, I[:, :, -1])
                D[:, :, -1] = torch.where(I[:, :, -1] < 0, 0x7FFFFFFF, D[:, :, -1])
                Ds.append(D.contiguous())
                Is.append(I.contiguous())
            torch.cuda.synchronize()
            outs = []
            for r in range(world):
                with torch.cuda.stream(streams[r]):
                    outs.append(peers[r].exchange(Ds[r], Is[r], k))
            torch.cuda.synchronize()
            keys = torch.stack([RowShardedSearch.pack_keys(Ds[r], Is[r]) for r in range(world)])  # [src, nw, nq, k]
            for r in range(world):
                mine = keys[:, :, r * qg:(r + 1) * qg]                                              # [src, nw, qg, k]
                best = mine.permute(1, 2, 0, 3).reshape(nw, qg, world * k).sort(dim=2).values[:, :, :k]
                De, Ie = RowShardedSearch.unpack_keys(best)
                assert torch.equal(outs[r][0], De) and torch.equal(outs[r][1], Ie), (batch, r)
    finally:
        for p in peers:
            p.close()


@pytest.mark.parametrize("d", [1030, 100, 1024, 33])
def test_dense_packed_rows_equal_strided_rows(d):
    """SNV_DT_PACKED_U32_DENSE: packed rows without the stride padding (the 132-byte wire format) for add, queries and
    masks (observed and missing) give exactly the results of the strided rows, from host and from device buffers"""
    import torch

    from rag_snvbert_b200 import _lib

    rng = np.random.default_rng(d)
    W, N, Q, k = 2, 600, 70, 8
    stride, words = _lib.packed_stride(d), _lib.packed_words(d)
    panel = (rng.random((W, N, d)) < 0.4).astype(np.uint8)
    q = (rng.random((W, Q, d)) < 0.4).astype(np.uint8)
    miss = (rng.random((W, Q, d)) < 0.3).astype(np.uint8)
    pk = lambda x: O.pack_bits_u32(x.reshape(-1, d), stride).reshape(x.shape[:-1] + (stride,))  # noqa: E731
    idx = _idx(d, W)
    idx.add(panel)
    D0, I0 = idx.search(q, k, missing=miss)
    D1, I1 = idx.search(q, k)
    qd, md = np.ascontiguousarray(pk(q)[..., :words]), np.ascontiguousarray(pk(miss)[..., :words])
    if words == stride:
        return  # no padding for this d: the dense and the strided layouts coincide
    for conv in (lambda x: x, lambda x: torch.from_numpy(x.view(np.int32)).cuda()):
        D, I = idx.search(conv(qd), k, missing=conv(md))
        D, I = (D.cpu().numpy(), I.cpu().numpy()) if hasattr(D, "cpu") else (D, I)
        np.testing.assert_array_equal(I, I0)
        np.testing.assert_array_equal(D, D0)
        D, I = idx.search(conv(qd), k)
        D, I = (D.cpu().numpy(), I.cpu().numpy()) if hasattr(D, "cpu") else (D, I)
        np.testing.assert_array_equal(I, I1)
        np.testing.assert_array_equal(D, D1)
    idx2 = _idx(d, W)
    idx2.add(np.ascontiguousarray(pk(panel)[..., :words]))   # dense rows into add()
    D, I = idx2.search(q, k)
    np.testing.assert_array_equal(I, I1)
    np.testing.assert_array_equal(D, D1)


@pytest.mark.parametrize("k", [33, 100])
@pytest.mark.parametrize("where", ["host", "device"])
def test_k_above_32_block_path(k, where):
    """faiss accepts any k: above the in-register limit the Python classes take the block path (panel as windows of 32 rows,
    one k = 32 launch, top-k over the keys) - still bit-exact vs the oracle, ties, masks, padding (k > rows) included"""
    import torch

    rng = np.random.default_rng(k)
    W, N, Q, d = 2, 70, 9, 200
    panel = (rng.random((W, N, d)) < 0.4).astype(np.uint8)
    panel[:, N // 2:] = panel[:, : N - N // 2]           # duplicates -> ties
    q = (rng.random((W, Q, d)) < 0.4).astype(np.uint8)
    obs = (rng.random((W, Q, d)) < 0.7).astype(np.uint8)
    idx = _idx(d, W)
    idx.add(panel)
    conv = (lambda x: x) if where == "host" else (lambda x: torch.from_numpy(x).cuda())
    for m in (None, obs):
        D, I = idx.search(conv(q), k, observed=None if m is None else conv(m))
        D, I = (D.cpu().numpy(), I.cpu().numpy()) if hasattr(D, "cpu") else (D, I)
        assert D.shape == (W, Q, k) and D.dtype == np.int32 and I.dtype == np.int64
        for w in range(W):
            De, Ie = O.hamming_topk(panel[w], q[w], k, None if m is None else m[w])
            np.testing.assert_array_equal(I[w], Ie)
            np.testing.assert_array_equal(D[w], De)
    # through the faiss surface (float 0/1 rows -> the L2 index's block path), single window
    import rag_snvbert_b200.faiss_compat as faiss

    f = faiss.IndexFlatL2(d)
    f.add(panel[0].astype(np.float32))
    Df, If = f.search(q[0].astype(np.float32), k)
    De, Ie = O.hamming_topk(panel[0], q[0], k)
    np.testing.assert_array_equal(If, Ie)
    np.testing.assert_array_equal(Df[Ie >= 0], De[Ie >= 0].astype(np.float32))
    assert (Df[Ie < 0] > 3e38).all()
