"""GPU: offline reference-DB workflow (build_ref_db_l2 / batch_test_faiss_l2 / intersect /
partial_faiss_intersect mirrors) against brute force in the scripts' own vector layout."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _data(seed=0, V=700, S_ref=120, S_t=9):
    rng = np.random.default_rng(seed)
    ref = O.hapgen(seed + 1, 2 * S_ref, V).reshape(S_ref, 2, V).transpose(2, 0, 1).copy()  # [V, S, 2]
    tgt = O.hapgen(seed + 2, 2 * S_t, V, founder_seed=seed + 1).reshape(S_t, 2, V).transpose(2, 0, 1).copy()
    win = np.array([[0, 250], [250, 480], [480, 700]])
    return rng, ref, tgt, win


def test_build_and_batch_search_match_faiss_layout(tmp_path):
    from rag_snvbert_b200 import refdb

    rng, ref, tgt, win = _data()
    index = refdb.build_ref_db(ref, win)
    D, I = refdb.batch_search(index, tgt, win, 4)
    for w, (a, b) in enumerate(win):
        # the literal vectors of build_ref_db_l2.py:86 / batch_test_faiss_l2.py:88
        xr = np.transpose(ref[a:b], (1, 0, 2)).reshape(ref.shape[1], -1).astype(np.float32)
        xq = np.transpose(tgt[a:b], (1, 0, 2)).reshape(tgt.shape[1], -1).astype(np.float32)
        De, Ie = O.l2_topk_f32_blas(xr, xq, 4)
        np.testing.assert_array_equal(I[w], Ie)
        np.testing.assert_array_equal(D[w], De)
    # window_{i}.npy round trip (build_ref_db_l2.py:77-78)
    for w, (a, b) in enumerate(win):
        np.save(tmp_path / f"window_{w}.npy", np.transpose(ref[a:b], (1, 0, 2)))
    again = refdb.load_ref_db(str(tmp_path), len(win))
    D2, I2 = refdb.batch_search(again, tgt, win, 4)
    np.testing.assert_array_equal(I2, I)
    np.testing.assert_array_equal(D2, D)


def test_expand_target_and_partial_search():
    from rag_snvbert_b200 import refdb

    rng, ref, tgt, win = _data(5)
    V = ref.shape[0]
    ref_pos = np.sort(rng.choice(10 * V, V, replace=False))
    keep = np.sort(rng.choice(V, int(0.6 * V), replace=False))
    tgt_pos = np.concatenate([ref_pos[keep], [10 * V + 5, 10 * V + 9]])  # plus sites absent from the ref
    tgt_data = np.concatenate([tgt[keep], np.ones((2,) + tgt.shape[1:], tgt.dtype)])
    expanded, missing = refdb.expand_target_to_ref(ref_pos, tgt_data, tgt_pos)
    # the reference's loop, literally (partial_faiss_intersect.py:63-78)
    tgt_dict = {p: i for i, p in enumerate(tgt_pos)}
    exp_e = np.zeros_like(expanded)
    mis_e = np.zeros_like(missing)
    for r_i, p in enumerate(ref_pos):
        if p in tgt_dict:
            exp_e[r_i] = tgt_data[tgt_dict[p]]
        else:
            mis_e[r_i] = 1
    np.testing.assert_array_equal(expanded, exp_e)
    np.testing.assert_array_equal(missing, mis_e)
    np.testing.assert_array_equal(refdb.intersect_windows(ref_pos, tgt_pos, win)[0, :250], 1 - mis_e[:250, 0])

    index = refdb.build_ref_db(ref, win)
    D, I = refdb.partial_search(index, expanded, missing, win, 5)
    for w, (a, b) in enumerate(win):
        for s in range(tgt.shape[1]):
            valid = np.where(missing[a:b, s] == 0)[0]
            sub_ref = np.transpose(ref[a:b], (1, 0, 2))[:, valid, :].reshape(ref.shape[1], -1).astype(np.float32)
            q = expanded[a:b, s][valid].reshape(1, -1).astype(np.float32)  # aligned (s0h0, s0h1, ...) columns
            De, Ie = O.l2_topk_f32_blas(sub_ref, q, 5)
            np.testing.assert_array_equal(I[w, s], Ie[0])
            np.testing.assert_array_equal(D[w, s], De[0])


def test_intersect_masks_on_device_match_host_and_drive_the_masked_search():
    """the ref/target position intersection as a device kernel -> per-window observed masks that feed
    the masked Hamming search (no host round trip); equals the numpy intersect1d mask and the host-mask search."""
    import torch

    from rag_snvbert_b200 import _lib
    from rag_snvbert_b200.index import pack_rows
    from rag_snvbert_b200.refdb import build_ref_db, intersect_windows, intersect_windows_device, sample_rows

    rng = np.random.default_rng(21)
    V, S_ref, S_tgt = 900, 120, 40
    ref_pos = np.sort(rng.choice(50000, size=V, replace=False)).astype(np.int64)
    tgt_pos = rng.permutation(np.concatenate([rng.choice(ref_pos, size=500, replace=False),
                                              rng.choice(50000, size=200)])).astype(np.int64)
    window_info = np.array([[0, 250], [250, 400], [400, 900]], dtype=np.int64)
    ref_gt = (rng.random((V, S_ref, 2)) < 0.3).astype(np.uint8)
    tgt_gt = (rng.random((V, S_tgt, 2)) < 0.3).astype(np.uint8)
    index = build_ref_db(ref_gt, window_info)
    d = index.d
    obs_host = intersect_windows(ref_pos, tgt_pos, window_info)            # [W, Lmax] per site
    obs_rows = np.zeros((3, d), dtype=np.uint8)
    obs_rows[:, : 2 * obs_host.shape[1]] = np.repeat(obs_host, 2, axis=1)  # both haplotypes share the site mask
    dev = torch.device("cuda", 0)
    masks = intersect_windows_device(torch.from_numpy(ref_pos).to(dev), torch.from_numpy(tgt_pos).to(dev),
                                     torch.from_numpy(window_info).to(dev), d, ploidy=2)
    exp = O.pack_bits_u32(obs_rows, _lib.packed_stride(d))
    np.testing.assert_array_equal(masks.cpu().numpy().view(np.uint32), exp)
    # drive the masked search with the device masks
    q = sample_rows(tgt_gt, window_info, d)
    qd = pack_rows(torch.from_numpy(q.reshape(-1, d)).to(dev), d).reshape(3, S_tgt, -1)
    Dd, Id = index.search(qd, 5, observed=masks)
    Dh, Ih = index.search(q, 5, observed=obs_rows)
    np.testing.assert_array_equal(Id.cpu().numpy(), Ih)
    np.testing.assert_array_equal(Dd.cpu().numpy(), Dh)


def test_packed_container_and_side_files_roundtrip(tmp_path):
    """the reference DB directory's side files (window_{i}_pos.npy, window_{i}_pop.npy: build_ref_db_intersect.py:73-75,
    build_ref_db_l2.py:80-83) and the bit-packed one-file container: same search results as the index it was saved from"""
    from rag_snvbert_b200 import refdb

    rng, ref, tgt, win = _data(9)
    V = ref.shape[0]
    ref_pos = np.sort(rng.choice(10 * V, V, replace=False))
    pop_labels = np.array(["AFR", "EUR", "EAS"])[rng.integers(0, 3, ref.shape[1])]
    for w, (a, b) in enumerate(win):
        np.save(tmp_path / f"window_{w}.npy", np.transpose(ref[a:b], (1, 0, 2)))
        np.save(tmp_path / f"window_{w}_pos.npy", ref_pos[a:b])
        np.save(tmp_path / f"window_{w}_pop.npy", pop_labels)
    index = refdb.load_ref_db(str(tmp_path), len(win))
    pos, pop = refdb.load_ref_db_meta(str(tmp_path), len(win))
    assert len(pos) == len(win) and all(np.array_equal(p, ref_pos[a:b]) for p, (a, b) in zip(pos, win))
    assert np.array_equal(pop.astype(str), pop_labels)
    n_sites = 2 * (win[:, 1] - win[:, 0])
    path = str(tmp_path / "panel.snvp")
    refdb.save_packed_db(path, index, n_sites=n_sites, positions=pos, pop=pop)
    raw = sum((tmp_path / f"window_{w}.npy").stat().st_size for w in range(len(win)))
    assert (tmp_path / "panel.snvp").stat().st_size < raw / 4   # 1 bit instead of 1 byte per allele (+ positions)
    again, ns2, pos2, pop2 = refdb.load_packed_db(path)
    assert again.ntotal == index.ntotal and again.d == index.d and np.array_equal(ns2, n_sites)
    assert all(np.array_equal(a, b) for a, b in zip(pos, pos2)) and np.array_equal(pop2, pop_labels)
    D, I = refdb.batch_search(index, tgt, win, 5)
    D2, I2 = refdb.batch_search(again, tgt, win, 5)
    np.testing.assert_array_equal(I2, I)
    np.testing.assert_array_equal(D2, D)
    # a directory written by build_ref_db_l2 has no position files
    (tmp_path / "l2").mkdir()
    np.save(tmp_path / "l2" / "window_0.npy", np.transpose(ref[0:10], (1, 0, 2)))
    np.save(tmp_path / "l2" / "window_0_pop.npy", pop_labels)
    pos3, pop3 = refdb.load_ref_db_meta(str(tmp_path / "l2"), 1)
    assert pos3 is None and np.array_equal(pop3.astype(str), pop_labels)
