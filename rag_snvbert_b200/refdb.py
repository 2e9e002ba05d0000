"""Offline reference-DB workflow on the GPU index (the callers on either side of the hot path).

Mirrors, without faiss / h5py / allel:
  * build_ref_db_l2.py:68-93       per window: slice -> [S, win_len, 2] -> rows [S, 2*win_len]
                                   (site-major, hap-minor interleave) -> IndexFlatL2.add
  * batch_test_faiss_l2.py:80-110  per window: target rows -> index.search(batch, top_k)
  * test_faiss_intersect.py:128-140 ref / target position intersection before the search
  * partial_faiss_intersect.py:46-80 expand_target_to_ref (target expanded to the ref site set
                                   with a missing mask) + per-sample observed-site search (:82-111)
The rows are 0/1 genotypes, so the whole sweep is ONE bit-packed Hamming search over all windows
(D equals faiss's squared L2 on these vectors).  Windows shorter than the longest are zero padded;
pad sites are equal in panel and queries and never contribute.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np

from .index import WindowedHammingIndex


def sample_rows(gt: np.ndarray, window_info: np.ndarray, d: Optional[int] = None) -> np.ndarray:
    """gt [V, S, 2] 0/1 (GT > 0 -> 1, build_ref_db_l2.py:50), window_info [W, 2] (start, end) ->
    uint8 rows [W, S, d] in the scripts' layout: row = sample, columns s0h0, s0h1, s1h0, ...
    (build_ref_db_l2.py:70-86), zero padded to d = 2 * max window length."""
    gt = (np.asarray(gt) > 0).astype(np.uint8)
    window_info = np.asarray(window_info)
    W = window_info.shape[0]
    dmax = int(2 * (window_info[:, 1] - window_info[:, 0]).max()) if W else 0
    d = dmax if d is None else int(d)
    if d < dmax:
        raise ValueError("d is smaller than the longest window")
    out = np.zeros((W, gt.shape[1], d), dtype=np.uint8)
    for w, (a, b) in enumerate(window_info):
        sub = np.transpose(gt[a:b], (1, 0, 2))  # [S, win_len, 2]
        out[w, :, : 2 * (b - a)] = sub.reshape(sub.shape[0], -1)
    return out


def build_ref_db(ref_gt: np.ndarray, window_info: np.ndarray, device: Optional[int] = None) -> WindowedHammingIndex:
    """build_ref_db_l2.py's loop: one windowed index instead of one .faiss file per window."""
    rows = sample_rows(ref_gt, window_info)
    index = WindowedHammingIndex(rows.shape[2], rows.shape[0], device)
    index.add(rows)
    return index


def load_ref_db(ref_db_dir: str, n_windows: int, device: Optional[int] = None) -> WindowedHammingIndex:
    """Reads the `window_{i}.npy` files ([S, win_len, 2]) written by build_ref_db_l2.py:77-78."""
    cubes = [np.load(os.path.join(ref_db_dir, f"window_{w}.npy")) for w in range(n_windows)]
    d = 2 * max(c.shape[1] for c in cubes)
    rows = np.zeros((n_windows, cubes[0].shape[0], d), dtype=np.uint8)
    for w, c in enumerate(cubes):
        rows[w, :, : 2 * c.shape[1]] = (c > 0).reshape(c.shape[0], -1)
    index = WindowedHammingIndex(d, n_windows, device)
    index.add(rows)
    return index


def load_ref_db_meta(ref_db_dir: str, n_windows: int):
    """The side files of the reference DB directory: `window_{i}_pos.npy` (site positions of the window, written by
    build_ref_db_intersect.py:73-74) and `window_{i}_pop.npy` (population label per reference sample,
    build_ref_db_l2.py:80-83 / build_ref_db_intersect.py:75).  Returns (pos, pop): pos = list of int64 arrays or None
    when the directory has no position files (the L2 builder does not write them); pop = array of labels (the same for
    every window: it is the panel's sample list) or None."""
    pos, pop = [], None
    for w in range(n_windows):
        pp = os.path.join(ref_db_dir, f"window_{w}_pos.npy")
        if not os.path.exists(pp):
            pos = None
            break
        pos.append(np.load(pp).astype(np.int64))
    lp = os.path.join(ref_db_dir, "window_0_pop.npy")
    if os.path.exists(lp):
        pop = np.load(lp, allow_pickle=True)
    return pos, pop


# ---- bit-packed panel container ------------------------------------------------------------------------------
# The .npy cubes cost 1 byte per allele (and the .faiss files 4): a 5,008-sample x 1,030-site window is 10.3 MB as
# window_{i}.npy, 41 MB as window_{i}.faiss and 0.36 MB packed.  `save_packed_db` writes the index's own device layout
# (uint32 words, site s -> word s / 32, bit s % 32, row stride snv_packed_stride(d)) for ALL windows into one file, so a
# later run adds it with a straight copy - no 0/1 expansion, no re-packing.  Layout (little endian):
#   magic "SNVP" | version u32 = 1 | n_windows u32 | rows i64 | d i64 | stride u32 | has_pos u8 | has_pop u8 | pad u16 |
#   n_sites i32 [n_windows] | panel u32 [n_windows][rows][stride] |
#   (has_pos) per window: count i64, positions i64 [count] | (has_pop) npy-serialised label array
_PACKED_MAGIC = b"SNVP"


def save_packed_db(path: str, index: WindowedHammingIndex, n_sites=None, positions=None, pop=None) -> None:
    """index -> one packed container file.  n_sites: real site count per window (default d), positions / pop as returned
    by load_ref_db_meta."""
    import io
    import struct

    W, n, d, stride = index.n_windows, index.ntotal, index.d, index.stride
    ns = np.full(W, d, np.int32) if n_sites is None else np.ascontiguousarray(np.asarray(n_sites, dtype=np.int32))
    if ns.shape != (W,):
        raise ValueError("n_sites must have one entry per window")
    with open(path, "wb") as f:
        f.write(_PACKED_MAGIC)
        f.write(struct.pack("<IIqqIBBH", 1, W, n, d, stride, 1 if positions is not None else 0, 1 if pop is not None else 0, 0))
        f.write(ns.astype("<i4").tobytes())
        for w in range(W):
            f.write(index.export_packed(w).astype("<u4").tobytes())
        if positions is not None:
            if len(positions) != W:
                raise ValueError("positions must have one array per window")
            for pw in positions:
                pw = np.ascontiguousarray(np.asarray(pw, dtype="<i8"))
                f.write(struct.pack("<q", pw.size))
                f.write(pw.tobytes())
        if pop is not None:
            buf = io.BytesIO()
            np.save(buf, np.asarray(pop).astype(str), allow_pickle=False)
            f.write(buf.getvalue())


def load_packed_db(path: str, device: Optional[int] = None):
    """-> (index, n_sites int32 [W], positions list | None, pop array | None).  The panel is added as packed rows: the
    bytes of the file go to the device as they are."""
    import io
    import struct

    with open(path, "rb") as f:
        if f.read(4) != _PACKED_MAGIC:
            raise ValueError(f"{path}: not a packed reference DB (bad magic)")
        ver, W, n, d, stride, has_pos, has_pop, _ = struct.unpack("<IIqqIBBH", f.read(4 + 4 + 8 + 8 + 4 + 1 + 1 + 2))
        if ver != 1:
            raise ValueError(f"{path}: unsupported container version {ver}")
        ns = np.frombuffer(f.read(4 * W), dtype="<i4").astype(np.int32)
        index = WindowedHammingIndex(int(d), int(W), device)
        if stride != index.stride:
            raise ValueError(f"{path}: row stride {stride} does not match this library's {index.stride} for d = {d}")
        panel = np.frombuffer(f.read(4 * W * n * stride), dtype="<u4")
        if panel.size != W * n * stride:
            raise ValueError(f"{path}: truncated panel")
        if n:
            index.add(panel.reshape(W, n, stride))
        positions = None
        if has_pos:
            positions = []
            for _w in range(W):
                (cnt,) = struct.unpack("<q", f.read(8))
                positions.append(np.frombuffer(f.read(8 * cnt), dtype="<i8").astype(np.int64))
        pop = np.load(io.BytesIO(f.read()), allow_pickle=False) if has_pop else None
    return index, ns, positions, pop


def batch_search(index: WindowedHammingIndex, target_gt: np.ndarray, window_info: np.ndarray, top_k: int,
                 samples=None) -> Tuple[np.ndarray, np.ndarray]:
    """batch_test_faiss_l2.py:80-110 for every window at once -> D float32 [W, nq, k], I int64."""
    rows = sample_rows(target_gt, window_info, index.d)
    if samples is not None:
        rows = np.ascontiguousarray(rows[:, list(samples)])
    return index.search(rows, top_k, dist_dtype=np.float32)


def expand_target_to_ref(ref_pos: np.ndarray, tgt_data: np.ndarray, tgt_pos: np.ndarray):
    """partial_faiss_intersect.py:46-80, vectorised: target genotypes placed on the ref site set.
    Returns (expanded [var_ref, S_t, 2] uint8, missing_mask [var_ref, S_t] uint8, 1 = missing).
    Duplicate target positions resolve to the LAST occurrence like the reference's dict."""
    ref_pos = np.asarray(ref_pos)
    tgt_pos = np.asarray(tgt_pos)
    tgt_data = np.asarray(tgt_data)
    order = np.argsort(tgt_pos, kind="stable")
    sp = tgt_pos[order]
    j = np.searchsorted(sp, ref_pos, side="right") - 1  # last occurrence <= p
    hit = (j >= 0) & (sp[np.clip(j, 0, None)] == ref_pos)
    expanded = np.zeros((ref_pos.shape[0],) + tgt_data.shape[1:], dtype=np.uint8)
    expanded[hit] = tgt_data[order[j[hit]]]
    missing = np.zeros((ref_pos.shape[0], tgt_data.shape[1]), dtype=np.uint8)
    missing[~hit] = 1
    return expanded, missing


def intersect_windows(ref_pos: np.ndarray, tgt_pos: np.ndarray, window_info: np.ndarray) -> np.ndarray:
    """test_faiss_intersect.py:128-140 as a mask: observed[w, s] = 1 where ref site
    window_info[w,0]+s is also a target position (np.intersect1d semantics)."""
    ref_pos = np.asarray(ref_pos)
    window_info = np.asarray(window_info)
    shared = np.isin(ref_pos, np.asarray(tgt_pos))
    L = int((window_info[:, 1] - window_info[:, 0]).max())
    obs = np.zeros((window_info.shape[0], L), dtype=np.uint8)
    for w, (a, b) in enumerate(window_info):
        obs[w, : b - a] = shared[a:b]
    return obs


def intersect_windows_device(ref_pos, tgt_pos, window_info, d: int, ploidy: int = 2):
    """intersect_windows on the device (snv_intersect_masks): CUDA int64 tensors in, packed per-window observed
    masks [W, snv_packed_stride(d)] out - what `index.search(packed_queries, k, observed=...)` takes, so the
    intersect workflow (test_faiss_intersect.py:128-181) needs no host round trip per window.  ploidy 2 = the
    scripts' sample rows (s0h0, s0h1, ...), 1 = haplotype rows."""
    from .index import intersect_masks

    return intersect_masks(ref_pos, tgt_pos, window_info, d, ploidy)


def partial_search(index: WindowedHammingIndex, expanded: np.ndarray, missing: np.ndarray,
                   window_info: np.ndarray, top_k: int):
    """Observed-site search of every target sample in every window
    (partial_faiss_intersect.py:145-172 with build_partial_index_l2 :82-111): distance restricted to
    sites where the sample's mask == 0, both haplotypes of a sample share its site mask.
    expanded [var_ref, S_t, 2], missing [var_ref, S_t] -> D float32 [W, S_t, k], I int64.
    (The script concatenates the query as [h1.., h2..] but the panel rows as s0h0,s0h1,..
    (:94 vs :101) — a layout slip; this follows the evident intent: aligned columns.)"""
    q = sample_rows(expanded, window_info, index.d)
    m2 = np.repeat(np.asarray(missing)[:, :, None], 2, axis=2)  # both haplotypes share the site mask
    miss = sample_rows(m2, window_info, index.d)
    return index.search(q, top_k, missing=miss, dist_dtype=np.float32)
