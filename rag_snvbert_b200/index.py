"""Index objects over the C ABI: the host-side mirror of the faiss surface the reference calls.

Reference call sites replaced (paths relative to the reference tree):
  faiss.IndexFlatL2(d).add / .search      build_ref_db_l2.py:89-90, batch_test_faiss_l2.py:110,
                                          src/dataset/rag_train_dataset.py:132-134,281
  faiss.IndexBinaryFlat(d).add / .search  test_faiss_intersect.py:173-181
  per-query observed-site search          partial_faiss_intersect.py:82-111
  torch.cdist + topk                      src/dataset/embedding_rag_dataset.py:392-402
  gather of retrieved haplotypes          src/dataset/rag_train_dataset.py:287-307,
                                          src/dataset/embedding_rag_dataset.py:406-438

Inputs may be numpy arrays (host; copied by the library on the call's stream) or torch CUDA
tensors (used in place, zero copy, on torch's current stream).  Outputs come back in the same
world as the queries.  No CPU fallback exists: without the CUDA library / device these raise.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import numpy as np

from . import _lib as L

try:  # torch is plumbing only (device memory + streams); numpy-only use works without it
    import torch
except Exception:  # pragma: no cover
    torch = None


MAX_K = 32  # neighbours the kernels select in registers; larger k takes the block path (_search_large_k)


def _is_torch(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


class _Arg:
    """A contiguous host (numpy) or device (torch.cuda) array ready to cross the C ABI."""

    __slots__ = ("arr", "ptr", "on_device", "shape", "np_dtype")

    def __init__(self, x):
        if _is_torch(x):
            if x.is_cuda:
                x = x.contiguous()
                self.arr, self.ptr, self.on_device = x, x.data_ptr(), True
                self.shape = tuple(x.shape)
                self.np_dtype = np.dtype(str(x.dtype).replace("torch.", "").replace("bool", "bool_"))
                return
            x = x.detach().numpy()
        x = np.ascontiguousarray(x)
        self.arr, self.ptr, self.on_device = x, x.ctypes.data, False
        self.shape = x.shape
        self.np_dtype = x.dtype


def _current_stream(device: int) -> int:
    if torch is not None and torch.cuda.is_available():
        return int(torch.cuda.current_stream(device).cuda_stream)
    return 0


def _hamming_dtype(a: _Arg, d: int, stride: int, what: str, codes: bool = False) -> Tuple[int, "_Arg"]:
    dt = a.np_dtype
    last = a.shape[-1]
    if codes:  # np.packbits rows, the input of faiss.IndexBinaryFlat (test_faiss_intersect.py:46-54)
        if dt != np.dtype(np.uint8) or last != (d + 7) // 8:
            raise ValueError(f"{what}: binary codes must be uint8 [.., {(d + 7) // 8}]")
        return L.DT_PACKED_U8, a
    if dt in (np.dtype(np.uint32), np.dtype(np.int32)):
        words = (d + 31) // 32
        if last == stride:
            return L.DT_PACKED_U32, a
        if last == words:  # dense rows without the stride padding: the compact wire format (re-strided on the device)
            return L.DT_PACKED_U32_DENSE, a
        raise ValueError(f"{what}: packed rows must have {stride} uint32 words (snv_packed_stride({d})) or, dense, {words}; got {last}")
    if last != d:
        raise ValueError(f"{what}: last dimension must be d={d}, got {last}")
    if dt == np.dtype(np.float32):
        return L.DT_F32, a
    if dt == np.dtype(np.int64):
        return L.DT_I64_TOKENS, a
    if dt in (np.dtype(np.uint8), np.dtype(np.bool_), np.dtype(np.int8)):
        if dt != np.dtype(np.uint8):
            arr = a.arr.view(torch.uint8) if _is_torch(a.arr) else a.arr.view(np.uint8)
            a = _Arg(arr)
        return L.DT_U8, a
    raise ValueError(f"{what}: unsupported dtype {dt} (use uint8/bool 0-1 sites, float32, packed uint32 or int64 tokens)")


class _IndexBase:
    kind = L.KIND_HAMMING

    def __init__(self, d: int, n_windows: int = 1, device: Optional[int] = None, l2_mode: int = L.L2_TF32X3):
        if int(d) <= 0:
            raise ValueError("d must be positive")
        lib = L.lib()
        if device is None:
            device = torch.cuda.current_device() if (torch is not None and torch.cuda.is_available()) else 0
        h = ctypes.c_void_p()
        L.check(lib.snv_index_create(self.kind, int(d), int(n_windows), int(device), int(l2_mode), ctypes.byref(h)),
                "snv_index_create")
        self._h = h
        self._lib = lib
        self.d = int(d)
        self.n_windows = int(n_windows)
        self.device = int(device)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self._lib.snv_index_free(h)
            except Exception:
                pass
            self._h = None

    @property
    def ntotal(self) -> int:
        return int(self._lib.snv_index_ntotal(self._h))

    def reset(self) -> None:
        L.check(self._lib.snv_index_reset(self._h), "snv_index_reset")

    # ---- helpers
    def _win_shape(self, a: _Arg, what: str):
        """[rows, last] for a single-window index or [W, rows, last] -> (W_in, rows)."""
        if len(a.shape) == 2:
            if self.n_windows != 1:
                raise ValueError(f"{what}: a {self.n_windows}-window index takes [windows, rows, ...] arrays")
            return 1, a.shape[0]
        if len(a.shape) == 3:
            return a.shape[0], a.shape[1]
        raise ValueError(f"{what}: expected a 2-D or 3-D array, got shape {a.shape}")

    def _alloc_out(self, like_device: bool, shape, np_dtype):
        if like_device:
            t = torch.empty(shape, dtype=getattr(torch, np.dtype(np_dtype).name), device=f"cuda:{self.device}")
            return t, t.data_ptr()
        arr = np.empty(shape, dtype=np_dtype)
        return arr, arr.ctypes.data


class WindowedHammingIndex(_IndexBase):
    """`n_windows` independent bit-packed haplotype panels of `n_sites` sites each, searched in one
    launch.  Distances are integer Hamming distances (== faiss squared L2 on 0/1 rows); with a
    mask, Hamming over the observed sites of each query (partial_faiss_intersect.py:82-111)."""

    kind = L.KIND_HAMMING

    def __init__(self, n_sites: int, n_windows: int = 1, device: Optional[int] = None):
        super().__init__(n_sites, n_windows, device)
        self.stride = L.packed_stride(n_sites)
        self.words = L.packed_words(n_sites)

    def add(self, x, codes: bool = False) -> None:
        """x: [W, n, d] (or [n, d] when W == 1) 0/1 uint8/bool/float32, int64 tokens, or packed
        uint32 [.., stride]; codes=True: np.packbits bytes [.., ceil(d/8)].  Appends n rows to
        every window."""
        a = _Arg(x)
        W, n = self._win_shape(a, "add")
        if W != self.n_windows:
            raise ValueError(f"add: array has {W} windows, index has {self.n_windows}")
        dt, a = _hamming_dtype(a, self.d, self.stride, "add", codes)
        flags = L.X_ON_DEVICE if a.on_device else 0
        L.check(self._lib.snv_index_add(self._h, a.ptr, n, dt, flags, _current_stream(self.device)), "snv_index_add")

    def search(self, q, k: int, observed=None, missing=None, w0: int = 0, id_offset: int = 0,
               dist_dtype=np.int32, codes: bool = False, out=None):
        """q: [nw, nq, d] (or [nq, d] for one window) in any add() dtype.  `observed` / `missing`:
        optional site mask, [nw, nq, d] per query or [nw, d] per window ([nq, d] / [d] for one
        window).  Returns (D [nw, nq, k] int32 or float32, I [nw, nq, k] int64); single-window
        calls drop the leading axis.  `out=(D, I)`: caller-owned result buffers (e.g. pinned host
        arrays, so the device-to-host copies of a host-buffer call overlap the scan)."""
        if int(k) < 1:
            raise ValueError("k must be >= 1")
        if int(k) > MAX_K:
            if out is not None or codes:
                raise ValueError(f"search: k > {MAX_K} takes the block path, which allocates its own results (no out=, no codes=)")
            return self._search_large_k(q, int(k), observed, missing, int(w0), int(id_offset), dist_dtype)
        a = _Arg(q)
        squeeze = len(a.shape) == 2
        if squeeze:
            nw, nq = 1, a.shape[0]
        elif len(a.shape) == 3:
            nw, nq = a.shape[0], a.shape[1]
        else:
            raise ValueError(f"search: expected a 2-D or 3-D query array, got shape {a.shape}")
        dt, a = _hamming_dtype(a, self.d, self.stride, "search", codes)
        flags = L.Q_ON_DEVICE | L.OUT_ON_DEVICE if a.on_device else 0
        m_ptr, m_mode = None, L.MASK_NONE
        if observed is not None and missing is not None:
            raise ValueError("search: give observed= or missing=, not both")
        mk = observed if observed is not None else missing
        if mk is not None:
            m = _Arg(mk)
            if m.on_device != a.on_device:
                raise ValueError("search: queries and mask must live in the same memory (both numpy or both CUDA)")
            mshape = m.shape
            per_query_shape = a.shape
            if tuple(mshape) == tuple(per_query_shape):
                m_mode = L.MASK_PER_QUERY
            elif tuple(mshape) == tuple(per_query_shape[:-2] + per_query_shape[-1:]):
                m_mode = L.MASK_PER_WINDOW
            else:
                raise ValueError(f"search: mask shape {mshape} matches neither queries {per_query_shape} nor one row per window")
            mdt, m = _hamming_dtype(m, self.d, self.stride, "search(mask)")
            if mdt != dt:
                raise ValueError("search: mask and queries must use the same dtype")
            m_ptr = m.ptr
            if missing is not None:
                flags |= L.MASK_IS_MISSING
        want_f = np.dtype(dist_dtype) == np.dtype(np.float32)
        shape = (nw, nq, int(k))
        if out is not None:
            Do, Io = _Arg(out[0]), _Arg(out[1])
            want_f = Do.np_dtype == np.dtype(np.float32)
            if (Do.on_device != a.on_device or Io.on_device != a.on_device or int(np.prod(Do.shape)) != int(np.prod(shape))
                    or int(np.prod(Io.shape)) != int(np.prod(shape)) or Io.np_dtype != np.dtype(np.int64)
                    or Do.np_dtype not in (np.dtype(np.int32), np.dtype(np.float32))
                    or Do.arr is not out[0] or Io.arr is not out[1]):
                raise ValueError("search: out=(D, I) must be contiguous int32|float32 / int64 arrays of the result size, "
                                 "in the same memory as the queries")
            D, Dp, I, Ip = out[0].reshape(shape), Do.ptr, out[1].reshape(shape), Io.ptr
        else:
            D, Dp = self._alloc_out(a.on_device, shape, np.float32 if want_f else np.int32)
            I, Ip = self._alloc_out(a.on_device, shape, np.int64)
        L.check(self._lib.snv_index_search(self._h, int(w0), nw, a.ptr, nq, dt, m_ptr, m_mode, int(k), int(id_offset),
                                           None if want_f else Dp, Dp if want_f else None, Ip, flags,
                                           _current_stream(self.device)), "snv_index_search")
        if squeeze:
            return D[0], I[0]
        return D, I

    # ---- k > 32 (faiss accepts any k; the kernels select at most 32 neighbours in registers)
    def _packed_on_device(self, x, what):
        """any accepted row dtype (numpy or CUDA tensor) -> packed int32 CUDA tensor [..., stride]"""
        a = _Arg(x)
        t = a.arr if a.on_device else torch.from_numpy(np.ascontiguousarray(a.arr)).to(f"cuda:{self.device}")
        dt, a2 = _hamming_dtype(_Arg(t), self.d, self.stride, what)
        t = a2.arr
        if dt == L.DT_PACKED_U32:
            return t.view(torch.int32) if t.dtype != torch.int32 else t
        if dt == L.DT_PACKED_U32_DENSE:
            pad = torch.zeros(t.shape[:-1] + (self.stride - t.shape[-1],), dtype=t.dtype, device=t.device)
            return torch.cat([t, pad], dim=-1).view(torch.int32)
        flat = pack_rows(t.reshape(-1, t.shape[-1]), self.d)
        return flat.reshape(t.shape[:-1] + (self.stride,))

    def _search_large_k(self, q, k, observed, missing, w0, id_offset, dist_dtype):
        """Exact top-k for k > 32 by the BLOCK path: the window's panel is re-added as ntotal / 32 windows of 32 rows, one
        launch searches every block with k = 32 - which returns ALL of a block's distances - and the k best of the
        (distance, id) keys are selected with one torch.topk.  Same total order, same padding as the in-kernel path; far
        slower per pair (every query meets every 32-row block as its own tiny window), which is acceptable for the rare
        large k (the reference asks for k <= 5)."""
        if torch is None or not torch.cuda.is_available():
            raise RuntimeError("search with k > 32 needs torch with CUDA")
        was_numpy = not _Arg(q).on_device
        qp = self._packed_on_device(q, "search")
        qdim = qp.dim()
        squeeze = qdim == 2
        if squeeze:
            qp = qp.unsqueeze(0)
        nw, nq = int(qp.shape[0]), int(qp.shape[1])
        mk = observed if observed is not None else missing
        if observed is not None and missing is not None:
            raise ValueError("search: give observed= or missing=, not both")
        mp = None
        if mk is not None:
            mp = self._packed_on_device(mk, "search(mask)")
            if mp.dim() == qdim:                      # one mask row per query
                mp = mp.unsqueeze(0) if squeeze else mp
            elif mp.dim() == qdim - 1:                # one mask row per window
                mp = mp.reshape(nw, 1, self.stride).expand(nw, nq, self.stride)
            else:
                raise ValueError("search: mask shape matches neither the queries nor one row per window")
            mp = mp.contiguous()
        n, B = self.ntotal, MAX_K
        nb = max(1, -(-n // B))
        dev = qp.device
        Ds, Is = [], []
        for w in range(nw):
            rows = torch.from_numpy(self.export_packed(w0 + w).view(np.int32)).to(dev)        # [n, stride]
            pad = torch.zeros((nb * B - n, self.stride), dtype=torch.int32, device=dev)
            blocks = WindowedHammingIndex(self.d, nb, self.device)
            blocks.add(torch.cat([rows, pad]).reshape(nb, B, self.stride))
            qrep = qp[w].unsqueeze(0).expand(nb, nq, self.stride).contiguous()
            kw = {}
            if mp is not None:
                kw["observed" if observed is not None else "missing"] = mp[w].unsqueeze(0).expand(nb, nq, self.stride).contiguous()
            Db, Ib = blocks.search(qrep, B, **kw)                                             # [nb, nq, 32]: every distance of a block
            gid = Ib + (torch.arange(nb, device=dev) * B).view(nb, 1, 1)
            key = (Db.to(torch.int64) << 32) | gid
            key = torch.where((Ib < 0) | (gid >= n), torch.full_like(key, torch.iinfo(torch.int64).max), key)
            key = key.permute(1, 0, 2).reshape(nq, nb * B)
            kk = min(k, nb * B)
            best = torch.topk(key, kk, dim=1, largest=False, sorted=True).values
            if kk < k:
                best = torch.cat([best, torch.full((nq, k - kk), torch.iinfo(torch.int64).max, dtype=torch.int64, device=dev)], dim=1)
            none = best == torch.iinfo(torch.int64).max
            Ds.append(torch.where(none, torch.full_like(best, 0x7FFFFFFF), best >> 32).to(torch.int32))
            Is.append(torch.where(none, torch.full_like(best, -1), (best & 0xFFFFFFFF) + id_offset))
        D, I = torch.stack(Ds), torch.stack(Is)
        if np.dtype(dist_dtype) == np.dtype(np.float32):
            D = torch.where(I < 0, torch.full_like(D, 0, dtype=torch.float32) + 3.4028234663852886e38, D.to(torch.float32))
        if squeeze:
            D, I = D[0], I[0]
        if was_numpy:
            return D.cpu().numpy(), I.cpu().numpy()
        return D, I

    def search_compact(self, q, k: int, observed=None, missing=None, w0: int = 0, out=None):
        """search() with compact results: (D uint16 [nw, nq, k], I int32 [nw, nq, k]) - the same distances and row ids in
        half the bytes (6 instead of 12 per neighbour), for host-buffer sweeps whose cost is the PCIe return trip
        (batch_test_faiss_l2.py:109-111 keeps every window's D, I).  Padding: I = -1, D = 0xFFFF.  `out=(D, I)`:
        caller-owned (e.g. pinned) buffers of those dtypes."""
        if int(k) < 1:
            raise ValueError("k must be >= 1")
        a = _Arg(q)
        squeeze = len(a.shape) == 2
        if squeeze:
            nw, nq = 1, a.shape[0]
        elif len(a.shape) == 3:
            nw, nq = a.shape[0], a.shape[1]
        else:
            raise ValueError(f"search_compact: expected a 2-D or 3-D query array, got shape {a.shape}")
        dt, a = _hamming_dtype(a, self.d, self.stride, "search_compact")
        flags = L.Q_ON_DEVICE | L.OUT_ON_DEVICE if a.on_device else 0
        m_ptr, m_mode = None, L.MASK_NONE
        if observed is not None and missing is not None:
            raise ValueError("search_compact: give observed= or missing=, not both")
        mk = observed if observed is not None else missing
        if mk is not None:
            m = _Arg(mk)
            if m.on_device != a.on_device:
                raise ValueError("search_compact: queries and mask must live in the same memory (both numpy or both CUDA)")
            if tuple(m.shape) == tuple(a.shape):
                m_mode = L.MASK_PER_QUERY
            elif tuple(m.shape) == tuple(a.shape[:-2] + a.shape[-1:]):
                m_mode = L.MASK_PER_WINDOW
            else:
                raise ValueError(f"search_compact: mask shape {m.shape} matches neither the queries nor one row per window")
            mdt, m = _hamming_dtype(m, self.d, self.stride, "search_compact(mask)")
            if mdt != dt:
                raise ValueError("search_compact: mask and queries must use the same dtype")
            m_ptr = m.ptr
            if missing is not None:
                flags |= L.MASK_IS_MISSING
        shape = (nw, nq, int(k))
        if out is not None:
            Do, Io = _Arg(out[0]), _Arg(out[1])
            if (Do.on_device != a.on_device or Io.on_device != a.on_device or Do.np_dtype not in (np.dtype(np.uint16), np.dtype(np.int16))
                    or Io.np_dtype != np.dtype(np.int32) or int(np.prod(Do.shape)) != int(np.prod(shape))
                    or int(np.prod(Io.shape)) != int(np.prod(shape)) or Do.arr is not out[0] or Io.arr is not out[1]):
                raise ValueError("search_compact: out=(D, I) must be contiguous uint16 / int32 arrays of the result size, "
                                 "in the same memory as the queries")
            D, Dp, I, Ip = out[0].reshape(shape), Do.ptr, out[1].reshape(shape), Io.ptr
        else:
            if a.on_device:
                D = torch.empty(shape, dtype=torch.int16, device=f"cuda:{self.device}")  # uint16 bit patterns
                Dp = D.data_ptr()
            else:
                D = np.empty(shape, dtype=np.uint16)
                Dp = D.ctypes.data
            I, Ip = self._alloc_out(a.on_device, shape, np.int32)
        L.check(self._lib.snv_index_search_compact(self._h, int(w0), nw, a.ptr, nq, dt, m_ptr, m_mode, int(k), Dp, Ip, flags,
                                                   _current_stream(self.device)), "snv_index_search_compact")
        if squeeze:
            return D[0], I[0]
        return D, I

    def search_grouped(self, q, window_ids, k: int, observed=None, missing=None, dist_dtype=np.int32):
        """Ragged per-window batch (a training batch regrouped by window_idx,
        src/dataset/rag_train_dataset.py:239-281): q [nq_total, d] in caller order, window_ids
        int [nq_total].  One launch; results [nq_total, k] in caller order."""
        if int(k) < 1:
            raise ValueError("k must be >= 1")
        a = _Arg(q)
        if len(a.shape) != 2:
            raise ValueError("search_grouped: expected [nq_total, d] queries")
        nq = a.shape[0]
        wid = np.ascontiguousarray(np.asarray(window_ids.cpu() if _is_torch(window_ids) else window_ids), dtype=np.int32)
        if wid.shape != (nq,):
            raise ValueError("search_grouped: window_ids must have one entry per query")
        dt, a = _hamming_dtype(a, self.d, self.stride, "search_grouped")
        flags = L.Q_ON_DEVICE | L.OUT_ON_DEVICE if a.on_device else 0
        if observed is not None and missing is not None:
            raise ValueError("search_grouped: give observed= or missing=, not both")
        mk = observed if observed is not None else missing
        m_ptr = None
        if mk is not None:
            m = _Arg(mk)
            if m.on_device != a.on_device or tuple(m.shape) != tuple(a.shape):
                raise ValueError("search_grouped: mask must match the queries (shape and memory)")
            mdt, m = _hamming_dtype(m, self.d, self.stride, "search_grouped(mask)")
            if mdt != dt:
                raise ValueError("search_grouped: mask and queries must use the same dtype")
            m_ptr = m.ptr
            if missing is not None:
                flags |= L.MASK_IS_MISSING
        want_f = np.dtype(dist_dtype) == np.dtype(np.float32)
        D, Dp = self._alloc_out(a.on_device, (nq, int(k)), np.float32 if want_f else np.int32)
        I, Ip = self._alloc_out(a.on_device, (nq, int(k)), np.int64)
        L.check(self._lib.snv_index_search_grouped(self._h, a.ptr, wid.ctypes.data, nq, dt, m_ptr, int(k),
                                                   None if want_f else Dp, Dp if want_f else None, Ip, flags,
                                                   _current_stream(self.device)), "snv_index_search_grouped")
        return D, I

    def gather_tokens_grouped(self, I, window_ids, n_sites=None, seq_len: int = 1030):
        """I [nq_total, k] + window_ids [nq_total] -> int64 tokens [nq_total, k, seq_len]."""
        a = _Arg(I)
        if a.np_dtype != np.dtype(np.int64) or len(a.shape) != 2:
            raise ValueError("gather_tokens_grouped: I must be int64 [nq_total, k]")
        nq, k = a.shape
        wid = np.ascontiguousarray(np.asarray(window_ids.cpu() if _is_torch(window_ids) else window_ids), dtype=np.int32)
        ns_ptr = None
        if n_sites is not None:
            ns = np.ascontiguousarray(np.broadcast_to(np.asarray(n_sites, dtype=np.int32), (self.n_windows,)))
            ns_ptr = ns.ctypes.data
        flags = L.Q_ON_DEVICE | L.OUT_ON_DEVICE if a.on_device else 0
        out, op = self._alloc_out(a.on_device, (nq, k, int(seq_len)), np.int64)
        L.check(self._lib.snv_index_gather_tokens_grouped(self._h, a.ptr, wid.ctypes.data, nq, k, ns_ptr, int(seq_len),
                                                          op, flags, _current_stream(self.device)),
                "snv_index_gather_tokens_grouped")
        return out

    def gather_tokens(self, I, n_sites=None, seq_len: int = 1030, w0: int = 0):
        """I [nw, nq, k] (or [nq, k]) -> int64 tokens [.., k, seq_len] in the model's input layout
        (src/dataset/rag_train_dataset.py:287-307): [SOS] + 5|6 per site + [EOS] + PAD."""
        a = _Arg(I)
        if a.np_dtype != np.dtype(np.int64):
            raise ValueError("gather_tokens: I must be int64")
        squeeze = len(a.shape) == 2
        nw, nq, k = (1,) + tuple(a.shape) if squeeze else tuple(a.shape)
        ns_ptr = None
        if n_sites is not None:
            ns = np.ascontiguousarray(np.broadcast_to(np.asarray(n_sites, dtype=np.int32), (nw,)))
            ns_ptr = ns.ctypes.data
        flags = L.Q_ON_DEVICE | L.OUT_ON_DEVICE if a.on_device else 0
        out, op = self._alloc_out(a.on_device, (nw, nq, k, int(seq_len)), np.int64)
        L.check(self._lib.snv_index_gather_tokens(self._h, int(w0), nw, a.ptr, nq, k, ns_ptr, int(seq_len), op, flags,
                                                  _current_stream(self.device)), "snv_index_gather_tokens")
        return out[0] if squeeze else out

    def export_packed(self, window: int = 0) -> np.ndarray:
        out = np.zeros((self.ntotal, self.stride), dtype=np.uint32)
        L.check(self._lib.snv_index_export(self._h, int(window), out.ctypes.data), "snv_index_export")
        return out


class IndexHamming(WindowedHammingIndex):
    """One window: the native replacement of IndexFlatL2 / IndexBinaryFlat on 0/1 haplotypes."""

    def __init__(self, n_sites: int, device: Optional[int] = None):
        super().__init__(n_sites, 1, device)


class WindowedL2Index(_IndexBase):
    """Float rows, squared L2 = |q|^2 + |r|^2 - 2 q.r with the cross term on tcgen05 (tf32 operands;
    precision 'tf32x3' = hi/lo split, fp32-faithful; 'tf32' = one pass, exact for small-integer
    inputs such as the V17 token vectors)."""

    kind = L.KIND_L2

    def __init__(self, d: int, n_windows: int = 1, device: Optional[int] = None, precision: str = "tf32x3",
                 center: Optional[bool] = None):
        """center=True subtracts the column means of the first rows added from panel and queries
        (squared L2 is translation invariant): embedding vectors share a large position /
        allele-frequency component, and without it |q|^2 + |r|^2 - 2 q.r cancels badly at depths like
        the reference's L*D = 197,760.  Integer-valued rows (tokens, genotypes) are exact as they are
        and must not be centred.  center=None (default) decides on the first add(): centred unless
        every value is a small integer."""
        modes = {"tf32": L.L2_TF32, "tf32x3": L.L2_TF32X3}
        if precision not in modes:
            raise ValueError("precision must be 'tf32' or 'tf32x3'")
        flag = L.L2_CENTER_AUTO if center is None else (L.L2_CENTER if center else 0)
        super().__init__(d, n_windows, device, modes[precision] | flag)
        self.precision = precision
        self.center = center

    def add(self, x) -> None:
        a = _Arg(x)
        if a.np_dtype != np.dtype(np.float32):
            a = _Arg(a.arr.float() if _is_torch(a.arr) else a.arr.astype(np.float32))
        W, n = self._win_shape(a, "add")
        if W != self.n_windows:
            raise ValueError(f"add: array has {W} windows, index has {self.n_windows}")
        if a.shape[-1] != self.d:
            raise ValueError(f"add: last dimension must be d={self.d}, got {a.shape[-1]}")
        flags = L.X_ON_DEVICE if a.on_device else 0
        L.check(self._lib.snv_index_add(self._h, a.ptr, n, L.DT_F32, flags, _current_stream(self.device)), "snv_index_add")

    def _search_large_k(self, q, k, w0, id_offset):
        """k > 32 by the block path (see WindowedHammingIndex._search_large_k): the window's rows re-added as blocks of 32, one
        launch with k = 32 returns every distance of every block, torch.topk on (float bits << 32 | id) keys selects."""
        if torch is None or not torch.cuda.is_available():
            raise RuntimeError("search with k > 32 needs torch with CUDA")
        a = _Arg(q)
        was_numpy = not a.on_device
        qt = a.arr if a.on_device else torch.from_numpy(np.ascontiguousarray(a.arr)).to(f"cuda:{self.device}")
        qt = qt.float()
        squeeze = qt.dim() == 2
        if squeeze:
            qt = qt.unsqueeze(0)
        nw, nq = int(qt.shape[0]), int(qt.shape[1])
        n, B = self.ntotal, MAX_K
        nb = max(1, -(-n // B))
        dev = qt.device
        Ds, Is = [], []
        for w in range(nw):
            rows = torch.from_numpy(self.export_rows(w0 + w)).to(dev)
            pad = torch.zeros((nb * B - n, self.d), dtype=torch.float32, device=dev)
            blocks = WindowedL2Index(self.d, nb, self.device, self.precision, center=self.center)
            blocks.add(torch.cat([rows, pad]).reshape(nb, B, self.d))
            Db, Ib = blocks.search(qt[w].unsqueeze(0).expand(nb, nq, self.d).contiguous(), B)
            gid = Ib + (torch.arange(nb, device=dev) * B).view(nb, 1, 1)
            key = (Db.clamp_min(0).view(torch.int32).to(torch.int64) << 32) | gid       # non-negative float bits order like the floats
            key = torch.where((Ib < 0) | (gid >= n), torch.full_like(key, torch.iinfo(torch.int64).max), key)
            key = key.permute(1, 0, 2).reshape(nq, nb * B)
            kk = min(k, nb * B)
            best = torch.topk(key, kk, dim=1, largest=False, sorted=True).values
            if kk < k:
                best = torch.cat([best, torch.full((nq, k - kk), torch.iinfo(torch.int64).max, dtype=torch.int64, device=dev)], dim=1)
            none = best == torch.iinfo(torch.int64).max
            Ds.append(torch.where(none, torch.full((1,), 3.4028234663852886e38, device=dev), (best >> 32).to(torch.int32).view(torch.float32)))
            Is.append(torch.where(none, torch.full_like(best, -1), (best & 0xFFFFFFFF) + id_offset))
        D, I = torch.stack(Ds), torch.stack(Is)
        if squeeze:
            D, I = D[0], I[0]
        if was_numpy:
            return D.cpu().numpy(), I.cpu().numpy()
        return D, I

    def search(self, q, k: int, w0: int = 0, id_offset: int = 0):
        if int(k) < 1:
            raise ValueError("k must be >= 1")
        if int(k) > MAX_K:
            return self._search_large_k(q, int(k), int(w0), int(id_offset))
        a = _Arg(q)
        if a.np_dtype != np.dtype(np.float32):
            a = _Arg(a.arr.float() if _is_torch(a.arr) else a.arr.astype(np.float32))
        squeeze = len(a.shape) == 2
        if squeeze:
            nw, nq = 1, a.shape[0]
        elif len(a.shape) == 3:
            nw, nq = a.shape[0], a.shape[1]
        else:
            raise ValueError(f"search: expected a 2-D or 3-D query array, got shape {a.shape}")
        if a.shape[-1] != self.d:
            raise ValueError(f"search: last dimension must be d={self.d}, got {a.shape[-1]}")
        flags = L.Q_ON_DEVICE | L.OUT_ON_DEVICE if a.on_device else 0
        shape = (nw, nq, int(k))
        D, Dp = self._alloc_out(a.on_device, shape, np.float32)
        I, Ip = self._alloc_out(a.on_device, shape, np.int64)
        L.check(self._lib.snv_index_search(self._h, int(w0), nw, a.ptr, nq, L.DT_F32, None, L.MASK_NONE, int(k),
                                           int(id_offset), None, Dp, Ip, flags, _current_stream(self.device)),
                "snv_index_search")
        if squeeze:
            return D[0], I[0]
        return D, I

    def gather_rows(self, I, w0: int = 0):
        """panel[I] -> float32 [.., k, d] (src/dataset/embedding_rag_dataset.py:406-438)."""
        a = _Arg(I)
        if a.np_dtype != np.dtype(np.int64):
            raise ValueError("gather_rows: I must be int64")
        squeeze = len(a.shape) == 2
        nw, nq, k = (1,) + tuple(a.shape) if squeeze else tuple(a.shape)
        flags = L.Q_ON_DEVICE | L.OUT_ON_DEVICE if a.on_device else 0
        out, op = self._alloc_out(a.on_device, (nw, nq, k, self.d), np.float32)
        L.check(self._lib.snv_index_gather_rows(self._h, int(w0), nw, a.ptr, nq, k, op, flags,
                                                _current_stream(self.device)), "snv_index_gather_rows")
        return out[0] if squeeze else out

    def export_rows(self, window: int = 0) -> np.ndarray:
        out = np.zeros((self.ntotal, self.d), dtype=np.float32)
        L.check(self._lib.snv_index_export(self._h, int(window), out.ctypes.data), "snv_index_export")
        return out


def topk_merge(D, I, k: int):
    """Merge per-shard results: D, I torch CUDA tensors [parts, nq, k_in] (global ids) ->
    (D [nq, k], I [nq, k]) by (distance, id).  Used after the all-gather of a row-sharded search."""
    if not (_is_torch(D) and D.is_cuda and _is_torch(I) and I.is_cuda):
        raise ValueError("topk_merge takes CUDA tensors")
    D = D.contiguous()
    I = I.contiguous()
    parts, nq, kin = D.shape
    dev = D.device.index
    Do = torch.empty((nq, k), dtype=D.dtype, device=D.device)
    Io = torch.empty((nq, k), dtype=torch.int64, device=D.device)
    is_i = D.dtype == torch.int32
    if not is_i and D.dtype != torch.float32:
        raise ValueError("topk_merge: D must be int32 or float32")
    L.check(L.lib().snv_topk_merge(dev, D.data_ptr() if is_i else None, None if is_i else D.data_ptr(), I.data_ptr(),
                                   parts, nq, kin, int(k), Do.data_ptr() if is_i else None,
                                   None if is_i else Do.data_ptr(), Io.data_ptr(), _current_stream(dev)),
            "snv_topk_merge")
    return Do, Io


def intersect_masks(ref_pos, tgt_pos, window_info, d: int, ploidy: int = 1):
    """Device-side position intersection (snv_intersect_masks): CUDA int64 tensors ref_pos [n_ref], tgt_pos [n_tgt]
    (any order; sorted here on the device), window_info [W, 2] (start, end) -> packed int32 observed masks
    [W, snv_packed_stride(d)], usable as `observed=` of a search over packed queries (one mask per window)."""
    for t in (ref_pos, tgt_pos, window_info):
        if not (_is_torch(t) and t.is_cuda and t.dtype == torch.int64):
            raise ValueError("intersect_masks takes CUDA int64 tensors")
    if window_info.dim() != 2 or window_info.shape[1] != 2:
        raise ValueError("intersect_masks: window_info must be [W, 2]")
    dev = ref_pos.device.index
    ref_pos = ref_pos.contiguous()
    tgt_sorted = torch.sort(tgt_pos.contiguous()).values
    window_info = window_info.contiguous()
    W = int(window_info.shape[0])
    out = torch.empty((W, L.packed_stride(int(d))), dtype=torch.int32, device=ref_pos.device)
    L.check(L.lib().snv_intersect_masks(dev, ref_pos.data_ptr(), int(ref_pos.numel()), tgt_sorted.data_ptr(), int(tgt_sorted.numel()),
                                        window_info.data_ptr(), W, int(d), int(ploidy), out.data_ptr(), _current_stream(dev)),
            "snv_intersect_masks")
    return out


def pack_rows(x, d: Optional[int] = None, invert: bool = False):
    """Device-side bit packing (snv_pack_rows): torch CUDA tensor [rows, d] (uint8/bool 0-1 sites,
    float32, or int64 tokens) -> packed int32 tensor [rows, snv_packed_stride(d)]."""
    if not (_is_torch(x) and x.is_cuda):
        raise ValueError("pack_rows takes a CUDA tensor")
    a = _Arg(x)
    if len(a.shape) != 2:
        raise ValueError("pack_rows: expected [rows, d]")
    d = a.shape[1] if d is None else int(d)
    stride = L.packed_stride(d)
    dt, a = _hamming_dtype(a, d, stride, "pack_rows")
    if dt == L.DT_PACKED_U32:
        raise ValueError("pack_rows: input is already packed")
    dev = a.arr.device.index
    out = torch.empty((a.shape[0], stride), dtype=torch.int32, device=a.arr.device)
    L.check(L.lib().snv_pack_rows(dev, a.ptr, a.shape[0], d, dt, int(bool(invert)), out.data_ptr(), None,
                                  _current_stream(dev)), "snv_pack_rows")
    return out
