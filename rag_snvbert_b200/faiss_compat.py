"""Drop-in for the subset of the faiss Python API that RAG-SNVBERT calls.

The one-line switch in the reference's callers is

    import faiss                      ->   import rag_snvbert_b200.faiss_compat as faiss

for build_ref_db_l2.py:5, batch_test_faiss_l2.py, test_faiss*.py, partial_faiss_intersect.py,
src/dataset/rag_train_dataset.py:7, src/dataset/rag_infer_dataset.py and
src/dataset/embedding_rag_infer_dataset.py.  Same names, argument meaning and error
behaviour as the faiss SWIG wrapper for: IndexFlatL2, IndexBinaryFlat, write_index,
read_index, StandardGpuResources, index_cpu_to_gpu, index_gpu_to_cpu, omp_set_num_threads.

Everything runs on the B200 through libsnvknn (no CPU fallback): IndexFlatL2 -> tcgen05
squared-L2 search (0/1 rows: the bit-packed Hamming engine, identical results), IndexBinaryFlat ->
bit-packed Hamming search.
"""
from __future__ import annotations

import struct

import numpy as np

from .index import IndexHamming, WindowedL2Index


def _is_binary(x) -> bool:
    """True when every entry of a float matrix is exactly 0 or 1 (numpy or torch)."""
    try:
        import torch

        if isinstance(x, torch.Tensor):
            return bool(((x == 0) | (x == 1)).all().item())
    except ImportError:  # pragma: no cover
        pass
    return bool(np.logical_or(x == 0, x == 1).all())


MAX_K = 32  # neighbours the kernels select in registers; larger k (faiss takes any) goes through the block path of index.py


class IndexFlatL2:
    """faiss.IndexFlatL2(d): exact squared-L2 search (build_ref_db_l2.py:89, batch_test_faiss_l2.py:110,
    src/dataset/rag_train_dataset.py:132-134,281).  `precision`: 'tf32x3' (default, fp32-faithful) or
    'tf32' (one pass; exact for the integer-valued token / genotype vectors the reference adds).

    Binary fast path: the offline scripts add raw 0/1 genotypes as float32 (build_ref_db_l2.py:86-90,
    batch_test_faiss_l2.py:83-88), for which squared L2 IS the Hamming distance.  While every added row is
    0/1 the index also keeps a bit-packed shadow (1/32 of the float bytes) and a search whose queries are 0/1
    as well - and large enough for the scan to dominate (`binary_min_work`) - runs on the Hamming engine: same
    float32 D (exact integers), same I, same (D, I) tie order, several times faster.  Anything else falls back to the float kernel.  `binary_fast_path=False` disables it."""

    def __init__(self, d: int, precision: str = "tf32x3", device=None, binary_fast_path: bool = True):
        self.d = int(d)
        self.is_trained = True
        self.metric_type = 1  # METRIC_L2
        self._impl = WindowedL2Index(self.d, 1, device, precision)
        self._device = device
        self._binary = bool(binary_fast_path)   # still true: every row added so far is 0/1
        self._shadow = None                     # IndexHamming over the same rows while _binary
        self.last_search_path = None            # "hamming" | "l2" (introspection for tests / benchmarks)
        self.binary_min_work = 2e10             # queries x rows x sites from which the Hamming engine is used

    @property
    def ntotal(self) -> int:
        return self._impl.ntotal

    def add(self, x) -> None:
        x = _check_matrix(x, self.d, np.float32, "add")
        if self._binary and _is_binary(x):
            if self._shadow is None:
                self._shadow = IndexHamming(self.d, self._device)
            self._shadow.add(x)
        else:
            self._binary = False
            self._shadow = None
        self._impl.add(x)

    def search(self, x, k: int):
        x = _check_matrix(x, self.d, np.float32, "search")
        assert k > 0
        # worth it only when the scan outweighs the 0/1 check and the packing of the queries
        # (BASELINE cfg 1, 1000 x 5008 x 1030, is faster on the float kernel: 0.40 vs 0.75 ms from numpy)
        big = float(x.shape[0]) * self.ntotal * self.d >= self.binary_min_work
        if self._binary and self._shadow is not None and big and int(k) <= MAX_K and _is_binary(x):
            self.last_search_path = "hamming"
            return self._shadow.search(x, int(k), dist_dtype=np.float32)
        self.last_search_path = "l2"
        return self._impl.search(x, int(k))

    def reset(self) -> None:
        self._impl.reset()
        if self._shadow is not None:
            self._shadow.reset()

    def reconstruct_n(self, n0: int = 0, ni: int = -1) -> np.ndarray:
        rows = self._impl.export_rows(0)
        return rows[n0:] if ni < 0 else rows[n0 : n0 + ni]


class IndexBinaryFlat:
    """faiss.IndexBinaryFlat(d_bits): Hamming search over np.packbits codes
    (test_faiss_intersect.py:164-181).  D is int32, I int64."""

    def __init__(self, d: int, device=None):
        if int(d) % 8 != 0:
            # faiss asserts d % 8 == 0 (the reference script never hits an odd case because
            # d_bits = 2 * intersection_len is even but may not be a multiple of 8)
            raise ValueError("IndexBinaryFlat: d must be a multiple of 8")
        self.d = int(d)
        self.code_size = self.d // 8
        self.is_trained = True
        self._impl = IndexHamming(self.d, device)

    @property
    def ntotal(self) -> int:
        return self._impl.ntotal

    def add(self, x) -> None:
        x = _check_matrix(x, self.code_size, np.uint8, "add")
        self._impl.add(x, codes=True)

    def search(self, x, k: int):
        x = _check_matrix(x, self.code_size, np.uint8, "search")
        assert k > 0
        return self._impl.search(x, int(k), codes=True)

    def reset(self) -> None:
        self._impl.reset()


def _check_matrix(x, d, dtype, what):
    """The SWIG wrapper's checks: 2-D, C-contiguous after conversion, right width and dtype."""
    try:
        import torch

        if isinstance(x, torch.Tensor):
            assert x.dim() == 2 and x.shape[1] == d, f"{what}: expected [n, {d}]"
            return x
    except ImportError:  # pragma: no cover
        pass
    x = np.ascontiguousarray(x, dtype=dtype)
    assert x.ndim == 2, f"{what}: expected a 2-D array"
    assert x.shape[1] == d, f"{what}: expected {d} columns, got {x.shape[1]}"
    return x


# ---- GPU plumbing names used by src/dataset/embedding_rag_infer_dataset.py:43,217-218 ------------
class StandardGpuResources:
    def __init__(self):
        pass

    def setTempMemory(self, *_):
        pass

    def noTempMemory(self):
        pass


def index_cpu_to_gpu(res, device, index, options=None):
    """Indexes already live on the GPU; kept so the caller's line stays valid."""
    return index


def index_gpu_to_cpu(index):
    return index


def omp_set_num_threads(n):
    return None


def get_num_gpus() -> int:
    from . import _lib

    return _lib.device_count()


# ---- write_index / read_index (build_ref_db_l2.py:93, batch_test_faiss_l2.py:94,
#      src/dataset/embedding_rag_infer_dataset.py:180,217) ------------------------------------------
# File layout follows faiss's native serialisation for the two flat index types as published
# upstream (impl/index_write.cpp; faiss is un-vendored and un-pinned in the reference, so this
# is best-effort compatibility):
#   IxF2: "IxF2", int32 d, int64 ntotal, int64 dummy, int64 dummy, uint8 is_trained,
#         int32 metric_type, uint64 n_float32, float32 data[ntotal * d]
#   IBxF: "IBxF", int32 d, int32 code_size, int64 ntotal, uint8 is_trained, int32 metric_type,
#         uint64 n_bytes, uint8 codes[ntotal * code_size]
def write_index(index, path: str) -> None:
    with open(path, "wb") as f:
        if isinstance(index, IndexFlatL2):
            rows = index._impl.export_rows(0)
            f.write(b"IxF2")
            f.write(struct.pack("<iqqqBi", index.d, index.ntotal, 1 << 20, 1 << 20, 1, 1))
            f.write(struct.pack("<Q", rows.size))
            f.write(rows.astype("<f4").tobytes())
        elif isinstance(index, IndexBinaryFlat):
            packed = index._impl.export_packed(0)  # uint32 words == the original bytes, LE
            codes = packed.view(np.uint8).reshape(packed.shape[0], -1)[:, : index.code_size]
            f.write(b"IBxF")
            f.write(struct.pack("<iiqBi", index.d, index.code_size, index.ntotal, 1, 1))
            f.write(struct.pack("<Q", codes.size))
            f.write(np.ascontiguousarray(codes).tobytes())
        else:
            raise TypeError(f"write_index: unsupported index type {type(index).__name__}")


def read_index(path: str, io_flags: int = 0):
    with open(path, "rb") as f:
        four = f.read(4)
        if four == b"IxF2":
            d, ntotal, _, _, _, _ = struct.unpack("<iqqqBi", f.read(4 + 8 * 3 + 1 + 4))
            (n,) = struct.unpack("<Q", f.read(8))
            data = np.frombuffer(f.read(n * 4), dtype="<f4").reshape(ntotal, d)
            idx = IndexFlatL2(d)
            if ntotal:
                idx.add(np.ascontiguousarray(data))
            return idx
        if four == b"IBxF":
            d, code_size, ntotal, _, _ = struct.unpack("<iiqBi", f.read(4 + 4 + 8 + 1 + 4))
            (n,) = struct.unpack("<Q", f.read(8))
            data = np.frombuffer(f.read(n), dtype=np.uint8).reshape(ntotal, code_size)
            idx = IndexBinaryFlat(d)
            if ntotal:
                idx.add(np.ascontiguousarray(data))
            return idx
        raise RuntimeError(f"read_index: unsupported index fourcc {four!r}")
