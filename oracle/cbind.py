"""ctypes binding of oracle/libsnv_oracle.so (C restatement; test infrastructure only)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsnv_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "snv_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        # -march=native binaries must not travel between hosts: always rebuild when the
        # host differs is impractical, so use a portable baseline ISA with popcnt.
        subprocess.check_call(
            ["gcc", "-O3", "-mpopcnt", "-mavx2", "-fopenmp", "-fPIC", "-shared", "-o", _SO, src]
        )
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.snv_oracle_hamming_topk.restype = ctypes.c_int
        _lib.snv_oracle_hamming_topk.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
            ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
        ]
        _lib.snv_oracle_max_threads.restype = ctypes.c_int
    return _lib


def max_threads() -> int:
    return int(lib().snv_oracle_max_threads())


def hamming_topk_packed(panel: np.ndarray, queries: np.ndarray, k: int, mask: np.ndarray | None = None,
                        words: int | None = None, n_threads: int = 0):
    """panel uint32 [W, N, stride] (or [N, stride]); queries uint32 [W, Q, stride];
    mask None or like queries.  Returns D int32 [W, Q, k], I int64 [W, Q, k]."""
    squeeze = panel.ndim == 2
    if squeeze:
        panel, queries = panel[None], queries[None]
        mask = None if mask is None else mask[None]
    panel = np.ascontiguousarray(panel, dtype=np.uint32)
    queries = np.ascontiguousarray(queries, dtype=np.uint32)
    W, N, stride = panel.shape
    Q = queries.shape[1]
    assert queries.shape == (W, Q, stride)
    if mask is not None:
        mask = np.ascontiguousarray(mask, dtype=np.uint32)
        assert mask.shape == queries.shape
    words = stride if words is None else words
    D = np.empty((W, Q, k), dtype=np.int32)
    I = np.empty((W, Q, k), dtype=np.int64)
    rc = lib().snv_oracle_hamming_topk(
        panel.ctypes.data, queries.ctypes.data, None if mask is None else mask.ctypes.data,
        W, N, Q, words, stride, k, D.ctypes.data, I.ctypes.data, n_threads)
    if rc != 0:
        raise ValueError("snv_oracle_hamming_topk: bad arguments")
    return (D[0], I[0]) if squeeze else (D, I)
