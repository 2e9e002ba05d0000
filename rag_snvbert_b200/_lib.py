"""ctypes binding of libsnvknn.so (the C ABI declared in include/snvknn.h).

There is no CPU fallback: if the shared library is missing the import of any compute entry
point raises, and on a machine without a CUDA device every compute call raises RuntimeError.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("SNVKNN_LIB") or os.path.join(_PKG, "libsnvknn.so")  # SNVKNN_LIB: tuning builds only
_CSRC = os.path.join(_PKG, "csrc")

# enums of include/snvknn.h
SNV_OK = 0
KIND_HAMMING, KIND_L2 = 0, 1
DT_U8, DT_F32, DT_PACKED_U32, DT_PACKED_U8, DT_I64_TOKENS, DT_PACKED_U32_DENSE = 0, 1, 2, 3, 4, 5
MASK_NONE, MASK_PER_WINDOW, MASK_PER_QUERY = 0, 1, 2
Q_ON_DEVICE, OUT_ON_DEVICE, MASK_IS_MISSING, X_ON_DEVICE = 0x1, 0x2, 0x4, 0x8
L2_TF32, L2_TF32X3 = 0, 1
L2_CENTER = 0x10
L2_CENTER_AUTO = 0x20

_c = ctypes
_vp, _i, _i64, _u = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_uint

# name -> (restype, argtypes): every symbol include/snvknn.h declares
SYMBOLS = {
    "snv_last_error": (_c.c_char_p, []),
    "snv_version": (_i, []),
    "snv_device_count": (_i, [_c.POINTER(_i)]),
    "snv_packed_words": (_i64, [_i64]),
    "snv_packed_stride": (_i64, [_i64]),
    "snv_index_create": (_i, [_i, _i64, _i, _i, _i, _c.POINTER(_vp)]),
    "snv_index_free": (None, [_vp]),
    "snv_index_ntotal": (_i64, [_vp]),
    "snv_index_d": (_i64, [_vp]),
    "snv_index_kind": (_i, [_vp]),
    "snv_index_n_windows": (_i, [_vp]),
    "snv_index_device": (_i, [_vp]),
    "snv_index_reset": (_i, [_vp]),
    "snv_index_add": (_i, [_vp, _vp, _i64, _i, _u, _vp]),
    "snv_index_search": (_i, [_vp, _i, _i, _vp, _i64, _i, _vp, _i, _i, _i64, _vp, _vp, _vp, _u, _vp]),
    "snv_index_search_compact": (_i, [_vp, _i, _i, _vp, _i64, _i, _vp, _i, _i, _vp, _vp, _u, _vp]),
    "snv_index_search_grouped": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _i, _vp, _vp, _vp, _u, _vp]),
    "snv_index_gather_tokens_grouped": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _i, _vp, _u, _vp]),
    "snv_index_gather_tokens": (_i, [_vp, _i, _i, _vp, _i64, _i, _vp, _i, _vp, _u, _vp]),
    "snv_index_gather_rows": (_i, [_vp, _i, _i, _vp, _i64, _i, _vp, _u, _vp]),
    "snv_index_export": (_i, [_vp, _i, _vp]),
    "snv_topk_merge": (_i, [_i, _vp, _vp, _vp, _i, _i64, _i, _i, _vp, _vp, _vp, _vp]),
    "snv_exchange_pack": (_i, [_i, _vp, _vp, _i, _i64, _i, _i, _vp, _vp]),
    "snv_exchange_merge": (_i, [_i, _vp, _i, _i64, _i, _i, _vp, _vp, _vp]),
    "snv_peer_create": (_i, [_i, _i, _i, _c.c_size_t, _c.POINTER(_vp), _vp]),
    "snv_peer_open": (_i, [_vp, _vp]),
    "snv_peer_open_local": (_i, [_vp, _vp]),
    "snv_peer_exchange": (_i, [_vp, _vp, _vp, _i, _i64, _i, _i, _vp, _vp, _vp]),
    "snv_peer_push": (_i, [_vp, _vp, _vp, _i, _i64, _i, _c.POINTER(_c.c_uint64), _vp]),
    "snv_peer_merge": (_i, [_vp, _c.c_uint64, _i, _i64, _i, _i, _vp, _vp, _vp]),
    "snv_peer_destroy": (_i, [_vp]),
    "snv_pack_rows": (_i, [_i, _vp, _i64, _i64, _i, _i, _vp, _vp, _vp]),
    "snv_intersect_masks": (_i, [_i, _vp, _i64, _vp, _i64, _vp, _i, _i64, _i, _vp, _vp]),
    "snv_launch_count": (_i64, []),
    "snv_profile_enable": (_i, [_i]),
    "snv_profile_last_ms": (_i, [_c.POINTER(_c.c_float)]),
    "snv_last_hamming_engine": (_i, []),
    "snv_debug_hamming_plan": (_i, [_i, _i, _i64, _i, _i, _vp, _vp, _i64, _vp]),
    "snv_debug_hamming_chunks": (_i, [_i, _i, _i64, _i, _i, _i, _vp, _i, _vp]),
    "snv_debug_tc_codes": (_i, [_i, _u, _u, _u, _vp, _vp]),
}

_lib = None


def build(force: bool = False, jobs: int = 8) -> str:
    """Compile csrc/*.cu for sm_100a into rag_snvbert_b200/libsnvknn.so (nvcc; no GPU needed)."""
    cmd = ["make", "-C", _CSRC, f"-j{jobs}"]
    if force:
        subprocess.check_call(["make", "-C", _CSRC, "clean"])
    subprocess.check_call(cmd)
    return _SO


def so_path() -> str:
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise RuntimeError(
                f"{_SO} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). rag_snvbert_b200 has no CPU fallback."
            )
        L = ctypes.CDLL(_SO)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class SnvError(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    if rc != SNV_OK:
        msg = lib().snv_last_error().decode("utf-8", "replace")
        if rc == 1:
            raise ValueError(f"{what}: {msg}")
        raise SnvError(f"{what}: {msg} (status {rc})")


def device_count() -> int:
    n = _i(0)
    check(lib().snv_device_count(ctypes.byref(n)), "snv_device_count")
    return int(n.value)


def packed_stride(d: int) -> int:
    return int(lib().snv_packed_stride(int(d)))


def packed_words(d: int) -> int:
    return int(lib().snv_packed_words(int(d)))


def launch_count() -> int:
    return int(lib().snv_launch_count())


def last_hamming_engine() -> int:
    """0 = popcount scan, 1 = tcgen05 fp8, 2 = its bring-up variant, 3 = tcgen05 fp4, 4 = fp4 on CTA pairs, -1 = none yet."""
    return int(lib().snv_last_hamming_engine())


def debug_hamming_plan(n_windows: int, nq: int, n: int, d: int, k: int, cap: int = 1 << 16):
    """Planner of the tensor-core engine, host only (no GPU needed): (plan dict, items int64 [n_items, 8]).
    items columns: window, query tile, first panel tile, tiles, piece, pieces, partial-key row base, CTA slot."""
    import numpy as np

    plan = np.zeros(12, np.int32)
    items = np.zeros((cap, 8), np.int64)
    cnt = _c.c_int64(0)
    check(lib().snv_debug_hamming_plan(n_windows, nq, n, d, k, plan.ctypes.data, items.ctypes.data, cap, _c.addressof(cnt)),
          "snv_debug_hamming_plan")
    if cnt.value > cap:
        return debug_hamming_plan(n_windows, nq, n, d, k, cap=int(cnt.value))
    names = ["engine", "kt", "kblocks", "qtiles", "n_tiles", "nsplit", "tiles_per_split", "idx_bits", "tail_items",
             "tail_split", "tail_tiles", "workspace_kib"]
    return dict(zip(names, (int(v) for v in plan))), items[: cnt.value]


def debug_tc_codes(fp4: bool, q_word: int, mask_word: int, panel_word: int):
    """Operand codes of one packed word as the tensor-core kernels compute them (host only): (query codes, panel
    codes) as uint8 arrays of 16 (fp4) or 32 (fp8) bytes."""
    import numpy as np

    n = 16 if fp4 else 32
    a = np.zeros(n, np.uint8)
    b = np.zeros(n, np.uint8)
    check(lib().snv_debug_tc_codes(1 if fp4 else 0, q_word, mask_word, panel_word, a.ctypes.data, b.ctypes.data), "snv_debug_tc_codes")
    return a, b


def debug_hamming_chunks(n_windows: int, nq: int, n: int, d: int, k: int, host_io: bool = True):
    """Window-chunk boundaries the host-buffer pipeline of snv_index_search uses for that search (host only)."""
    import numpy as np

    cap = n_windows + 2
    out = np.zeros(cap, np.int32)
    cnt = _c.c_int(0)
    check(lib().snv_debug_hamming_chunks(n_windows, nq, n, d, k, 1 if host_io else 0, out.ctypes.data, cap, _c.addressof(cnt)),
          "snv_debug_hamming_chunks")
    return out[: cnt.value].copy()


def profile_enable(on: bool) -> None:
    check(lib().snv_profile_enable(1 if on else 0), "snv_profile_enable")


def profile_last_ms() -> float:
    ms = _c.c_float(0.0)
    check(lib().snv_profile_last_ms(ctypes.byref(ms)), "snv_profile_last_ms")
    return float(ms.value)
