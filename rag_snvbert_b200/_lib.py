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
DT_U8, DT_F32, DT_PACKED_U32, DT_PACKED_U8, DT_I64_TOKENS = 0, 1, 2, 3, 4
MASK_NONE, MASK_PER_WINDOW, MASK_PER_QUERY = 0, 1, 2
Q_ON_DEVICE, OUT_ON_DEVICE, MASK_IS_MISSING, X_ON_DEVICE = 0x1, 0x2, 0x4, 0x8
L2_TF32, L2_TF32X3 = 0, 1
L2_CENTER = 0x10

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
    "snv_index_search_grouped": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _i, _vp, _vp, _vp, _u, _vp]),
    "snv_index_gather_tokens_grouped": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _i, _vp, _u, _vp]),
    "snv_index_gather_tokens": (_i, [_vp, _i, _i, _vp, _i64, _i, _vp, _i, _vp, _u, _vp]),
    "snv_index_gather_rows": (_i, [_vp, _i, _i, _vp, _i64, _i, _vp, _u, _vp]),
    "snv_index_export": (_i, [_vp, _i, _vp]),
    "snv_topk_merge": (_i, [_i, _vp, _vp, _vp, _i, _i64, _i, _i, _vp, _vp, _vp, _vp]),
    "snv_pack_rows": (_i, [_i, _vp, _i64, _i64, _i, _i, _vp, _vp, _vp]),
    "snv_intersect_masks": (_i, [_i, _vp, _i64, _vp, _i64, _vp, _i, _i64, _i, _vp, _vp]),
    "snv_launch_count": (_i64, []),
    "snv_profile_enable": (_i, [_i]),
    "snv_profile_last_ms": (_i, [_c.POINTER(_c.c_float)]),
    "snv_last_hamming_engine": (_i, []),
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
    """0 = popcount scan, 1 = tensor cores, 2 = bring-up variant, -1 = no Hamming search yet."""
    return int(lib().snv_last_hamming_engine())


def profile_enable(on: bool) -> None:
    check(lib().snv_profile_enable(1 if on else 0), "snv_profile_enable")


def profile_last_ms() -> float:
    ms = _c.c_float(0.0)
    check(lib().snv_profile_last_ms(ctypes.byref(ms)), "snv_profile_last_ms")
    return float(ms.value)
