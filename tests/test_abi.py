"""CPU: the C-ABI library loads without a GPU and exports every symbol include/snvknn.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "snvknn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(snv_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    from rag_snvbert_b200 import _lib

    if not os.path.exists(_lib.so_path()):
        _lib.build()
    L = ctypes.CDLL(_lib.so_path())
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/snvknn.h but not exported"
    assert set(names) == set(_lib.SYMBOLS), "ctypes table and header disagree"


def test_no_cpu_fallback_without_device():
    from rag_snvbert_b200 import _lib, IndexHamming

    assert _lib.lib().snv_version() == 100
    assert _lib.packed_words(1030) == 33 and _lib.packed_stride(1030) == 36
    assert _lib.packed_stride(1) == 4 and _lib.packed_stride(2060) == 68 and _lib.packed_stride(5000) == 160
    if _lib.device_count() == 0:
        with pytest.raises(RuntimeError, match="no CUDA device"):
            IndexHamming(1030)


def test_peer_exchange_argument_checks_and_no_device():
    """snv_peer_*: bad arguments are refused before any device work; without a CUDA device creation fails loudly
    (no host-memory stand-in for the exchange buffer)"""
    import ctypes

    from rag_snvbert_b200 import _lib

    L = _lib.lib()
    h = ctypes.c_void_p()
    raw = ctypes.create_string_buffer(64)
    assert L.snv_peer_create(0, 2, 2, 1024, ctypes.byref(h), raw) == 1        # rank outside the world
    assert L.snv_peer_create(0, 0, 0, 1024, ctypes.byref(h), raw) == 1        # empty world
    assert L.snv_peer_create(0, 0, 2, 0, ctypes.byref(h), raw) == 1           # no slot
    assert L.snv_peer_open(None, raw) == 1 and L.snv_peer_exchange(None, None, None, 1, 2, 8, 8, None, None, None) == 1
    assert L.snv_peer_destroy(None) == 0
    if _lib.device_count() == 0:
        assert L.snv_peer_create(0, 0, 2, 1024, ctypes.byref(h), raw) != 0
        assert not h.value


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "rag_snvbert_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("the oracle", "").replace("oracle (", ""), f"{f} mentions oracle/"
