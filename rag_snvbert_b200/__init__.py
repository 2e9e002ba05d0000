"""rag_snvbert_b200 — B200-native exact per-window k-NN for haplotype retrieval.

The package holds only the hot path of wangbaonan/RAG-SNVBERT: hand-written
sm_100a kernels behind a C ABI (csrc/, include/snvknn.h) and the host-side mirror of the faiss
surface the reference calls (index.py, faiss_compat.py, refdb.py; collate.py / embedding_rag.py mirror the
V17 / V18 retrieval inside the training loop; sharding.py is the multi-GPU plumbing).  Importing the package
does not load the CUDA library; the first index construction does, and fails loudly when the
library or a CUDA device is missing (there is no CPU fallback).
"""
from .index import (  # noqa: F401
    IndexHamming,
    WindowedHammingIndex,
    WindowedL2Index,
    topk_merge,
)
from . import _lib  # noqa: F401

__all__ = ["IndexHamming", "WindowedHammingIndex", "WindowedL2Index", "topk_merge"]
__version__ = "0.1.0"
