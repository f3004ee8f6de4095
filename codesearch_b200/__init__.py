"""codesearch_b200 — B200-native vector-retrieval hot path for flupkede/codesearch.

Only what the path needs: csrc/ (CUDA kernels + the C ABI of include/csgpu.h, built into
libcsgpu.so) and store.py (host-side mirror of the reference's VectorStore API).
"""
from ._lib import CsgpuError, load as load_library  # noqa: F401
from .store import (Chunk, EmbeddedChunk, RowFilter, SearchResult, StoreStats,  # noqa: F401
                    VectorStore)

__all__ = ["VectorStore", "SearchResult", "StoreStats", "Chunk", "EmbeddedChunk", "RowFilter",
           "CsgpuError", "load_library"]
