"""Host-side mirror of the reference's `VectorStore` (src/vectordb/store.rs) over libcsgpu.so.

Same method names, argument meaning and error text as the Rust struct so the parity tests read
like the reference's own tests (store.rs:826-1029):

    VectorStore.new(db_path, dimensions)            store.rs:110
    insert_chunks / insert_chunks_with_ids           store.rs:334,618
    delete_chunks(ids) -> count                      store.rs:548
    build_index()                                    store.rs:386
    search(query_embedding, limit) -> [SearchResult] store.rs:431   <- the CUDA path
    search_filtered / search_batch                   new, additive
    get_chunk / get_chunk_as_result / get_chunks_by_file / stats / clear / is_indexed

Chunk metadata (the LMDB "chunks" table, store.rs:97) stays on the host, here as a dict keyed by
chunk id; only the arroy block (store.rs:446-459) runs on the GPU. This file contains no scoring
arithmetic and no fallback: every search is one call into the C ABI.
"""
from __future__ import annotations

import ctypes
import json
import os
from dataclasses import asdict, dataclass, field
from typing import Iterable, Sequence

import numpy as np

from . import _lib
from ._lib import CsgpuError
from .tags import FileTable, TagPredicate


@dataclass
class Chunk:
    """chunker::Chunk (src/chunker/mod.rs:22-63), the fields the store persists."""
    content: str
    start_line: int
    end_line: int
    kind: str
    path: str
    signature: str | None = None
    docstring: str | None = None
    context: str | None = None
    hash: str = ""
    context_prev: str | None = None
    context_next: str | None = None


@dataclass
class EmbeddedChunk:
    """embed::EmbeddedChunk (src/embed/batch.rs:47-51)."""
    chunk: Chunk
    embedding: Sequence[float]


@dataclass
class SearchResult:
    """store.rs:753-772."""
    id: int
    content: str
    path: str
    start_line: int
    end_line: int
    kind: str
    signature: str | None
    docstring: str | None
    context: str | None
    hash: str
    distance: float
    score: float
    context_prev: str | None = None
    context_next: str | None = None


@dataclass
class StoreStats:
    """store.rs:784-792."""
    total_chunks: int
    total_files: int
    indexed: bool
    dimensions: int
    max_chunk_id: int


@dataclass
class RowFilter:
    """Allow-set over chunk ids for search_filtered (bit i set <=> chunk id i may be returned)."""
    bitmap: np.ndarray  # uint64 words
    n_bits: int

    @staticmethod
    def from_ids(ids: Iterable[int], n_bits: int) -> "RowFilter":
        bm = np.zeros((n_bits + 63) // 64, dtype=np.uint64)
        ids = np.asarray(list(ids) if not isinstance(ids, np.ndarray) else ids, dtype=np.uint64)
        ids = ids[ids < n_bits]
        np.bitwise_or.at(bm, (ids >> np.uint64(6)).astype(np.int64), np.uint64(1) << (ids & np.uint64(63)))
        return RowFilter(bm, n_bits)

    @staticmethod
    def from_mask(mask: np.ndarray) -> "RowFilter":
        mask = np.asarray(mask, dtype=bool)
        n_bits = mask.size
        padded = np.zeros(((n_bits + 63) // 64) * 64, dtype=np.uint8)
        padded[:n_bits] = mask
        bm = np.packbits(padded.reshape(-1, 8), axis=1, bitorder="little").reshape(-1).view(np.uint64)
        return RowFilter(np.ascontiguousarray(bm), n_bits)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


class VectorStore:
    def __init__(self, dimensions: int, devices: Sequence[int] | None = None, db_path=None, dtype: str = "fp32"):
        self._lib = _lib.load()
        self.dimensions = int(dimensions)
        self.db_path = db_path
        self.next_id = 0
        self._chunks: dict[int, Chunk] = {}
        self.files = FileTable()   # path -> file_id / language for the row tags
        self._h = ctypes.c_void_p()
        devs = None
        n = 1
        if devices is not None:
            n = len(devices)
            devs = (ctypes.c_int32 * n)(*devices)
        # CODESEARCH_GPU_INDEX_DTYPE=fp32|bf16 in the reference's env-knob style; bf16 is opt-in
        self.dtype = dtype
        code = {"fp32": _lib.DTYPE_F32, "bf16": _lib.DTYPE_BF16}[dtype]
        _lib.check(self._lib.csgpu_create(ctypes.byref(self._h), self.dimensions, code, devs, n))
        self.read_only = False
        if db_path is not None:
            # store.rs:116 create_dir_all; :141-163 next_id = last key + 1, indexed = "Reader::open succeeds".
            # Here: the device snapshot <db>/gpu (csgpu_save at build_index) plays arroy's part and
            # <db>/chunks.jsonl the LMDB "chunks" table's (store.rs:135-136).
            os.makedirs(db_path, exist_ok=True)
            if os.path.exists(os.path.join(self._gpu_dir(), "meta.json")):
                _lib.check(self._lib.csgpu_load(self._h, self._gpu_dir().encode()))
                with open(os.path.join(db_path, "chunks.jsonl")) as f:
                    for line in f:
                        rec = json.loads(line)
                        cid = int(rec.pop("id"))
                        self._chunks[cid] = Chunk(**rec)
                # the tags in <db>/gpu/tags.u32 carry the file ids assigned at INSERT time; after delete + re-insert
                # cycles those are not re-derivable from the surviving chunks, so the table is persisted verbatim
                files_json = os.path.join(db_path, "files.json")
                if os.path.exists(files_json):
                    with open(files_json) as f:
                        self.files = FileTable.from_paths(json.load(f)["paths"])
                else:   # snapshot written before the table was persisted: first-seen order over surviving chunks
                    for cid in sorted(self._chunks):
                        self.files.file_id(self._chunks[cid].path)
                if self._chunks:
                    self.next_id = max(self._chunks) + 1

    # -- lifecycle ---------------------------------------------------------------------------
    @classmethod
    def new(cls, db_path, dimensions: int, devices: Sequence[int] | None = None, dtype: str = "fp32") -> "VectorStore":
        return cls(dimensions, devices=devices, db_path=db_path, dtype=dtype)

    @classmethod
    def open_readonly(cls, db_path, dimensions: int, devices: Sequence[int] | None = None, dtype: str = "fp32") -> "VectorStore":
        """store.rs:183-250: searches while another process writes; mutations are refused."""
        st = cls(dimensions, devices=devices, db_path=db_path, dtype=dtype)
        st.read_only = True
        return st

    def _gpu_dir(self) -> str:
        return os.path.join(self.db_path, "gpu")

    def _check_writable(self) -> None:
        if self.read_only:
            raise CsgpuError(_lib.ERR_ARG, "VectorStore is open read-only")

    def close(self) -> None:
        if self._h:
            self._lib.csgpu_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def is_indexed(self) -> bool:
        return bool(self._stats().built)

    # -- write side --------------------------------------------------------------------------
    def insert_chunks_with_ids(self, chunks: Sequence[EmbeddedChunk]) -> list[int]:
        self._check_writable()
        if not chunks:
            return []
        for c in chunks:
            if len(c.embedding) != self.dimensions:
                raise ValueError(
                    f"Embedding dimension mismatch: expected {self.dimensions}, got {len(c.embedding)}")
        start = self.next_id
        ids = np.arange(start, start + len(chunks), dtype=np.uint32)
        rows = _f32([c.embedding for c in chunks])
        tags = np.array([self.files.tag(c.chunk.path) for c in chunks], dtype=np.uint32)
        self.append_rows(rows, ids, tags)
        for i, c in zip(ids, chunks):
            self._chunks[int(i)] = c.chunk
        self.next_id = start + len(chunks)
        return [int(i) for i in ids]

    def insert_chunks(self, chunks: Sequence[EmbeddedChunk]) -> int:
        return len(self.insert_chunks_with_ids(chunks))

    def append_rows(self, rows: np.ndarray, ids: np.ndarray, tags: np.ndarray | None = None) -> None:
        """Bulk path: raw [n, dim] embeddings with explicit chunk ids (no metadata); optional packed row tags."""
        rows = _f32(rows)
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        if rows.ndim != 2 or rows.shape[1] != self.dimensions:
            raise ValueError(
                f"Embedding dimension mismatch: expected {self.dimensions}, got {rows.shape[-1]}")
        assert ids.shape == (rows.shape[0],)
        if tags is None:
            _lib.check(self._lib.csgpu_append(self._h, rows.ctypes.data_as(_lib._f32p),
                                              ids.ctypes.data_as(_lib._u32p), rows.shape[0]))
        else:
            tags = np.ascontiguousarray(tags, dtype=np.uint32)
            assert tags.shape == ids.shape
            _lib.check(self._lib.csgpu_append_tagged(self._h, rows.ctypes.data_as(_lib._f32p), ids.ctypes.data_as(_lib._u32p),
                                                     tags.ctypes.data_as(_lib._u32p), rows.shape[0]))
        if ids.size:
            self.next_id = max(self.next_id, int(ids.max()) + 1)

    def append_synthetic(self, seed: int, first_row: int, n: int, id_base: int = 0, tagged: bool = False) -> None:
        fn = self._lib.csgpu_append_synthetic_tagged if tagged else self._lib.csgpu_append_synthetic
        _lib.check(fn(self._h, seed, first_row, n, id_base))

    def get_tags(self, ids) -> np.ndarray:
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        out = np.empty(ids.size, dtype=np.uint32)
        _lib.check(self._lib.csgpu_get_tags(self._h, ids.ctypes.data_as(_lib._u32p), ids.size, out.ctypes.data_as(_lib._u32p)))
        return out

    def reserve(self, total_rows: int) -> None:
        _lib.check(self._lib.csgpu_reserve(self._h, total_rows))

    def delete_chunks(self, chunk_ids: Sequence[int]) -> int:
        self._check_writable()
        if len(chunk_ids) == 0:
            return 0
        ids = np.ascontiguousarray(chunk_ids, dtype=np.uint32)
        removed = ctypes.c_uint64(0)
        _lib.check(self._lib.csgpu_remove(self._h, ids.ctypes.data_as(_lib._u32p), ids.size, ctypes.byref(removed)))
        for i in ids:
            self._chunks.pop(int(i), None)
        return int(removed.value)

    def build_index(self) -> None:
        self._check_writable()
        _lib.check(self._lib.csgpu_build(self._h))
        if self.db_path is not None:
            self.save_snapshot()

    def save_snapshot(self) -> None:
        """Sidecar snapshot: device rows/ids via csgpu_save + the chunk table."""
        tmp = os.path.join(self.db_path, "chunks.jsonl.tmp")
        with open(tmp, "w") as f:
            for i in sorted(self._chunks):
                f.write(json.dumps({"id": i, **asdict(self._chunks[i])}) + "\n")
        os.replace(tmp, os.path.join(self.db_path, "chunks.jsonl"))
        tmp = os.path.join(self.db_path, "files.json.tmp")
        with open(tmp, "w") as f:   # path list in file-id order: the row tags in HBM / tags.u32 index into it
            json.dump({"paths": self.files.paths}, f)
        os.replace(tmp, os.path.join(self.db_path, "files.json"))
        _lib.check(self._lib.csgpu_save(self._h, self._gpu_dir().encode()))

    def clear(self) -> None:
        self._check_writable()
        _lib.check(self._lib.csgpu_clear(self._h))
        self._chunks.clear()
        self.files = FileTable()
        self.next_id = 0
        if self.db_path is not None:   # store.rs:690-706 clears both LMDB tables
            for name in ("gpu/meta.json", "gpu/ids.u32", "gpu/rows.f32", "gpu/rows.bf16", "gpu/tags.u32", "gpu/zero.u32", "chunks.jsonl", "files.json"):
                try:
                    os.remove(os.path.join(self.db_path, name))
                except FileNotFoundError:
                    pass

    # -- search ------------------------------------------------------------------------------
    def search_ids(self, query_embedding, limit: int, filter: RowFilter | None = None):
        """The raw hot path: (ids[u32], distance[f32]) ascending (distance, id)."""
        q = _f32(query_embedding).reshape(-1)
        k = int(limit)
        oi = np.empty(max(k, 1), dtype=np.uint32)
        od = np.empty(max(k, 1), dtype=np.float32)
        on = ctypes.c_uint32(0)
        if filter is None:
            rc = self._lib.csgpu_search(self._h, q.ctypes.data_as(_lib._f32p), q.size, k,
                                        oi.ctypes.data_as(_lib._u32p), od.ctypes.data_as(_lib._f32p), ctypes.byref(on))
        else:
            bm = np.ascontiguousarray(filter.bitmap, dtype=np.uint64)
            rc = self._lib.csgpu_search_filtered(self._h, q.ctypes.data_as(_lib._f32p), q.size, k,
                                                 bm.ctypes.data_as(_lib._u64p), filter.n_bits,
                                                 oi.ctypes.data_as(_lib._u32p), od.ctypes.data_as(_lib._f32p),
                                                 ctypes.byref(on))
        _lib.check(rc)
        return oi[: on.value].copy(), od[: on.value].copy()

    def search_tagged_ids(self, query_embedding, limit: int, pred: TagPredicate):
        """Predicate-filtered hot path (csgpu_search_tagged): language mask / file range / per-file bitmap evaluated on
        the device from the row tags; rows that fail are never read."""
        q = _f32(query_embedding).reshape(-1)
        k = int(limit)
        oi = np.empty(max(k, 1), dtype=np.uint32)
        od = np.empty(max(k, 1), dtype=np.float32)
        on = ctypes.c_uint32(0)
        cp = _lib.Predicate(pred.lang_mask & 0xFFFFFFFF, pred.file_lo, pred.file_hi & 0xFFFFFFFF, 0, None, 0)
        bm = None
        if pred.file_bitmap is not None:
            bm = np.ascontiguousarray(pred.file_bitmap, dtype=np.uint64)
            cp.file_bitmap = bm.ctypes.data
            cp.n_file_bits = int(pred.n_file_bits)
        _lib.check(self._lib.csgpu_search_tagged(self._h, q.ctypes.data_as(_lib._f32p), q.size, k, ctypes.byref(cp),
                                                 oi.ctypes.data_as(_lib._u32p), od.ctypes.data_as(_lib._f32p), ctypes.byref(on)))
        return oi[: on.value].copy(), od[: on.value].copy()

    def search_tagged(self, query_embedding, limit: int, languages=None, path_prefix: str | None = None,
                      path_contains: str | None = None, project_root: str = "") -> list[SearchResult]:
        """search() restricted to files of the given languages and/or under a path prefix / containing a substring —
        the filters the reference applies AFTER search on the host (src/search/mod.rs:727-737, src/server/mod.rs:553-559),
        applied here BEFORE scoring so that `limit` results always survive."""
        pred = self.files.predicate(languages, path_prefix, path_contains, project_root)
        return self._join(*self.search_tagged_ids(query_embedding, limit, pred))

    def search_batch_ids(self, queries, limit: int):
        q = _f32(queries)
        b, d = q.shape
        k = int(limit)
        oi = np.zeros((b, max(k, 1)), dtype=np.uint32)
        od = np.zeros((b, max(k, 1)), dtype=np.float32)
        on = np.zeros(b, dtype=np.uint32)
        _lib.check(self._lib.csgpu_search_batch(self._h, q.ctypes.data_as(_lib._f32p), d, b, k,
                                                oi.ctypes.data_as(_lib._u32p), od.ctypes.data_as(_lib._f32p),
                                                on.ctypes.data_as(_lib._u32p)))
        return oi, od, on

    def search_variants_ids(self, queries, limit: int):
        """<= 16 query variants of one user query -> one deduplicated list (src/search/mod.rs:508-590 on the device)."""
        q = _f32(queries)
        b, d = q.shape
        k = int(limit)
        oi = np.empty(max(k, 1), dtype=np.uint32)
        od = np.empty(max(k, 1), dtype=np.float32)
        on = ctypes.c_uint32(0)
        _lib.check(self._lib.csgpu_search_variants(self._h, q.ctypes.data_as(_lib._f32p), d, b, k,
                                                   oi.ctypes.data_as(_lib._u32p), od.ctypes.data_as(_lib._f32p),
                                                   ctypes.byref(on)))
        return oi[: on.value].copy(), od[: on.value].copy()

    def search_variants(self, queries, limit: int) -> list[SearchResult]:
        return self._join(*self.search_variants_ids(queries, limit))

    def search_variants_tagged(self, queries, limit: int, languages=None, path_prefix: str | None = None,
                               path_contains: str | None = None, project_root: str = "") -> list[SearchResult]:
        """search_variants() restricted by language / path BEFORE scoring (see search_tagged): hybrid search under a filter."""
        pred = self.files.predicate(languages, path_prefix, path_contains, project_root)
        return self._join(*self.search_variants_tagged_ids(queries, limit, pred))

    def search_variants_tagged_ids(self, queries, limit: int, pred: TagPredicate):
        """search_variants_ids under a row-tag predicate (csgpu_search_variants_tagged): the reference's hybrid search with a
        language / path filter (src/search/mod.rs:508-590 + the post-filters at :727-737) as one call."""
        q = _f32(queries)
        b, d = q.shape
        k = int(limit)
        oi = np.empty(max(k, 1), dtype=np.uint32)
        od = np.empty(max(k, 1), dtype=np.float32)
        on = ctypes.c_uint32(0)
        cp = _lib.Predicate(pred.lang_mask & 0xFFFFFFFF, pred.file_lo, pred.file_hi & 0xFFFFFFFF, 0, None, 0)
        bm = None
        if pred.file_bitmap is not None:
            bm = np.ascontiguousarray(pred.file_bitmap, dtype=np.uint64)
            cp.file_bitmap = bm.ctypes.data
            cp.n_file_bits = int(pred.n_file_bits)
        _lib.check(self._lib.csgpu_search_variants_tagged(self._h, q.ctypes.data_as(_lib._f32p), d, b, k, ctypes.byref(cp),
                                                          oi.ctypes.data_as(_lib._u32p), od.ctypes.data_as(_lib._f32p),
                                                          ctypes.byref(on)))
        return oi[: on.value].copy(), od[: on.value].copy()

    def set_tensor_prefilter(self, enabled: bool) -> None:
        """Opt-in (fp32 index): batches run on the tensor cores against a bf16 shadow as a filter, survivors are rescored
        in fp32 with the single-query kernel's arithmetic — results bit-identical to search() (csrc/rescore.cuh)."""
        _lib.check(self._lib.csgpu_set_tensor_prefilter(self._h, 1 if enabled else 0))

    def set_byte_prefilter(self, enabled: bool) -> None:
        """Opt-in (fp32 index): search() streams a 1-byte-per-element shadow as a filter with a proven per-row bound and
        rescores the survivors in fp32 inside the same launch — results bit-identical to the default path
        (csrc/scan_i8.cuh), a quarter of the HBM bytes per query."""
        _lib.check(self._lib.csgpu_set_byte_prefilter(self._h, 1 if enabled else 0))

    def set_coalescing(self, enabled: bool, window_us: int = 0) -> None:
        """Host micro-batcher: concurrent search() calls with the same limit share one pass over the corpus."""
        _lib.check(self._lib.csgpu_set_coalescing(self._h, 1 if enabled else 0, int(window_us)))

    def _join(self, ids, dist) -> list[SearchResult]:
        out = []
        for i, d in zip(ids, dist):
            c = self._chunks.get(int(i))
            if c is None:  # store.rs:465 silently drops ids without metadata
                continue
            d = np.float32(d)
            out.append(SearchResult(int(i), c.content, c.path, c.start_line, c.end_line, c.kind, c.signature,
                                    c.docstring, c.context, c.hash, float(d), float(np.float32(1.0) - d),
                                    c.context_prev, c.context_next))
        return out

    def search(self, query_embedding, limit: int) -> list[SearchResult]:
        return self._join(*self.search_ids(query_embedding, limit))

    def search_filtered(self, query_embedding, limit: int, filter: RowFilter) -> list[SearchResult]:
        return self._join(*self.search_ids(query_embedding, limit, filter))

    def search_batch(self, queries, limit: int) -> list[list[SearchResult]]:
        oi, od, on = self.search_batch_ids(queries, limit)
        return [self._join(oi[j, : on[j]], od[j, : on[j]]) for j in range(len(on))]

    # -- metadata / stats --------------------------------------------------------------------
    def get_chunk(self, id: int) -> Chunk | None:
        return self._chunks.get(int(id))

    def get_chunk_as_result(self, id: int) -> SearchResult | None:
        c = self._chunks.get(int(id))
        if c is None:
            return None
        return SearchResult(int(id), c.content, c.path, c.start_line, c.end_line, c.kind, c.signature, c.docstring,
                            c.context, c.hash, 0.0, 0.0, c.context_prev, c.context_next)

    def get_chunks_by_file(self) -> dict[str, list[int]]:
        out: dict[str, list[int]] = {}
        for i in sorted(self._chunks):
            out.setdefault(self._chunks[i].path, []).append(i)
        return out

    def _stats(self) -> _lib.Stats:
        s = _lib.Stats()
        _lib.check(self._lib.csgpu_stats(self._h, ctypes.byref(s)))
        return s

    def device_stats(self) -> _lib.Stats:
        return self._stats()

    def stats(self) -> StoreStats:
        s = self._stats()
        return StoreStats(total_chunks=len(self._chunks), total_files=len({c.path for c in self._chunks.values()}),
                          indexed=bool(s.built), dimensions=self.dimensions,
                          max_chunk_id=max(self._chunks) if self._chunks else 0)

    @property
    def handle(self) -> ctypes.c_void_p:
        return self._h
