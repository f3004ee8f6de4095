"""CPU oracle for the vector-retrieval hot path — TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module. The product package (codesearch_b200) never does.

PARITY STATUS: "parity unpinned" for the (1-cos)/2 distance scale and for arroy's
ANN recall — the reference (Rust + arroy 0.5.0 + LMDB) cannot be built or run here
(SURVEY.md §8c). What IS pinned: every known-answer the reference's own tests hold
for this path (tests/test_oracle.py).

Two independent restatements live here so they can check each other:
  * liboracle.so (oracle.c, gcc) via ctypes — fast, used at all sizes;
  * numpy/pure-Python functions below — small cases and golden-vector generation.

Reference anchors (all under /root/reference):
  examples/benchmark_models.rs:323-328  cosine_similarity = dot/(|a||b|), sequential f32 sums
  examples/benchmark_models.rs:155-165  linear scan, strict '>' keeps first index on ties
  src/embed/batch.rs:316-324            zero-magnitude => 0.0 (test helper)
  src/vectordb/store.rs:446-459         arroy nns().by_vector(): ascending (distance, id)
  src/vectordb/store.rs:477-478         score = 1 - distance
  Cargo.lock:162-165                    arroy 0.5.0: distance = pn*qn != 0 ? (1 - cos)/2 : 0
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle.c -> liboracle.so (building the checker is not using it)."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        u32p = ctypes.POINTER(ctypes.c_uint32)
        u64p = ctypes.POINTER(ctypes.c_uint64)
        f32p = ctypes.POINTER(ctypes.c_float)
        f64p = ctypes.POINTER(ctypes.c_double)
        L.cs_philox4x32_10.argtypes = [u32p, u32p, u32p]
        L.cs_synth_rows.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32, f32p]
        for name in ("cs_ref_cosine_similarity", "cs_ref_cosine_similarity_guarded", "cs_ref_distance_f32"):
            getattr(L, name).argtypes = [f32p, f32p, ctypes.c_uint32]
            getattr(L, name).restype = ctypes.c_float
        L.cs_ref_distance_f64.argtypes = [f32p, f32p, ctypes.c_uint32]
        L.cs_ref_distance_f64.restype = ctypes.c_double
        L.cs_oracle_search.argtypes = [f32p, u32p, ctypes.c_uint64, ctypes.c_uint32, f32p, ctypes.c_uint32,
                                       ctypes.c_int, u64p, ctypes.c_uint64, u32p, f32p, f64p]
        L.cs_oracle_search.restype = ctypes.c_uint32
        L.cs_oracle_search_synth.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32,
                                             f32p, ctypes.c_uint32, ctypes.c_uint32, u32p, f32p, f64p, u32p]
        L.cs_cpu_baseline_search.argtypes = [f32p, ctypes.c_uint64, ctypes.c_uint32, f32p, ctypes.c_uint32, u32p, f32p]
        L.cs_cpu_baseline_search.restype = ctypes.c_uint32
        L.cs_oracle_threads.restype = ctypes.c_int
        L.cs_oracle_set_threads.argtypes = [ctypes.c_int]
        _lib = L
    return _lib


def _p(a: np.ndarray, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct)) if a is not None else None


# --------------------------------------------------------------------------- #
# C-backed entry points
# --------------------------------------------------------------------------- #
def synth_rows(seed: int, first_row: int, n: int, dim: int) -> np.ndarray:
    """Raw (un-normalised) synthetic rows; bit-identical to the CUDA generator."""
    assert dim % 4 == 0
    out = np.empty((n, dim), dtype=np.float32)
    lib().cs_synth_rows(seed, first_row, n, dim, _p(out, ctypes.c_float))
    return out


def search(rows: np.ndarray, q: np.ndarray, k: int, ids: np.ndarray | None = None, mode: int = 1,
           bitmap: np.ndarray | None = None, n_bits: int = 0):
    """Top-k ascending (distance, id). Returns (ids[u32], dist[f32], dist64[f64])."""
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float32)
    n, d = rows.shape
    assert q.shape == (d,)
    if ids is not None:
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
    if bitmap is not None:
        bitmap = np.ascontiguousarray(bitmap, dtype=np.uint64)
    oi = np.zeros(max(k, 1), dtype=np.uint32)
    od = np.zeros(max(k, 1), dtype=np.float32)
    o64 = np.zeros(max(k, 1), dtype=np.float64)
    m = lib().cs_oracle_search(_p(rows, ctypes.c_float), _p(ids, ctypes.c_uint32) if ids is not None else None,
                               n, d, _p(q, ctypes.c_float), k, mode,
                               _p(bitmap, ctypes.c_uint64) if bitmap is not None else None, n_bits,
                               _p(oi, ctypes.c_uint32), _p(od, ctypes.c_float), _p(o64, ctypes.c_double))
    return oi[:m].copy(), od[:m].copy(), o64[:m].copy()


def search_synth(seed: int, first_row: int, n: int, dim: int, queries: np.ndarray, k: int):
    """Streaming f64 oracle over the never-materialised synthetic corpus; ids = global row."""
    queries = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, dim)
    b = queries.shape[0]
    oi = np.zeros((b, k), dtype=np.uint32)
    od = np.zeros((b, k), dtype=np.float32)
    o64 = np.zeros((b, k), dtype=np.float64)
    on = np.zeros(b, dtype=np.uint32)
    lib().cs_oracle_search_synth(seed, first_row, n, dim, _p(queries, ctypes.c_float), b, k,
                                 _p(oi, ctypes.c_uint32), _p(od, ctypes.c_float), _p(o64, ctypes.c_double),
                                 _p(on, ctypes.c_uint32))
    return oi, od, o64, on


def cpu_baseline_search(rows: np.ndarray, q: np.ndarray, k: int):
    """The timed CPU baseline: f32 exact scan + top-k over all host threads."""
    n, d = rows.shape
    oi = np.zeros(k, dtype=np.uint32)
    od = np.zeros(k, dtype=np.float32)
    m = lib().cs_cpu_baseline_search(_p(rows, ctypes.c_float), n, d, _p(q, ctypes.c_float), k,
                                     _p(oi, ctypes.c_uint32), _p(od, ctypes.c_float))
    return oi[:m], od[:m]


def threads() -> int:
    return lib().cs_oracle_threads()


def set_threads(n: int) -> None:
    lib().cs_oracle_set_threads(n)


def c_cosine(a, b, guarded=False) -> float:
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    f = lib().cs_ref_cosine_similarity_guarded if guarded else lib().cs_ref_cosine_similarity
    return float(f(_p(a, ctypes.c_float), _p(b, ctypes.c_float), a.size))


def c_distance_f32(p, q) -> float:
    p = np.ascontiguousarray(p, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float32)
    return float(lib().cs_ref_distance_f32(_p(p, ctypes.c_float), _p(q, ctypes.c_float), p.size))


def c_philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().cs_philox4x32_10(_p(c, ctypes.c_uint32), _p(k, ctypes.c_uint32), _p(o, ctypes.c_uint32))
    return o


# --------------------------------------------------------------------------- #
# Independent numpy / pure-Python restatement (small cases, golden generation)
# --------------------------------------------------------------------------- #
_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def np_philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Salmon et al., SC'11). All args uint32 arrays/scalars."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) for x in (c0, c1, c2, c3))
    k0 = np.uint64(k0)
    k1 = np.uint64(k1)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(_M0) * c0
        p1 = np.uint64(_M1) * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ k0
        n1 = p1 & mask
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ k1
        n3 = p0 & mask
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + np.uint64(_W0)) & mask
        k1 = (k1 + np.uint64(_W1)) & mask
    return tuple(x.astype(np.uint32) for x in (c0, c1, c2, c3))


def np_synth_rows(seed: int, first_row: int, n: int, dim: int) -> np.ndarray:
    rows = (np.arange(n, dtype=np.uint64) + np.uint64(first_row))[:, None]
    c4 = np.arange(dim // 4, dtype=np.uint64)[None, :]
    r_lo = np.broadcast_to(rows & np.uint64(0xFFFFFFFF), (n, dim // 4))
    r_hi = np.broadcast_to(rows >> np.uint64(32), (n, dim // 4))
    cc = np.broadcast_to(c4, (n, dim // 4))
    xs = np_philox4x32_10(r_lo, r_hi, cc, np.zeros_like(cc), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    out = np.empty((n, dim // 4, 4), dtype=np.float32)
    for j, x in enumerate(xs):
        s = (x & 0xFF).astype(np.int64) + ((x >> 8) & 0xFF) + ((x >> 16) & 0xFF) + (x >> 24)
        out[:, :, j] = (s - 510).astype(np.float32)
    return out.reshape(n, dim)


def py_cosine_similarity(a, b) -> np.float32:
    """benchmark_models.rs:323-328, sequential f32."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    dot = np.float32(0)
    ma = np.float32(0)
    mb = np.float32(0)
    for x, y in zip(a, b):
        dot = np.float32(dot + np.float32(x * y))
    for x in a:
        ma = np.float32(ma + np.float32(x * x))
    for y in b:
        mb = np.float32(mb + np.float32(y * y))
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.float32(dot / np.float32(np.sqrt(ma) * np.sqrt(mb)))


def py_distance_f32(p, q) -> np.float32:
    """arroy 0.5.0 Cosine::built_distance, sequential f32."""
    p = np.asarray(p, dtype=np.float32)
    q = np.asarray(q, dtype=np.float32)
    pq = np.float32(0)
    pp = np.float32(0)
    qq = np.float32(0)
    for x, y in zip(p, q):
        pq = np.float32(pq + np.float32(x * y))
    for x in p:
        pp = np.float32(pp + np.float32(x * x))
    for y in q:
        qq = np.float32(qq + np.float32(y * y))
    pnqn = np.float32(np.sqrt(pp) * np.sqrt(qq))
    if pnqn != 0:
        return np.float32((np.float32(1) - np.float32(pq / pnqn)) / np.float32(2))
    return np.float32(0)


def np_search(rows, q, k, ids=None, allowed=None):
    """f64 referee in numpy: ascending (f32-rounded distance, id), min(k, passing) results.

    allowed: optional boolean array indexed by chunk id.
    """
    rows = np.asarray(rows, dtype=np.float32)
    q = np.asarray(q, dtype=np.float32)
    n = rows.shape[0]
    ids = np.arange(n, dtype=np.uint32) if ids is None else np.asarray(ids, dtype=np.uint32)
    r64 = rows.astype(np.float64)
    q64 = q.astype(np.float64)
    pq = r64 @ q64
    pn = np.sqrt(np.einsum("ij,ij->i", r64, r64))
    qn = np.sqrt(q64 @ q64)
    pnqn = pn * qn
    with np.errstate(invalid="ignore", divide="ignore"):
        d64 = np.where(pnqn != 0, (1.0 - pq / np.where(pnqn != 0, pnqn, 1.0)) / 2.0, 0.0)
    d32 = d64.astype(np.float32)
    keep = np.ones(n, dtype=bool)
    if allowed is not None:
        allowed = np.asarray(allowed, dtype=bool)
        keep = (ids < allowed.size) & allowed[np.minimum(ids, allowed.size - 1)]
    idx = np.nonzero(keep)[0]
    order = idx[np.lexsort((ids[idx], d32[idx]))][:k]
    return ids[order], d32[order], d64[order]


def bf16_round(x: np.ndarray) -> np.ndarray:
    """float32 -> nearest bfloat16 (round to nearest even), returned as float32."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + ((u >> np.uint64(16)) & np.uint64(1)) + np.uint64(0x7FFF)) & np.uint64(0xFFFF0000)
    return u.astype(np.uint32).view(np.float32).reshape(np.shape(x))


def np_search_bf16(rows, q, k, ids=None):
    """Restatement of the opt-in bf16 index (no reference counterpart): rows and
    query are unit-normalised (f64 norm) and rounded f32 -> bf16; distance = (1 - dot)/2 with the dot
    product of the ROUNDED vectors taken exactly (f64), NOT re-normalised. Ascending (f32 distance, id).
    Returns (ids, dist32, dist64)."""
    rows = np.asarray(rows, dtype=np.float32)
    q = np.asarray(q, dtype=np.float32)
    n = rows.shape[0]
    ids = np.arange(n, dtype=np.uint32) if ids is None else np.asarray(ids, dtype=np.uint32)
    r64 = rows.astype(np.float64)
    rn = np.sqrt(np.einsum("ij,ij->i", r64, r64))
    keep = rn > 0
    ru = bf16_round((r64 / np.where(keep, rn, 1.0)[:, None]).astype(np.float32)).astype(np.float64)
    q64 = q.astype(np.float64)
    qu = bf16_round((q64 / np.sqrt(q64 @ q64)).astype(np.float32)).astype(np.float64)
    d64 = np.where(keep, (1.0 - ru @ qu) / 2.0, 0.0)
    d32 = d64.astype(np.float32)
    order = np.lexsort((ids, d32))[:k]
    return ids[order], d32[order], d64[order]


def dedup_variants(lists, k):
    """src/search/mod.rs:513-590 restated: the per-variant result lists of ONE user query are merged keeping, per
    chunk id, the entry with the best score (= smallest distance), then the top `k` of the union are returned best
    first. The reference pops a BinaryHeap fed from a HashMap, so ties come out in arbitrary order; here ties break
    by ascending id (the store's own order). lists: iterable of (ids, dist32). Returns (ids u32, dist f32)."""
    best = {}
    for ids, dist in lists:
        for i, d in zip(np.asarray(ids).tolist(), np.asarray(dist, dtype=np.float32).tolist()):
            if i not in best or d < best[i]:
                best[i] = d
    order = sorted(best.items(), key=lambda t: (np.float32(t[1]), t[0]))[:k]
    return (np.array([i for i, _ in order], dtype=np.uint32), np.array([d for _, d in order], dtype=np.float32))


def score_from_distance(distance):
    """store.rs:478."""
    return np.float32(1.0) - np.asarray(distance, dtype=np.float32)


def rrf_scores(ranked_ids, k_rrf: float = 60.0):
    """rerank/mod.rs:57-59 — consumer contract: position in the list is the rank."""
    return {int(i): 1.0 / (k_rrf + r + 1.0) for r, i in enumerate(ranked_ids)}


# ---- row tags and the file/language predicate -------------------------------------------------
# Independent restatement (the product's host half is codesearch_b200/tags.py, its device half csrc/scan.cuh).
# lang_id = position in `enum Language` (/root/reference/src/file/language.rs:5-29).
LANGUAGE_ORDER = ["Rust", "Python", "JavaScript", "TypeScript", "Go", "Java", "C", "Cpp", "CSharp", "Ruby", "Php",
                  "Swift", "Kotlin", "Shell", "Markdown", "Json", "Yaml", "Toml", "Sql", "Html", "Css", "Xml", "Unknown"]


def language_from_path(path: str) -> str:
    """`Language::from_path` (src/file/language.rs:31-44): `from_extension` (:60-88, lower-cased) first, then
    `from_filename` (:47-57) on the exact file name. Path::extension is the text after the last dot of the file
    name, where a leading dot does not start an extension."""
    name = path.replace("\\", "/").rstrip("/").split("/")[-1]
    body = name[1:] if name.startswith(".") else name
    ext = body.rsplit(".", 1)[1].lower() if "." in body else ""
    table = [("Rust", ["rs"]), ("Python", ["py", "pyw", "pyi"]), ("JavaScript", ["js", "mjs", "cjs"]),
             ("TypeScript", ["ts", "mts", "cts", "tsx", "jsx"]), ("Go", ["go"]), ("Java", ["java"]), ("C", ["c", "h"]),
             ("Cpp", ["cpp", "cc", "cxx", "hpp", "hxx"]), ("CSharp", ["cs"]), ("Ruby", ["rb", "rake"]), ("Php", ["php"]),
             ("Swift", ["swift"]), ("Kotlin", ["kt", "kts"]), ("Shell", ["sh", "bash", "zsh"]),
             ("Markdown", ["md", "markdown", "txt"]), ("Json", ["json"]), ("Yaml", ["yaml", "yml"]), ("Toml", ["toml"]),
             ("Sql", ["sql"]), ("Html", ["html", "htm"]), ("Css", ["css", "scss", "sass", "less"]),
             ("Xml", ["xml", "csproj", "props", "targets", "resx", "config"])]
    for lang, exts in table:
        if ext in exts:
            return lang
    if name in ("Dockerfile", "Containerfile", "Makefile", "GNUmakefile", "makefile", ".env", ".envrc", "CMakeLists"):
        return "Shell"
    if name in ("Jenkinsfile", "Vagrantfile", "Fastfile", "Appfile", "Podfile"):
        return "Ruby"
    return "Unknown"


def synth_tags(first_row: int, n: int) -> np.ndarray:
    """file_id = row // 37, lang_id = fmix32(file_id) % 23 (MurmurHash3 32-bit finaliser);
    tag = lang_id << 27 | file_id. Pure-Python integers (slow, small n only) so it shares nothing with the product."""
    out = np.empty(n, dtype=np.uint32)
    for i in range(n):
        f = ((first_row + i) // 37) & 0x07FFFFFF
        h = f
        h ^= h >> 16
        h = (h * 0x85EBCA6B) & 0xFFFFFFFF
        h ^= h >> 13
        h = (h * 0xC2B2AE35) & 0xFFFFFFFF
        h ^= h >> 16
        out[i] = ((h % 23) << 27) | f
    return out


def tag_predicate_mask(tags, lang_mask=0xFFFFFFFF, file_lo=0, file_hi=0xFFFFFFFF, file_bitmap=None, n_file_bits=0):
    """Row passes iff language bit set AND file_lo <= file <= file_hi AND (no bitmap OR file bit set)
    (include/csgpu.h csgpu_predicate_t). The reference's equivalents are host post-filters: path prefix
    src/search/mod.rs:727-737, path substring src/server/mod.rs:553-559."""
    out = np.zeros(len(tags), dtype=bool)
    for i, t in enumerate(np.asarray(tags, dtype=np.uint64).tolist()):
        lang, f = t >> 27, t & 0x07FFFFFF
        ok = ((lang_mask >> lang) & 1) == 1 and file_lo <= f <= file_hi
        if ok and file_bitmap is not None:
            ok = f < n_file_bits and ((int(file_bitmap[f >> 6]) >> (f & 63)) & 1) == 1
        out[i] = ok
    return out
