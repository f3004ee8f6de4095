"""Parity of the CUDA path (through the C ABI) against the oracle. GPU box only.

Bar (BASELINE.json north_star): identical top-k chunk ids (ties by id), |distance - oracle| <= 1e-5.
Tests restate the reference's own tests where they exist (src/vectordb/store.rs:826-1029).
"""
import ctypes

import numpy as np
import pytest

from parity import check_topk

pytestmark = pytest.mark.gpu

MARGIN = 8


@pytest.fixture(scope="module")
def cs():
    import codesearch_b200 as m
    m.load_library()
    return m


def make_store(cs, rows, ids=None):
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    st = cs.VectorStore.new(None, rows.shape[1])
    if ids is None:
        ids = np.arange(rows.shape[0], dtype=np.uint32)
    st.append_rows(rows, np.asarray(ids, dtype=np.uint32))
    st.build_index()
    return st


def assert_parity(oracle, st, rows, q, k, ids=None, allowed=None, flt=None):
    gi, gd = st.search_ids(q, k, flt)
    n_pass = rows.shape[0] if allowed is None else int(np.count_nonzero(allowed[ids if ids is not None else np.arange(rows.shape[0])]))
    k_eff = min(k, n_pass)
    oi, od, o64 = oracle.np_search(rows, q, k + MARGIN, ids=ids, allowed=allowed)
    return check_topk(gi, gd, oi, od, o64, k_eff)


# ---- the reference's own tests, restated (store.rs:826-1029) ---------------------------------
def test_vector_store_creation(cs):
    st = cs.VectorStore.new(None, 384)                      # store.rs:833-844
    assert st.stats().dimensions == 384
    assert not st.is_indexed()


def test_insert_and_search(cs, oracle):
    st = cs.VectorStore.new(None, 4)                        # store.rs:846-893
    chunks = [
        cs.EmbeddedChunk(cs.Chunk("fn authenticate() {}", 0, 1, "Function", "auth.rs"), [1.0, 0.0, 0.0, 0.0]),
        cs.EmbeddedChunk(cs.Chunk("fn calculate() {}", 2, 3, "Function", "math.rs"), [0.0, 1.0, 0.0, 0.0]),
    ]
    assert st.insert_chunks(chunks) == 2
    st.build_index()
    assert st.is_indexed()
    results = st.search([0.9, 0.1, 0.0, 0.0], 2)
    assert len(results) == 2
    assert "authenticate" in results[0].content
    assert results[0].score > results[1].score
    # numeric values implied by the reference arithmetic, tolerance 1e-5
    assert abs(results[0].distance - 0.0030581355) <= 1e-5 and abs(results[1].distance - 0.44478422) <= 1e-5
    assert abs(results[0].score - 0.99694186) <= 1e-5 and abs(results[1].score - 0.5552158) <= 1e-5
    assert [r.id for r in results] == [0, 1]
    assert len(st.search([0.9, 0.1, 0.0, 0.0], 10)) == 2    # min(limit, N)


def test_stats_clear_get_chunk(cs):
    st = cs.VectorStore.new(None, 4)                        # store.rs:895-1028
    ids = st.insert_chunks_with_ids([
        cs.EmbeddedChunk(cs.Chunk("a", 0, 1, "Function", "a.rs"), [1, 0, 0, 0]),
        cs.EmbeddedChunk(cs.Chunk("b", 0, 1, "Function", "a.rs"), [0, 1, 0, 0]),
        cs.EmbeddedChunk(cs.Chunk("c", 0, 1, "Function", "b.rs"), [0, 0, 1, 0])])
    assert ids == [0, 1, 2]
    assert not st.is_indexed()
    st.build_index()
    s = st.stats()
    assert (s.total_chunks, s.total_files, s.indexed, s.max_chunk_id) == (3, 2, True, 2)
    assert st.get_chunk(1).content == "b" and st.get_chunk(99) is None
    assert st.get_chunks_by_file() == {"a.rs": [0, 1], "b.rs": [2]}
    st.clear()
    assert not st.is_indexed() and st.stats().total_chunks == 0
    with pytest.raises(cs.CsgpuError) as e:
        st.search([1, 0, 0, 0], 1)
    assert e.value.code == 2


def test_guard_clauses_verbatim(cs):
    st = cs.VectorStore.new(None, 4)                        # store.rs:432-444
    st.insert_chunks([cs.EmbeddedChunk(cs.Chunk("a", 0, 1, "Function", "a.rs"), [1, 0, 0, 0])])
    with pytest.raises(cs.CsgpuError) as e:
        st.search([1, 0, 0, 0], 1)
    assert e.value.code == 2 and str(e.value) == "Index not built. Call build_index() after inserting chunks."
    st.build_index()
    with pytest.raises(cs.CsgpuError) as e:
        st.search([1, 0, 0], 1)
    assert e.value.code == 1 and str(e.value) == "Query embedding dimension mismatch: expected 4, got 3"
    with pytest.raises(ValueError) as e2:
        st.insert_chunks([cs.EmbeddedChunk(cs.Chunk("a", 0, 1, "Function", "a.rs"), [1, 0, 0])])
    assert str(e2.value) == "Embedding dimension mismatch: expected 4, got 3"
    with pytest.raises(cs.CsgpuError) as e:
        st.search([float("nan"), 0, 0, 0], 1)
    assert e.value.code == 6
    # insert/delete dirty the index again (store.rs:682, 605-607)
    st.insert_chunks([cs.EmbeddedChunk(cs.Chunk("b", 0, 1, "Function", "a.rs"), [0, 1, 0, 0])])
    assert not st.is_indexed()
    st.build_index()
    assert st.delete_chunks([0]) == 1
    assert not st.is_indexed()
    st.build_index()
    assert [r.id for r in st.search([1, 0, 0, 0], 5)] == [1]


# ---- seeded parity sweeps ---------------------------------------------------------------------
@pytest.mark.parametrize("n,d,k", [
    (1, 4, 1), (2, 4, 10), (33, 8, 32), (1000, 384, 10), (4097, 384, 10), (20000, 384, 1),
    (20000, 384, 32), (20000, 384, 33), (20000, 384, 100), (20000, 384, 200), (5000, 384, 1024),
    (3000, 768, 200), (3000, 1024, 25), (3000, 128, 10), (3000, 256, 64), (3000, 512, 10),
    (2000, 100, 10), (2000, 6, 5), (2000, 644, 50), (2000, 1000, 7), (300, 384, 1000),
])
def test_parity_gaussian(cs, oracle, n, d, k):
    rng = np.random.default_rng(n * 7919 + d * 31 + k)
    rows = rng.standard_normal((n, d)).astype(np.float32)
    st = make_store(cs, rows)
    for j in range(3):
        q = rng.standard_normal(d).astype(np.float32) * np.float32(10.0 ** (j - 1))   # query scale must not matter
        assert_parity(oracle, st, rows, q, k)


def test_parity_synthetic_100k_config0(cs, oracle):
    """BASELINE.json configs[0]: 100k x 384, single query, top-10 (the reference's CPU-runnable case)."""
    n, d = 100_000, 384
    st = cs.VectorStore.new(None, d)
    st.append_synthetic(1234, 0, n)
    st.build_index()
    rows = oracle.synth_rows(1234, 0, n, d)
    qs = oracle.synth_rows(4321, 0, 8, d)
    swaps = 0
    for q in qs:
        swaps += assert_parity(oracle, st, rows, q, 10)
        swaps += assert_parity(oracle, st, rows, q, 100)
    assert swaps == 0   # these seeds have no near-ties: ids are bit-exact
    # the committed golden fixture (tests/golden/c1_100k_top10.json): ids identical, distances within 1e-5
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c1_100k_top10.json")))
    for q, want in zip(qs, g["queries"]):
        gi, gd = st.search_ids(q, 10)
        assert gi.tolist() == want["ids"]
        assert np.abs(gd.astype(np.float64) - np.array(want["distance_f64"])).max() <= 1e-5
    # device generator == CPU generator, bit for bit
    host = np.empty((257, d), dtype=np.float32)
    from codesearch_b200 import _lib
    _lib.check(_lib.load().csgpu_synth_rows_host(st.handle, 1234, 99_000, 257, host.ctypes.data_as(_lib._f32p)))
    assert np.array_equal(host, rows[99_000:99_257])


def test_arbitrary_ids_and_gaps(cs, oracle):
    rng = np.random.default_rng(5)
    n, d = 5000, 384
    rows = rng.standard_normal((n, d)).astype(np.float32)
    ids = rng.permutation(1 << 20)[:n].astype(np.uint32)
    ids[0] = 0xFFFFFFFE
    st = make_store(cs, rows, ids)
    for k in (10, 100):
        assert_parity(oracle, st, rows, rng.standard_normal(d).astype(np.float32), k, ids=ids)


def test_duplicates_tie_break_by_id(cs, oracle):
    rng = np.random.default_rng(6)
    base = rng.standard_normal((50, 384)).astype(np.float32)
    rows = np.concatenate([base] * 8)                          # every vector 8 times, far apart in the scan
    ids = rng.permutation(rows.shape[0]).astype(np.uint32)
    st = make_store(cs, rows, ids)
    for k in (10, 32, 64, 400):
        q = base[3] + 0.01 * rng.standard_normal(384).astype(np.float32)
        gi, gd = st.search_ids(q, k)
        oi, od, o64 = oracle.np_search(rows, q, k, ids=ids)
        assert np.array_equal(gi, oi)                           # exact, including inside tie groups
        for g in range(0, (k // 8) * 8, 8):
            assert len(set(gd[g:g + 8].tolist())) == 1          # duplicates tie bit-exactly on the GPU
            assert gi[g:g + 8].tolist() == sorted(gi[g:g + 8].tolist())


def test_all_rows_tie(cs):
    rows = np.tile(np.arange(1, 385, dtype=np.float32), (3000, 1))
    ids = np.random.default_rng(7).permutation(3000).astype(np.uint32)
    st = make_store(cs, rows, ids)
    for k in (10, 100):
        gi, gd = st.search_ids(rows[0], k)
        assert gi.tolist() == list(range(k))                    # all distances equal -> smallest ids
        assert np.abs(gd).max() <= 1e-6


def test_zero_norm_rows_and_query(cs, oracle):
    rng = np.random.default_rng(8)
    rows = rng.standard_normal((2000, 384)).astype(np.float32)
    rows[[5, 700, 1999]] = 0.0                                  # arroy: pn*qn == 0 -> distance 0.0
    st = make_store(cs, rows)
    assert st.device_stats().zero_norm_rows == 3 and st.device_stats().live_rows == 2000
    q = rng.standard_normal(384).astype(np.float32)
    for k in (2, 10, 100):
        gi, gd = st.search_ids(q, k)
        oi, od, o64 = oracle.np_search(rows, q, k)
        assert np.array_equal(gi, oi) and np.abs(gd - od).max() <= 1e-5
        assert gi[:min(k, 3)].tolist() == [5, 700, 1999][:min(k, 3)] and (gd[:min(k, 3)] == 0.0).all()
    # zero query: every distance is 0.0, order = ascending id
    gi, gd = st.search_ids(np.zeros(384, np.float32), 7)
    assert gi.tolist() == list(range(7)) and (gd == 0.0).all()
    # deleting a zero-norm row removes it
    assert st.delete_chunks([700]) == 1
    st.build_index()
    gi, _ = st.search_ids(q, 3)
    assert gi[:2].tolist() == [5, 1999]


def test_nonfinite_rows_dropped(cs):
    rows = np.eye(8, dtype=np.float32)
    rows[2, 1] = np.nan
    rows[4, 0] = np.inf
    st = make_store(cs, rows)
    s = st.device_stats()
    assert s.nonfinite_rows == 2 and s.live_rows == 6
    gi, _ = st.search_ids(np.ones(8, np.float32), 8)
    assert sorted(gi.tolist()) == [0, 1, 3, 5, 6, 7]


def test_negative_and_opposite(cs):
    rows = np.array([[1, 0, 0, 0], [-1, 0, 0, 0], [0, 1, 0, 0]], dtype=np.float32)
    st = make_store(cs, rows)
    gi, gd = st.search_ids([2, 0, 0, 0], 3)
    assert gi.tolist() == [0, 2, 1]
    assert gd.tolist() == [0.0, 0.5, 1.0]


# ---- mutation protocol (store.rs:548-686) ------------------------------------------------------
def test_append_remove_rebuild_matches_oracle(cs, oracle):
    rng = np.random.default_rng(9)
    d = 384
    rows = rng.standard_normal((6000, d)).astype(np.float32)
    ids = np.arange(6000, dtype=np.uint32)
    st = cs.VectorStore.new(None, d)
    st.append_rows(rows[:4000], ids[:4000])
    st.build_index()
    dead = rng.permutation(4000)[:1500].astype(np.uint32)
    assert st.delete_chunks(dead) == 1500
    assert st.delete_chunks(dead[:10]) == 0                     # already gone
    st.append_rows(rows[4000:], ids[4000:])
    st.build_index()
    keep = np.ones(6000, bool)
    keep[dead] = False
    assert st.device_stats().live_rows == keep.sum()
    for k in (10, 150):
        q = rng.standard_normal(d).astype(np.float32)
        gi, gd = st.search_ids(q, k)
        oi, od, o64 = oracle.np_search(rows[keep], q, k + MARGIN, ids=ids[keep])
        check_topk(gi, gd, oi, od, o64, k)
    # replace semantics: re-appending a live id supersedes the old vector
    new_vec = rng.standard_normal((1, d)).astype(np.float32)
    live_id = int(ids[keep][0])
    st.append_rows(new_vec, np.array([live_id], np.uint32))
    st.build_index()
    assert st.device_stats().live_rows == keep.sum()
    gi, gd = st.search_ids(new_vec[0], 1)
    assert gi.tolist() == [live_id] and gd[0] <= 1e-6
    # duplicate ids inside one batch: the last wins
    st2 = cs.VectorStore.new(None, 4)
    st2.append_rows(np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]], np.float32), np.array([7, 3, 7], np.uint32))
    st2.build_index()
    gi, gd = st2.search_ids([0, 0, 1, 0], 3)
    assert gi.tolist() == [7, 3] and gd[0] == 0.0


def test_empty_index_and_k0(cs):
    st = cs.VectorStore.new(None, 384)
    st.build_index()
    gi, gd = st.search_ids(np.ones(384, np.float32), 10)
    assert len(gi) == 0
    st.append_rows(np.ones((3, 384), np.float32), np.arange(3, dtype=np.uint32))
    st.build_index()
    gi, _ = st.search_ids(np.ones(384, np.float32), 0)
    assert len(gi) == 0
    with pytest.raises(cs.CsgpuError):
        st.search_ids(np.ones(384, np.float32), 1025)


# ---- filtered variant (new) -----------------------------------------------------
@pytest.mark.parametrize("density", [1.0, 0.25, 0.01, 0.0])
def test_filtered_parity(cs, oracle, density):
    rng = np.random.default_rng(int(density * 1000) + 11)
    n, d = 20000, 384
    rows = rng.standard_normal((n, d)).astype(np.float32)
    ids = rng.permutation(n * 2)[:n].astype(np.uint32)
    st = make_store(cs, rows, ids)
    allowed = rng.random(n * 2) < density
    flt = cs.RowFilter.from_mask(allowed)
    for k in (10, 200):
        q = rng.standard_normal(d).astype(np.float32)
        assert_parity(oracle, st, rows, q, k, ids=ids, allowed=allowed, flt=flt)
    # superset-preserving w.r.t. the reference's host post-filter (src/search/mod.rs:727-737)
    q = rng.standard_normal(d).astype(np.float32)
    full, _ = st.search_ids(q, 1024)
    filt, _ = st.search_ids(q, 50, flt)
    post = [i for i in full.tolist() if allowed[i]]
    assert filt.tolist()[:len(post[:50])] == post[:50][:len(filt)]
    # n_bits smaller than the id space: ids beyond it are excluded
    small = cs.RowFilter(np.full(4, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64), 200)
    gi, _ = st.search_ids(q, 1000, small)
    assert (gi < 200).all() and len(gi) == int((ids < 200).sum())


# ---- batch -------------------------------------------------------------------------------------
def test_empty_filters_allow_nothing_on_a_fresh_store(cs):
    """Round-1 advisor finding: with n_bits == 0 (or a file bitmap of 0 bits) nothing was uploaded, a fresh context had no
    bitmap pointer, and the launch silently took the UNFILTERED path. The header's contract: an empty filter allows no row
    (ids >= n_bits are excluded) — on a fresh context and on one that held a bitmap before alike, zero-norm rows included."""
    from codesearch_b200.tags import TagPredicate
    rng = np.random.default_rng(5)
    rows = rng.standard_normal((3000, 32)).astype(np.float32)
    rows[7] = 0.0
    for warm in (False, True):
        st = cs.VectorStore.new(None, 32)
        st.append_rows(rows, np.arange(3000, dtype=np.uint32), np.full(3000, 5, dtype=np.uint32))
        st.build_index()
        q = rng.standard_normal(32).astype(np.float32)
        if warm:
            assert len(st.search_ids(q, 10, cs.RowFilter.from_mask(np.ones(3000, bool)))[0]) == 10
        ids, dist = st.search_ids(q, 10, cs.RowFilter(np.zeros(0, dtype=np.uint64), 0))
        assert len(ids) == 0 and len(dist) == 0
        ids, _ = st.search_ids(q, 10, cs.RowFilter(np.zeros(1, dtype=np.uint64), 0))
        assert len(ids) == 0
        ids, _ = st.search_tagged_ids(q, 10, TagPredicate(file_bitmap=np.zeros(1, dtype=np.uint64), n_file_bits=0))
        assert len(ids) == 0
        assert len(st.search_ids(q, 10)[0]) == 10


@pytest.fixture
def scan_route():
    """The tests below pin the multi-query SCAN kernels (scan_multi.cuh). On its own the library routes a batch by a cost
    model and may answer it on the tensor cores instead (tests/test_gpu_tf32_batch.py covers that route)."""
    import os
    old = os.environ.get("CSGPU_GEMM_MIN_BATCH")
    os.environ["CSGPU_GEMM_MIN_BATCH"] = "100000"
    yield
    if old is None:
        os.environ.pop("CSGPU_GEMM_MIN_BATCH", None)
    else:
        os.environ["CSGPU_GEMM_MIN_BATCH"] = old


def test_batch_matches_single(cs, oracle, scan_route):
    rng = np.random.default_rng(12)
    n, d = 10000, 384
    rows = rng.standard_normal((n, d)).astype(np.float32)
    st = make_store(cs, rows)
    qs = rng.standard_normal((9, d)).astype(np.float32)          # <= 9 query variants, src/search/mod.rs:508-511
    oi, od, on = st.search_batch_ids(qs, 100)
    for j in range(9):
        gi, gd = st.search_ids(qs[j], 100)
        # the multi-query kernel is bit-identical to the single-query kernel (same FMA chain, same tree)
        assert on[j] == 100 and np.array_equal(oi[j], gi) and np.array_equal(od[j], gd)
        ri, rd, r64 = oracle.np_search(rows, qs[j], 100 + MARGIN)
        check_topk(oi[j], od[j], ri, rd, r64, 100)


@pytest.mark.parametrize("n,d,b,k", [
    (20000, 384, 2, 10), (20000, 384, 3, 32), (20000, 384, 8, 10), (20000, 384, 9, 33), (20000, 384, 17, 100),
    (20000, 384, 5, 256), (5000, 384, 4, 300), (5000, 768, 8, 200), (5000, 1024, 7, 25), (5000, 128, 16, 64),
    (5000, 256, 3, 10), (5000, 512, 8, 128), (3000, 100, 5, 10), (3, 384, 8, 10), (1, 384, 2, 1),
])
def test_batch_parity(cs, oracle, scan_route, n, d, b, k):
    from codesearch_b200 import _lib
    rng = np.random.default_rng(n + d * 7 + b * 13 + k)
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[n // 2] = 0.0                                           # a zero-norm row rides along
    ids = rng.permutation(n * 2)[:n].astype(np.uint32)
    st = make_store(cs, rows, ids)
    qs = rng.standard_normal((b, d)).astype(np.float32)
    if b > 2:
        qs[1] = 0.0                                              # and a zero-norm query
    launches0 = _lib.load().csgpu_kernel_launches()
    oi, od, on = st.search_batch_ids(qs, k)
    launches = _lib.load().csgpu_kernel_launches() - launches0
    if d % 128 == 0 and k <= 256:                                # one pass per <= 16 queries (8-query kernel up to 8, 16-query kernel above)
        assert launches == (b // 16) + (1 if b % 16 else 0)
    k_eff = min(k, n)
    for j in range(b):
        ri, rd, r64 = oracle.np_search(rows, qs[j], k + MARGIN, ids=ids)
        assert on[j] == k_eff
        check_topk(oi[j, :k_eff], od[j, :k_eff], ri, rd, r64, k_eff)
        gi, gd = st.search_ids(qs[j], k)
        assert np.array_equal(oi[j, :k_eff], gi) and np.array_equal(od[j, :k_eff], gd)


@pytest.mark.parametrize("d", [128, 384, 768])
def test_sixteen_query_pass_bit_identical_to_single(cs, scan_route, d):
    """Round 2: 9..16 queries share ONE pass (scan_multi.cuh, MQ = 16 x R = 2) — the reference's default hybrid search is
    <= 9 query variants x limit 200 (src/search/mod.rs:498-511). Same FMA chain and shuffle tree per (row, query), so every
    list is bit-identical to csgpu_search, for the per-warp lists (k <= 32) and the CTA buffers (k > 32) alike."""
    from codesearch_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(900 + d)
    n = 60_000
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[[3, 40_000]] = 0.0
    st = make_store(cs, rows)
    qs = rng.standard_normal((16, d)).astype(np.float32)
    qs[5] = 0.0
    for b, k in ((9, 200), (16, 10), (12, 32), (10, 33), (16, 256), (13, 100), (16, 1)):
        l0 = lib.csgpu_kernel_launches()
        oi, od, on = st.search_batch_ids(qs[:b], k)
        assert lib.csgpu_kernel_launches() - l0 == 1, (b, k)
        for j in range(b):
            gi, gd = st.search_ids(qs[j], k)
            assert on[j] == len(gi) and np.array_equal(oi[j, : on[j]], gi) and np.array_equal(od[j, : on[j]].view(np.uint32), gd.view(np.uint32)), (b, k, j)
    # the variants entry point: 9 variants x top-200 = one pass + the dedup kernel
    l0 = lib.csgpu_kernel_launches()
    vi, vd = st.search_variants_ids(qs[6:15], 200)
    assert lib.csgpu_kernel_launches() - l0 == 2
    best = {}
    for q in qs[6:15]:
        for i, x in zip(*st.search_ids(q, 200)):
            best[int(i)] = min(best.get(int(i), 9.0), float(x))
    want = sorted(best.items(), key=lambda t: (t[1], t[0]))[:200]
    assert [int(i) for i in vi] == [w[0] for w in want]


# ---- device entry points + cross-shard merge ---------------------------------------------------
def test_device_entry_points_and_merge(cs, oracle):
    import torch
    from codesearch_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(13)
    n, d, k = 30000, 384, 10
    rows = rng.standard_normal((n, d)).astype(np.float32)
    q = rng.standard_normal(d).astype(np.float32)
    # 3 "ranks": row-sharded stores with global ids; local top-k -> concatenated -> merged
    bounds = [0, 9000, 21000, n]
    stores = [make_store(cs, rows[a:b], np.arange(a, b, dtype=np.uint32)) for a, b in zip(bounds, bounds[1:])]
    qd = torch.from_numpy(q).cuda()
    gathered = torch.empty((3, k), dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    for r, st in enumerate(stores):
        _lib.check(lib.csgpu_search_keys_device(st.handle, qd.data_ptr(), k, gathered[r].data_ptr(), stream))
    out = torch.empty(k, dtype=torch.int64, device="cuda")
    _lib.check(lib.csgpu_merge_keys_device(stores[0].handle, gathered.data_ptr(), 3, k, out.data_ptr(), stream))
    torch.cuda.synchronize()
    keys = out.cpu().numpy().view(np.uint64)
    ids = np.zeros(k, np.uint32); dist = np.zeros(k, np.float32); m = ctypes.c_uint32()
    lib.csgpu_decode_keys(keys.ctypes.data_as(_lib._u64p), k, ids.ctypes.data_as(_lib._u32p),
                          dist.ctypes.data_as(_lib._f32p), ctypes.byref(m))
    assert m.value == k
    oi, od, o64 = oracle.np_search(rows, q, k + MARGIN)
    check_topk(ids, dist, oi, od, o64, k)
    # sharded result is bit-identical to the single-index result (fixed reduction tree)
    whole = make_store(cs, rows)
    gi, gd = whole.search_ids(q, k)
    assert np.array_equal(gi, ids) and np.array_equal(gd, dist)


def test_concurrent_searches_from_threads(cs, oracle):
    """search(&self) is called from rayon/tokio threads concurrently (src/search/mod.rs:508-511)."""
    import threading
    rng = np.random.default_rng(14)
    n, d = 50000, 384
    rows = rng.standard_normal((n, d)).astype(np.float32)
    st = make_store(cs, rows)
    qs = rng.standard_normal((9, d)).astype(np.float32)
    want = [oracle.np_search(rows, q, 20)[0] for q in qs]
    errs = []

    def work(j):
        try:
            for _ in range(20):
                gi, _ = st.search_ids(qs[j], 20)
                if not np.array_equal(gi, want[j]):
                    errs.append((j, gi))
        except Exception as e:  # noqa: BLE001
            errs.append((j, e))
    ts = [threading.Thread(target=work, args=(j,)) for j in range(9)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs[:2]


def test_kernels_actually_launch(cs):
    from codesearch_b200 import _lib
    lib = _lib.load()
    st = make_store(cs, np.eye(4, dtype=np.float32))
    before = lib.csgpu_kernel_launches()
    st.search_ids([1, 0, 0, 0], 2)
    assert lib.csgpu_kernel_launches() == before + 1        # one fused scan+top-k+merge kernel per query
    assert st.device_stats().last_search_us > 0


# ---- full size (BASELINE.json configs[1]): 10M x 384, top-10 ------------------------------------
def test_full_size_10m_parity_and_properties(cs, oracle):
    """10M x 384 fp32 on one GPU. Checked three ways: (1) the streaming f64 oracle over the same
    counter-based corpus (ids bit-exact, |d| <= 1e-5); (2) planted neighbours: rows appended with
    known ids must come back first; (3) shard-union property: top-k of the whole == merge of top-k of
    two half stores (size-independent; this is what multi-GPU relies on)."""
    n, d, k = 10_000_000, 384, 10
    st = cs.VectorStore.new(None, d)
    st.reserve(n + 16)
    st.append_synthetic(1234, 0, n)
    st.build_index()
    assert st.device_stats().live_rows == n
    qs = oracle.synth_rows(4321, 0, 4, d)
    oi, od, o64, on = oracle.search_synth(1234, 0, n, d, qs, k + MARGIN)
    swaps = 0
    for j in range(4):
        gi, gd = st.search_ids(qs[j], k)
        swaps += check_topk(gi, gd, oi[j], od[j], o64[j], k)
    assert swaps == 0
    # k = 100 through the big-k path
    gi, gd = st.search_ids(qs[0], 100)
    o2 = oracle.search_synth(1234, 0, n, d, qs[:1], 100 + MARGIN)
    check_topk(gi, gd, o2[0][0], o2[1][0], o2[2][0], 100)
    # (2) planted neighbours at ids beyond the corpus
    rng = np.random.default_rng(21)
    q = qs[1]
    planted = np.stack([q + np.float32(s) * rng.standard_normal(d).astype(np.float32) * np.abs(q).mean()
                        for s in (0.0, 0.05, 0.1, 0.2)]).astype(np.float32)
    pid = np.array([n + 3, n + 1, n + 2, n + 0], dtype=np.uint32)
    st.append_rows(planted, pid)
    st.build_index()
    gi, gd = st.search_ids(q, k)
    assert gi[:4].tolist() == [n + 3, n + 1, n + 2, n + 0]
    assert gd[0] <= 1e-6
    assert gi[4:].tolist() == oi[1][:k - 4].tolist()
