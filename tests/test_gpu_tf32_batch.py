"""Batches on the DEFAULT fp32 index (no opt-in, no shadow copy): gemm_tf32_topk_kernel (csrc/gemm_tf32.cuh) contracts
queries x rows on the tensor cores straight off the fp32 rows (tcgen05 kind::tf32) as a FILTER with a proven margin, and
the survivors are rescored from the same fp32 rows with the single-query kernel's arithmetic (csrc/rescore.cuh).
Bar: ids AND distances bit-identical to csgpu_search on every query, oracle parity like every other fp32 path, and the
largest |d_tf32 - d_f32| the rescoring saw stays below the proven margin TF_MARGIN = 1.1e-3. GPU box only.
BASELINE configs[2]; the reference answers query variants one arroy search at a time (src/search/mod.rs:508-511).
"""
import os

import numpy as np
import pytest

from parity import check_topk

pytestmark = pytest.mark.gpu

MARGIN = 8
TF_MARGIN = 1.1e-3
ROUTE_SIMT, ROUTE_TF32 = 1, 3


@pytest.fixture(scope="module")
def cs():
    import codesearch_b200 as m
    m.load_library()
    return m


@pytest.fixture(autouse=True)
def gemm_route_for_every_batch():
    """Every batch of >= 2 queries takes the GEMM-shaped route (the library's own routing sends small batches over small
    corpora to the multi-query scan), and the SIMT override is off."""
    old = {k: os.environ.get(k) for k in ("CSGPU_GEMM_MIN_BATCH", "CSGPU_BATCH_SIMT")}
    os.environ["CSGPU_GEMM_MIN_BATCH"] = "2"
    os.environ["CSGPU_BATCH_SIMT"] = "0"
    yield
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def _store(cs, rows, ids=None):
    st = cs.VectorStore.new(None, rows.shape[1])
    st.append_rows(rows, np.arange(rows.shape[0], dtype=np.uint32) if ids is None else ids)
    st.build_index()
    return st


def _assert_batch_equals_single(st, qs, k):
    oi, od, on = st.search_batch_ids(qs, k)
    s = st.device_stats()
    assert s.batch_route == ROUTE_TF32, s.batch_route
    assert s.shadow_bytes == 0
    assert 0.0 <= s.filter_max_err < TF_MARGIN, s.filter_max_err
    for j in range(qs.shape[0]):
        gi, gd = st.search_ids(qs[j], k)
        assert on[j] == len(gi), j
        assert np.array_equal(oi[j, : on[j]], gi), (j, k)
        assert np.array_equal(od[j, : on[j]].view(np.uint32), gd.view(np.uint32)), (j, k)   # bit-identical distances
    return oi, od, on


@pytest.mark.parametrize("n,d,b,k", [
    (200_000, 384, 300, 100),
    (200_000, 384, 130, 10),
    (50_000, 128, 64, 33),
    (30_000, 320, 40, 200),      # dim4 = 80: predicated lanes in the rescoring, 10 K chunks
    (30_000, 512, 17, 1000),
    (9_000, 768, 64, 200),       # V = 6 rescoring
    (12_345, 1024, 48, 7),       # the margin's largest dim
    (5_000, 100, 50, 10),        # dim % 32 != 0: the last K chunk is zero-filled by the TMA unit
    (5_000, 36, 41, 33),
    (5_000, 34, 9, 10),          # dim % 4 != 0: rows and queries padded to 36 floats
    (5_000, 64, 2, 10),          # smallest batch: 8 query rows per chunk load
    (300, 384, 50, 500),         # k > rows
    (129, 128, 1024, 5),         # eight query blocks over one tile
    (70_000, 64, 300, 100),
])
def test_tf32_batch_bit_identical_to_single_query(cs, oracle, n, d, b, k):
    rng = np.random.default_rng(n + d + b + k)
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[n // 2] = 0.0                                            # a zero-norm row rides along (distance 0.0)
    rows[n // 3] = rows[n // 3 + 1]                               # and an exact duplicate (ties by id)
    ids = rng.permutation(n * 2)[:n].astype(np.uint32)
    st = _store(cs, rows, ids)
    qs = rng.standard_normal((b, d)).astype(np.float32)
    qs[1] = rows[n // 3]                                          # query equal to the duplicated row
    l0 = cs._lib.load().csgpu_kernel_launches()
    oi, od, on = st.search_batch_ids(qs, k)
    assert cs._lib.load().csgpu_kernel_launches() - l0 <= 40      # phases, not one scan per query
    oi, od, on = _assert_batch_equals_single(st, qs, k)
    assert st.device_stats().prefilter_rescored > 0               # the rescoring really ran
    for j in (0, 1, b - 1):
        ri, rd, r64 = oracle.np_search(rows, qs[j], min(k, n) + MARGIN, ids=ids)
        check_topk(oi[j, : on[j]], od[j, : on[j]], ri, rd, r64, min(k, n))
    a, c = sorted((int(ids[n // 3]), int(ids[n // 3 + 1])))
    if min(k, n) >= 3:                                            # (the zero-norm row, distance 0.0, may sit among them)
        top3 = oi[1, :3].tolist()
        assert a in top3 and c in top3 and top3.index(a) < top3.index(c)
        assert od[1, top3.index(a)] == od[1, top3.index(c)]


def test_tf32_randomised_shapes_against_the_single_query_kernel(cs):
    """Random (rows, dim, batch, k) — including rows < one tile, k > rows, duplicated and zero rows, zero and repeated queries,
    dims that need the zero-filled last K chunk or padded columns: every list of every batch equals csgpu_search's."""
    rng = np.random.default_rng(2024)
    dims = [32, 36, 64, 100, 128, 200, 384, 385, 512, 768, 1000, 1024]
    for it in range(14):
        d = int(rng.choice(dims))
        n = int(rng.choice([1, 7, 255, 256, 257, 1000, 4097, 20_000, 33_333]))
        b = int(rng.choice([2, 3, 8, 9, 31, 127, 128, 129, 260]))
        k = int(rng.choice([1, 2, 10, 32, 33, 100, 256, 257, 700]))
        rows = rng.standard_normal((n, d)).astype(np.float32)
        if n > 10:
            rows[n // 2] = 0.0
            rows[1] = rows[0]
            rows[n - 1] = rows[0] * np.float32(3.0)            # same direction, other norm: equal distance up to rounding
        ids = rng.permutation(3 * n)[:n].astype(np.uint32)
        st = _store(cs, rows, ids)
        qs = rng.standard_normal((b, d)).astype(np.float32)
        qs[0] = rows[0]
        if b > 2:
            qs[2] = 0.0
            qs[b - 1] = qs[1]
        oi, od, on = st.search_batch_ids(qs, k)
        s = st.device_stats()
        assert s.batch_route == ROUTE_TF32 and s.filter_max_err < TF_MARGIN, (it, n, d, b, k, s.batch_route, s.filter_max_err)
        for j in range(b):
            gi, gd = st.search_ids(qs[j], k)
            assert on[j] == len(gi), (it, n, d, b, k, j)
            assert np.array_equal(oi[j, : on[j]], gi) and np.array_equal(od[j, : on[j]].view(np.uint32), gd.view(np.uint32)), (it, n, d, b, k, j)
        st.close()


def test_tf32_route_is_the_default_and_simt_is_the_override(cs):
    rng = np.random.default_rng(3)
    rows = rng.standard_normal((20_000, 384)).astype(np.float32)
    st = _store(cs, rows)
    qs = rng.standard_normal((40, 384)).astype(np.float32)
    os.environ.pop("CSGPU_GEMM_MIN_BATCH", None)                  # the library's own routing: 40 queries -> GEMM-shaped
    a = st.search_batch_ids(qs, 10)
    assert st.device_stats().batch_route == ROUTE_TF32
    os.environ["CSGPU_BATCH_SIMT"] = "1"
    b = st.search_batch_ids(qs, 10)
    assert st.device_stats().batch_route == ROUTE_SIMT
    assert np.array_equal(a[0], b[0]) and np.abs(a[1] - b[1]).max() <= 2e-6


def test_tf32_near_ties_inside_the_margin(cs, oracle):
    """Adversarial for the filter: clusters of near-duplicates whose distances to the query differ by far less than the
    tf32 margin, so the margin lets many candidates through and only the fp32 rescoring decides the order."""
    rng = np.random.default_rng(7)
    d, n_clusters, per = 384, 40, 600
    centres = rng.standard_normal((n_clusters, d)).astype(np.float32)
    rows = (np.repeat(centres, per, axis=0) + 0.05 * rng.standard_normal((n_clusters * per, d))).astype(np.float32)
    rows[100] = rows[99]
    rows[5000] = rows[4999]
    rows = rows[rng.permutation(rows.shape[0])]
    st = _store(cs, rows)
    qs = np.concatenate([centres[:20] + 0.05 * rng.standard_normal((20, d)), rng.standard_normal((30, d))]).astype(np.float32)
    for k in (10, 100):
        oi, od, on = _assert_batch_equals_single(st, qs, k)
        for j in (0, 5, 25):
            ri, rd, r64 = oracle.np_search(rows, qs[j], k + 64)
            check_topk(oi[j, : on[j]], od[j, : on[j]], ri, rd, r64, k)
    assert st.device_stats().prefilter_rescored > 50 * 100        # the margin really let the clusters through


def test_tf32_huge_cluster_takes_the_careful_path(cs, oracle):
    """20 000 near-duplicate rows (licence headers, generated code) all inside the margin of a query: more candidates than
    a segment or the sort buffer holds, so the optimistic run raises the overflow flag and the batch is repeated in the
    careful mode (ranges halved until they fit). Slow, but exact."""
    rng = np.random.default_rng(17)
    d, n_dup, n_other = 384, 20_000, 30_000
    centre = rng.standard_normal(d).astype(np.float32)
    dup = (centre + 0.01 * rng.standard_normal((n_dup, d))).astype(np.float32)
    rows = np.concatenate([dup, rng.standard_normal((n_other, d)).astype(np.float32)])
    rows = rows[rng.permutation(rows.shape[0])]
    st = _store(cs, rows)
    qs = np.concatenate([(centre + 0.01 * rng.standard_normal((8, d))), rng.standard_normal((8, d))]).astype(np.float32)
    for k in (10, 100):
        oi, od, on = _assert_batch_equals_single(st, qs, k)
        for j in (0, 12):
            ri, rd, r64 = oracle.np_search(rows, qs[j], k + 64)
            check_topk(oi[j, : on[j]], od[j, : on[j]], ri, rd, r64, k)


def test_tf32_zero_norm_rows_and_queries(cs, oracle):
    rng = np.random.default_rng(8)
    n, d = 20_000, 384
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[[3, 19_000]] = 0.0
    st = _store(cs, rows)
    qs = rng.standard_normal((48, d)).astype(np.float32)
    qs[7] = 0.0                                                   # zero-norm query: every distance 0.0, ids ascending
    oi, od, on = _assert_batch_equals_single(st, qs, 10)
    assert oi[0, 0] == 3 and od[0, 0] == 0.0 and oi[0, 1] == 19_000
    assert np.array_equal(oi[7], np.arange(10)) and (od[7] == 0.0).all()


def test_tf32_follows_delete_append_rebuild_and_snapshot(cs, oracle, tmp_path):
    rng = np.random.default_rng(9)
    n, d = 40_000, 384
    rows = rng.standard_normal((n, d)).astype(np.float32)
    st = cs.VectorStore.new(str(tmp_path / "db"), d)
    st.append_rows(rows[:30_000], np.arange(30_000, dtype=np.uint32))
    st.build_index()
    qs = rng.standard_normal((64, d)).astype(np.float32)
    _assert_batch_equals_single(st, qs, 50)
    st.delete_chunks(np.arange(0, 30_000, 3, dtype=np.uint32))
    st.append_rows(rows[30_000:], np.arange(30_000, n, dtype=np.uint32))
    st.build_index()                                              # compaction moves the rows: the TMA map is re-encoded
    oi, od, on = _assert_batch_equals_single(st, qs, 50)
    live = np.ones(n, bool); live[0:30_000:3] = False
    ri, rd, r64 = oracle.np_search(rows[live], qs[0], 50 + MARGIN, ids=np.nonzero(live)[0].astype(np.uint32))
    check_topk(oi[0], od[0], ri, rd, r64, 50)
    st2 = cs.VectorStore.new(str(tmp_path / "db"), d)             # hydrate from the snapshot
    b = st2.search_batch_ids(qs, 50)
    assert st2.device_stats().batch_route == ROUTE_TF32
    assert np.array_equal(b[0], oi) and np.array_equal(b[1].view(np.uint32), od.view(np.uint32))


def test_tf32_full_size_10m(cs, oracle):
    """BASELINE configs[2] at full size on the default index: 10M x 384 fp32, 256 queries x top-100 in one batch.
    Size-independent property: every checked query's batch result is bit-identical to the single-query scan kernel; two
    queries are also held to the streaming f64 oracle over the same counter-based corpus."""
    n, d, b, k = 10_000_000, 384, 256, 100
    st = cs.VectorStore.new(None, d)
    st.reserve(n)
    st.append_synthetic(1234, 0, n)
    st.build_index()
    qs = oracle.synth_rows(4321, 0, b, d)
    oi, od, on = st.search_batch_ids(qs, k)
    s = st.device_stats()
    assert s.batch_route == ROUTE_TF32 and s.shadow_bytes == 0 and s.filter_max_err < TF_MARGIN
    assert (on == k).all()
    for j in range(0, b, 16):
        gi, gd = st.search_ids(qs[j], k)
        assert np.array_equal(oi[j], gi) and np.array_equal(od[j].view(np.uint32), gd.view(np.uint32)), j
    ri, rd, r64, rn = oracle.search_synth(1234, 0, n, d, qs[:2], k + MARGIN)
    for j in range(2):
        check_topk(oi[j], od[j], ri[j], rd[j], r64[j], k)
    assert 100 < s.prefilter_rescored / b < 2000                  # ~640 fp32 rows read per query, not the corpus
