"""Tensor prefilter on the fp32 index (csgpu_set_tensor_prefilter, csrc/rescore.cuh): batches are contracted on the
tensor cores against a bf16 shadow as a FILTER with a proven margin; survivors are rescored from the fp32 rows with the
single-query kernel's arithmetic. The bar is stronger than the 1e-5 tolerance: ids AND distances bit-identical to
csgpu_search on every query, and oracle parity like every other fp32 path. GPU box only.
"""
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


def _store(cs, rows, ids=None, prefilter=True):
    st = cs.VectorStore.new(None, rows.shape[1])
    st.append_rows(rows, np.arange(rows.shape[0], dtype=np.uint32) if ids is None else ids)
    if prefilter:
        st.set_tensor_prefilter(True)
    st.build_index()
    return st


def _assert_batch_equals_single(st, qs, k):
    oi, od, on = st.search_batch_ids(qs, k)
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
    (30_000, 320, 40, 200),      # dim4 = 80: predicated lanes (not a multiple of 32 float4)
    (30_000, 512, 17, 1000),
    (5_000, 64, 9, 10),
    (5_000, 64, 2, 10),          # smallest batch routed to the tensor cores
    (300, 384, 50, 500),         # k > rows
])
def test_prefilter_bit_identical_to_single_query(cs, oracle, n, d, b, k):
    rng = np.random.default_rng(n + d + b + k)
    rows = rng.standard_normal((n, d)).astype(np.float32)
    st = _store(cs, rows)
    s = st.device_stats()
    assert s.shadow_bytes == n * d * 2
    qs = rng.standard_normal((b, d)).astype(np.float32)
    launches0 = cs._lib.load().csgpu_kernel_launches()
    oi, od, on = _assert_batch_equals_single(st, qs, k)
    assert st.device_stats().prefilter_rescored > 0          # the tensor path really ran
    for j in (0, b - 1):
        ri, rd, r64 = oracle.np_search(rows, qs[j], min(k, n) + MARGIN)
        check_topk(oi[j, : on[j]], od[j, : on[j]], ri, rd, r64, min(k, n))


def test_prefilter_near_ties_inside_the_margin(cs, oracle):
    """Adversarial for the filter: thousands of rows whose distance to the query differs by far less than the bf16
    margin (clusters of near-duplicates), so the margin lets many candidates through and the order is decided only
    by the fp32 rescoring."""
    rng = np.random.default_rng(7)
    d, n_clusters, per = 384, 40, 600
    centres = rng.standard_normal((n_clusters, d)).astype(np.float32)
    # noise 0.05: distances inside a cluster spread over ~5e-4 (resolvable in fp32, spacing >> 4e-7) yet all 600 rows
    # of the query's cluster sit inside the 4.0e-3 bf16 margin of each other
    rows = (np.repeat(centres, per, axis=0) + 0.05 * rng.standard_normal((n_clusters * per, d))).astype(np.float32)
    rows[100] = rows[99]                                       # exact duplicates: tie broken by id
    rows[5000] = rows[4999]
    perm = rng.permutation(rows.shape[0])
    rows = rows[perm]
    st = _store(cs, rows)
    qs = np.concatenate([centres[:20] + 0.05 * rng.standard_normal((20, d)), rng.standard_normal((30, d))]).astype(np.float32)
    for k in (10, 100):
        oi, od, on = _assert_batch_equals_single(st, qs, k)
        for j in (0, 5, 25):
            ri, rd, r64 = oracle.np_search(rows, qs[j], k + 64)
            check_topk(oi[j, : on[j]], od[j, : on[j]], ri, rd, r64, k)
    assert st.device_stats().prefilter_rescored > 50 * 300     # the margin really let the clusters through


def test_prefilter_huge_cluster_takes_the_careful_path(cs, oracle):
    """Boilerplate-like corpus: 20 000 near-duplicate rows (licence headers, generated code) all inside the bf16 margin of
    a query. Far more candidates than a segment or the sort buffer holds: the optimistic run raises the overflow flag and
    the batch is repeated in the careful mode (ranges halved until they fit). Slow, but exact."""
    rng = np.random.default_rng(17)
    d, n_dup, n_other = 384, 20_000, 30_000
    centre = rng.standard_normal(d).astype(np.float32)
    dup = (centre + 0.02 * rng.standard_normal((n_dup, d))).astype(np.float32)
    rows = np.concatenate([dup, rng.standard_normal((n_other, d)).astype(np.float32)])
    rows = rows[rng.permutation(rows.shape[0])]
    st = _store(cs, rows)
    qs = np.concatenate([(centre + 0.02 * rng.standard_normal((8, d))), rng.standard_normal((8, d))]).astype(np.float32)
    for k in (10, 100):
        oi, od, on = _assert_batch_equals_single(st, qs, k)
        for j in (0, 12):
            ri, rd, r64 = oracle.np_search(rows, qs[j], k + 64)
            check_topk(oi[j, : on[j]], od[j, : on[j]], ri, rd, r64, k)


def test_prefilter_zero_norm_rows_and_queries(cs, oracle):
    rng = np.random.default_rng(8)
    n, d = 20_000, 384
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[[3, 19_000]] = 0.0
    st = _store(cs, rows)
    qs = rng.standard_normal((48, d)).astype(np.float32)
    qs[7] = 0.0                                                # zero-norm query: every distance 0.0, ids ascending
    oi, od, on = _assert_batch_equals_single(st, qs, 10)
    assert oi[0, 0] == 3 and od[0, 0] == 0.0 and oi[0, 1] == 19_000
    assert np.array_equal(oi[7], np.arange(10)) and (od[7] == 0.0).all()


def test_prefilter_follows_rebuild_toggle_and_snapshot(cs, oracle, tmp_path):
    rng = np.random.default_rng(9)
    n, d = 40_000, 384
    rows = rng.standard_normal((n, d)).astype(np.float32)
    st = cs.VectorStore.new(str(tmp_path / "db"), d)
    st.append_rows(rows[:30_000], np.arange(30_000, dtype=np.uint32))
    st.build_index()
    qs = rng.standard_normal((64, d)).astype(np.float32)
    plain = st.search_batch_ids(qs, 50)                         # fp32 SIMT path
    st.set_tensor_prefilter(True)                               # enabling after build creates the shadow now
    assert st.device_stats().shadow_bytes == 30_000 * d * 2
    a = _assert_batch_equals_single(st, qs, 50)
    assert np.array_equal(a[0], plain[0])                       # same ids as the SIMT path
    assert np.abs(a[1] - plain[1]).max() <= 1e-6
    st.delete_chunks(np.arange(0, 30_000, 3, dtype=np.uint32))
    st.append_rows(rows[30_000:], np.arange(30_000, n, dtype=np.uint32))
    st.build_index()                                            # shadow rebuilt with the compacted rows
    assert st.device_stats().shadow_bytes == st.device_stats().live_rows * d * 2
    _assert_batch_equals_single(st, qs, 50)
    live = np.ones(n, bool); live[0:30_000:3] = False
    ri, rd, r64 = oracle.np_search(rows[live], qs[0], 50 + MARGIN, ids=np.nonzero(live)[0].astype(np.uint32))
    oi, od, on = st.search_batch_ids(qs, 50)
    check_topk(oi[0], od[0], ri, rd, r64, 50)
    st2 = cs.VectorStore.new(str(tmp_path / "db"), d)           # hydrate from the snapshot, then opt in
    st2.set_tensor_prefilter(True)
    b = st2.search_batch_ids(qs, 50)
    assert np.array_equal(b[0], oi) and np.array_equal(b[1], od)
    st.set_tensor_prefilter(False)
    assert st.device_stats().shadow_bytes == 0
    c = st.search_batch_ids(qs, 50)
    assert np.array_equal(c[0], oi)


def test_prefilter_rejects_unsupported(cs):
    st = cs.VectorStore.new(None, 100)
    with pytest.raises(cs.CsgpuError):
        st.set_tensor_prefilter(True)                           # dim % 64 != 0
    st = cs.VectorStore.new(None, 384, dtype="bf16")
    with pytest.raises(cs.CsgpuError):
        st.set_tensor_prefilter(True)


def test_prefilter_full_size_10m(cs, oracle):
    """BASELINE configs[2] at full size: 10M x 384 fp32 index + tensor prefilter, 256 queries x top-100 in one batch.
    Size-independent property: every query's batch result is bit-identical to the single-query scan kernel; two queries
    are also held to the streaming f64 oracle over the same counter-based corpus."""
    n, d, b, k = 10_000_000, 384, 256, 100
    st = cs.VectorStore.new(None, d)
    st.reserve(n)
    st.append_synthetic(1234, 0, n)
    st.set_tensor_prefilter(True)
    st.build_index()
    assert st.device_stats().shadow_bytes == n * d * 2
    qs = oracle.synth_rows(4321, 0, b, d)
    oi, od, on = st.search_batch_ids(qs, k)
    assert (on == k).all()
    for j in range(0, b, 16):
        gi, gd = st.search_ids(qs[j], k)
        assert np.array_equal(oi[j], gi) and np.array_equal(od[j].view(np.uint32), gd.view(np.uint32)), j
    ri, rd, r64, rn = oracle.search_synth(1234, 0, n, d, qs[:2], k + MARGIN)
    for j in range(2):
        check_topk(oi[j], od[j], ri[j], rd[j], r64[j], k)
    assert 100 < st.device_stats().prefilter_rescored / b < 2000      # ~640 fp32 rows read per query, not the corpus


@pytest.mark.parametrize("b,k", [(9, 200), (2, 10), (16, 100)])
def test_prefilter_search_variants_equals_the_scan_route(cs, oracle, b, k):
    """csgpu_search_variants with the tensor prefilter on: the variants run as one tensor-core batch and are deduplicated
    on the host — same result, bit for bit, as the multi-query-scan + device-dedup route (and as the oracle's
    restatement of src/search/mod.rs:513-590)."""
    rng = np.random.default_rng(b * 100 + k)
    n, d = 60_000, 384
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[11] = 0.0
    fast = _store(cs, rows, prefilter=True)
    plain = _store(cs, rows, prefilter=False)
    base = rng.standard_normal(d).astype(np.float32)
    qs = np.stack([base] + [base + np.float32(0.3) * rng.standard_normal(d).astype(np.float32) for _ in range(b - 1)])
    if b == 16:
        qs[5] = 0.0                                              # a zero-norm variant: every distance 0.0, ids ascending
    r0 = fast.device_stats().prefilter_rescored
    gi, gd = fast.search_variants_ids(qs, k)
    pi, pd = plain.search_variants_ids(qs, k)
    assert np.array_equal(gi, pi) and np.array_equal(gd.view(np.uint32), pd.view(np.uint32))
    assert fast.device_stats().prefilter_rescored != r0 or b == 16   # the tensor route really ran
    lists = [plain.search_ids(q, k) for q in qs]
    wi, wd = oracle.dedup_variants(lists, k)
    assert np.array_equal(gi, wi) and np.array_equal(gd, wd)
