"""The callers either side of the path, through the C ABI. GPU only.

N2  host micro-batcher: concurrent VectorStore.search calls with the same limit share one corpus pass
    (src/search/mod.rs:508-511: rayon par_iter over <= 9 query variants; MCP/HTTP readers).
N3  device-side dedup of the per-variant lists (src/search/mod.rs:513-590).
"""
import threading

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
    st = cs.VectorStore.new(None, rows.shape[1])
    st.append_rows(rows, np.arange(rows.shape[0], dtype=np.uint32) if ids is None else ids)
    st.build_index()
    return st


@pytest.mark.parametrize("shadow", ["none", "byte", "tensor"])
def test_coalescing_concurrent_searches_bit_identical(cs, oracle, shadow):
    """`shadow`: which opt-in shadow is on while the callers are coalesced — a coalesced group is a small batch and takes
    whatever route csgpu_search_batch picks for it (multi-query scan / int8 singles / one tensor-core batch); every
    route must hand each caller exactly what its own uncoalesced csgpu_search returns."""
    rng = np.random.default_rng(51)
    n, d, k = 400_000, 384, 200                                  # hybrid default retrieval limit (search/mod.rs:498-501)
    rows = rng.standard_normal((n, d)).astype(np.float32)
    st = make_store(cs, rows)
    qs = rng.standard_normal((9, d)).astype(np.float32)          # <= 9 query variants
    want = [st.search_ids(q, k) for q in qs]                     # uncoalesced, one fp32 scan each
    if shadow == "byte":
        st.set_byte_prefilter(True)
    elif shadow == "tensor":
        st.set_tensor_prefilter(True)
    st.set_coalescing(True)
    s0 = st.device_stats()
    got = [None] * 9
    errs = []

    def work(j):
        try:
            for _ in range(10):
                got[j] = st.search_ids(qs[j], k)
                if not (np.array_equal(got[j][0], want[j][0]) and np.array_equal(got[j][1], want[j][1])):
                    errs.append(j)
        except Exception as e:  # noqa: BLE001
            errs.append((j, e))
    ts = [threading.Thread(target=work, args=(j,)) for j in range(9)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs[:3]
    s1 = st.device_stats()
    q_done = s1.coalesced_queries - s0.coalesced_queries
    passes = s1.coalesced_passes - s0.coalesced_passes
    assert q_done == 90
    assert passes < 90, "concurrent searches were never batched"
    print(f"coalescer: {q_done} searches in {passes} corpus passes")
    # mixed limits: requests with different k never share a pass, all still answered correctly
    ks = [10, 200, 10, 33, 200, 10, 33, 10, 200]
    want2 = [st.search_ids(qs[j], ks[j]) for j in range(9)]
    got2 = [None] * 9

    def work2(j):
        got2[j] = st.search_ids(qs[j], ks[j])
    ts = [threading.Thread(target=work2, args=(j,)) for j in range(9)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for j in range(9):
        assert np.array_equal(got2[j][0], want2[j][0]) and np.array_equal(got2[j][1], want2[j][1])
    # errors reach the caller that made them (guards verbatim, store.rs:432-438), others are unaffected
    with pytest.raises(cs.CsgpuError) as e:
        st.search_ids(qs[0][:100], k)
    assert e.value.code == 1 and "expected 384, got 100" in str(e.value)
    st.set_coalescing(False)
    gi, gd = st.search_ids(qs[0], k)
    assert np.array_equal(gi, want[0][0])
    oi, od, o64 = oracle.np_search(rows, qs[0], k + MARGIN)
    check_topk(gi, gd, oi, od, o64, k)


def test_many_threads_of_multi_query_scans(cs):
    """24 host threads run multi-query scans at once on one index. The scans' last CTAs wait, resident, for the rest of their
    grid (scan_multi.cuh finishers), so the library bounds how many run at a time (MultiSlot): all of them finish, every
    list equals the single-query kernel's."""
    rng = np.random.default_rng(78)
    n, d, T = 60_000, 384, 24
    rows = rng.standard_normal((n, d)).astype(np.float32)
    st = make_store(cs, rows)
    qs = rng.standard_normal((T, 9, d)).astype(np.float32)
    ks = [10, 50, 200]
    want = {(j, k): [st.search_ids(qs[j, v], k) for v in range(9)] for j in range(0, T, 6) for k in ks}
    errs = []

    def work(j):
        try:
            for rep in range(6):
                k = ks[(j + rep) % 3]
                oi, od, on = st.search_batch_ids(qs[j], k)             # 9 queries: one 16-query pass (small corpus: scan route)
                vi, vd = st.search_variants_ids(qs[j, :5], k)           # and the variants path
                if (j, k) in want:
                    for v in range(9):
                        if not (np.array_equal(oi[v], want[(j, k)][v][0]) and np.array_equal(od[v].view(np.uint32), want[(j, k)][v][1].view(np.uint32))):
                            errs.append((j, k, v))
        except Exception as e:  # noqa: BLE001
            errs.append((j, repr(e)))
    ts = [threading.Thread(target=work, args=(j,), daemon=True) for j in range(T)]
    [t.start() for t in ts]
    [t.join(timeout=120) for t in ts]
    assert not any(t.is_alive() for t in ts), "multi-query scans hung"
    assert not errs, errs[:3]


def test_coalescing_many_callers_share_one_tensor_core_batch(cs, oracle):
    """Round 2: on an fp32 index whose batches run the tf32 tensor-core filter (csrc/gemm_tf32.cuh) a coalesced group is up
    to 128 callers — one 128-query block costs about what one query costs. 48 threads hammer csgpu_search (the reference's
    server shape: `&self` from many threads, src/server/mod.rs:545-548); every caller gets exactly what its own uncoalesced
    search returns, and the searches ride in far fewer passes than a 16-wide group would allow."""
    rng = np.random.default_rng(77)
    n, d, k, T, R = 1_500_000, 384, 10, 48, 6
    st = cs.VectorStore.new(None, d)
    st.append_synthetic(99, 0, n)
    st.build_index()
    qs = rng.standard_normal((T, d)).astype(np.float32)
    want = [st.search_ids(q, k) for q in qs]
    st.set_coalescing(True)
    s0 = st.device_stats()
    errs = []
    start = threading.Barrier(T)

    def work(j):
        try:
            start.wait()
            for _ in range(R):
                gi, gd = st.search_ids(qs[j], k)
                if not (np.array_equal(gi, want[j][0]) and np.array_equal(gd.view(np.uint32), want[j][1].view(np.uint32))):
                    errs.append(j)
        except Exception as e:  # noqa: BLE001
            errs.append((j, e))
    ts = [threading.Thread(target=work, args=(j,)) for j in range(T)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs[:3]
    s1 = st.device_stats()
    q_done = s1.coalesced_queries - s0.coalesced_queries
    passes = s1.coalesced_passes - s0.coalesced_passes
    print(f"coalescer: {q_done} searches in {passes} passes, last batch route {s1.batch_route}")
    assert q_done == T * R
    assert passes < q_done / 2, (passes, q_done)                  # callers really shared passes (typically ~30 per pass)
    assert s1.batch_route in (0, 3)                               # groups of >= 9 ran as tf32 tensor-core batches (0: none formed)
    st.set_coalescing(False)


@pytest.mark.parametrize("n,d,b,k", [(50000, 384, 9, 200), (50000, 384, 3, 10), (20000, 768, 16, 100),
                                     (5000, 100, 5, 50), (300, 384, 9, 1000), (20000, 384, 1, 25)])
def test_search_variants_dedup(cs, oracle, n, d, b, k):
    rng = np.random.default_rng(n + d + b + k)
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[3] = 0.0
    ids = rng.permutation(2 * n)[:n].astype(np.uint32)
    st = make_store(cs, rows, ids)
    base = rng.standard_normal(d).astype(np.float32)
    # variants of one query: the original plus perturbed copies (their result lists overlap heavily)
    qs = np.stack([base] + [base + np.float32(0.3) * rng.standard_normal(d).astype(np.float32) for _ in range(b - 1)])
    gi, gd = st.search_variants_ids(qs, k)
    lists = [st.search_ids(q, k) for q in qs]                    # what the reference's par_iter would collect
    wi, wd = oracle.dedup_variants(lists, k)
    assert np.array_equal(gi, wi) and np.array_equal(gd, wd)
    assert len(set(gi.tolist())) == len(gi)
    # and against the oracle end to end: per-variant exact lists -> dedup
    olists = [oracle.np_search(rows, q, k, ids=ids)[:2] for q in qs]
    oi, od = oracle.dedup_variants(olists, k)
    assert np.array_equal(gi, oi)
    assert np.abs(gd - od).max() <= 1e-5
