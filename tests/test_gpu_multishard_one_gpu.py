"""The in-process multi-device index with BOTH shards on device 0 (csgpu_create accepts a repeated ordinal): the whole
multi-shard host logic — concurrent per-shard launches, peer copies, k-way merge, per-shard dedup of query variants,
the byte prefilter's per-shard status words — runs on the one-GPU test box. Results must equal the single-shard index
bit for bit. (tests/test_gpu_multidevice_batch.py is the same on two real GPUs.)"""
import os

import numpy as np
import pytest

from parity import check_topk

pytestmark = pytest.mark.gpu

MARGIN = 8
os.environ.setdefault("CSGPU_I8_MIN_ROWS", "4096")


@pytest.fixture(scope="module")
def cs():
    import codesearch_b200 as m
    m.load_library()
    return m


@pytest.fixture(scope="module")
def pair(cs):
    rng = np.random.default_rng(77)
    n, d = 50_000, 384
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[[5, 30_000]] = 0.0
    out = []
    for devs in ([0], [0, 0]):
        st = cs.VectorStore.new(None, d, devices=devs)
        st.append_rows(rows, np.arange(n, dtype=np.uint32))
        st.build_index()
        out.append(st)
    assert out[1].device_stats().n_devices == 2 and min(out[1].device_stats().rows_per_device[:2]) > 0
    return rows, out[0], out[1], rng


def _same(a, b):
    return np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))


def test_two_shards_every_search_entry_point(pair, oracle):
    rows, one, two, rng = pair
    d = rows.shape[1]
    qs = rng.standard_normal((130, d)).astype(np.float32)
    # batch through the SIMT GEMM path with a small k first: it used to leave merge_keys_kernel's dynamic shared-memory
    # ceiling at ITS size, and the larger merge of the next single query then failed to launch ("invalid argument")
    a = one.search_batch_ids(qs, 50)
    c = two.search_batch_ids(qs, 50)
    assert np.array_equal(a[2], c[2]) and _same(a, c)
    for k in (10, 200, 1000):
        g, h = two.search_ids(qs[1], k), one.search_ids(qs[1], k)
        assert _same(g, h), k
        oi, od, o64 = oracle.np_search(rows, qs[1], k + MARGIN)
        check_topk(g[0], g[1], oi, od, o64, k)
    for b, k in ((8, 10), (5, 100), (9, 256), (16, 256), (12, 10), (24, 100)):   # multi-query scan: 8- and 16-query passes, 16 + 8
        a = one.search_batch_ids(qs[:b], k)
        c = two.search_batch_ids(qs[:b], k)
        assert np.array_equal(a[2], c[2]) and _same(a, c), (b, k)
    flt = cs_filter(rows.shape[0])
    assert _same(two.search_ids(qs[2], 100, flt), one.search_ids(qs[2], 100, flt))


def cs_filter(n):
    import codesearch_b200 as m
    return m.RowFilter.from_mask(np.arange(n) % 5 != 0)


@pytest.mark.parametrize("b,k", [(9, 200), (3, 10), (16, 100), (16, 1024)])
def test_two_shards_search_variants(pair, oracle, b, k):
    rows, one, two, rng = pair
    d = rows.shape[1]
    base = rng.standard_normal(d).astype(np.float32)
    qs = np.stack([base] + [base + np.float32(0.3) * rng.standard_normal(d).astype(np.float32) for _ in range(b - 1)])
    a = one.search_variants_ids(qs, k)
    c = two.search_variants_ids(qs, k)
    assert _same(a, c)
    lists = [one.search_ids(q, k) for q in qs]
    wi, wd = oracle.dedup_variants(lists, k)
    assert np.array_equal(c[0], wi) and np.array_equal(c[1], wd)


def test_two_shards_byte_prefilter(pair):
    rows, one, two, rng = pair
    d = rows.shape[1]
    two.set_byte_prefilter(True)
    try:
        qs = rng.standard_normal((5, d)).astype(np.float32)
        s0 = two.device_stats()
        for j, k in enumerate((10, 1, 100, 128, 10)):
            assert _same(two.search_ids(qs[j], k), one.search_ids(qs[j], k)), k
        s1 = two.device_stats()
        assert s1.byte_searches - s0.byte_searches == 5 and s1.byte_fallbacks - s0.byte_fallbacks <= 1
        z = np.zeros(d, np.float32)                              # zero-norm query: both shards report it, fp32 kernel answers
        assert _same(two.search_ids(z, 10), one.search_ids(z, 10))
        assert two.device_stats().byte_fallbacks == s1.byte_fallbacks + 1
    finally:
        two.set_byte_prefilter(False)


def test_in_process_fused_route_and_launch_count(cs, pair):
    """csgpu_search on an N-device index = N scan launches (gather exchange fused into their tails), nothing else: no peer
    memcpy, no merge launch. Also from several host threads at once (&self + rayon in the reference,
    src/search/mod.rs:508-511): every thread takes its own wired context group."""
    import threading
    from codesearch_b200 import _lib
    lib = _lib.load()
    rows, one, two, rng = pair
    d = rows.shape[1]
    qs = rng.standard_normal((12, d)).astype(np.float32)
    two.search_ids(qs[0], 10)                                    # group creation + kernel preload happen here
    for k in (10, 100):
        l0 = lib.csgpu_kernel_launches()
        got = two.search_ids(qs[1], k)
        assert lib.csgpu_kernel_launches() - l0 == 2, "expected exactly one scan launch per shard"
        assert _same(got, one.search_ids(qs[1], k))
    want = [one.search_ids(q, 50) for q in qs]
    bad = []

    def worker(t):
        for rep in range(6):
            for j in range(t, 12, 4):
                if not _same(two.search_ids(qs[j], 50), want[j]):
                    bad.append((t, rep, j))

    ths = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert not bad


def test_in_process_fused_three_shards_filters_tags(cs, oracle):
    rng = np.random.default_rng(78)
    n, d = 30_000, 100                                           # padded dim, uneven shards
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[11] = 0.0
    ids = np.arange(n, dtype=np.uint32)
    from codesearch_b200 import tags as T
    tg = T.synth_tags(0, n)
    one = cs.VectorStore.new(None, d, devices=[0])
    three = cs.VectorStore.new(None, d, devices=[0, 0, 0])
    for st in (one, three):
        st.append_rows(rows[:20_000], ids[:20_000], tg[:20_000])
        st.append_rows(rows[20_000:20_100], ids[20_000:20_100], tg[20_000:20_100])   # small append: least-loaded shard
        st.append_rows(rows[20_100:], ids[20_100:], tg[20_100:])
        st.delete_chunks(list(range(100, 300)))
        st.build_index()
    per = three.device_stats().rows_per_device[:3]
    assert three.device_stats().n_devices == 3 and min(per) > 0 and len(set(per)) > 1
    q = rng.standard_normal(d).astype(np.float32)
    for k in (1, 10, 64, 300):
        assert _same(three.search_ids(q, k), one.search_ids(q, k)), k
    flt = cs.RowFilter.from_mask(np.arange(n) % 7 != 3)
    assert _same(three.search_ids(q, 40, flt), one.search_ids(q, 40, flt))
    pred = T.TagPredicate(lang_mask=0b1011, file_lo=10, file_hi=600)
    assert _same(three.search_tagged_ids(q, 25, pred), one.search_tagged_ids(q, 25, pred))
    z = np.zeros(d, np.float32)
    assert _same(three.search_ids(z, 10), one.search_ids(z, 10))


def test_in_process_exchange_timeout_is_an_error(cs, pair, monkeypatch):
    """A shard that never delivers (fault hook: its launch is skipped) -> CSGPU_ERR_NCCL after the configured bound, the
    poisoned group is retired, and the next search (fresh group) is correct again."""
    from codesearch_b200 import _lib
    lib = _lib.load()
    rows, one, two, rng = pair
    q = rng.standard_normal(rows.shape[1]).astype(np.float32)
    _lib.check(lib.csgpu_exchange_set_timeout_ms(two.handle, 100))
    monkeypatch.setenv("CSGPU_FAULT_SKIP_SHARD", "1")
    with pytest.raises(_lib.CsgpuError) as ei:
        two.search_ids(q, 10)
    assert ei.value.code == _lib.ERR_NCCL and "timed out" in str(ei.value)
    monkeypatch.delenv("CSGPU_FAULT_SKIP_SHARD")
    assert _same(two.search_ids(q, 10), one.search_ids(q, 10))
    _lib.check(lib.csgpu_exchange_set_timeout_ms(two.handle, 4000))


def test_uneven_shards_snapshot_roundtrip(cs, tmp_path):
    """Round-1 advisor finding: the version-1 snapshot checksum folded ids/rows/tags shard by shard at each shard's real
    size but was re-folded at an even split on load, so a multi-device store could not reopen its own DB. Version 2 hashes
    per file; saving with uneven shards and loading with 1, 2 or 3 devices all verify and return identical results."""
    rng = np.random.default_rng(79)
    n, d = 9001, 64
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[17] = 0.0
    db = str(tmp_path / "uneven.db")
    st = cs.VectorStore.new(db, d, devices=[0, 0])
    st.append_rows(rows[:8192], np.arange(8192, dtype=np.uint32))
    st.append_rows(rows[8192:], np.arange(8192, n, dtype=np.uint32))        # small append -> one shard only
    st.delete_chunks([1, 2, 3, 5000])
    st.build_index()
    per = st.device_stats().rows_per_device[:2]
    assert per[0] != per[1]
    q = rng.standard_normal(d).astype(np.float32)
    want = st.search_ids(q, 50)
    st.close()
    for devs in ([0, 0], [0], [0, 0, 0]):
        st2 = cs.VectorStore.new(db, d, devices=devs)
        assert st2.is_indexed()
        assert _same(st2.search_ids(q, 50), want), devs
        st2.close()
