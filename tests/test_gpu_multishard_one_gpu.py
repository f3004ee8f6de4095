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
    for b, k in ((8, 10), (5, 100), (9, 256)):                  # multi-query scan, 8 + 1 split
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
