"""In-process multi-device index (csgpu_create with n devices) on the batched paths: every shard contracts the whole
batch against its rows concurrently, per-shard lists are gathered on device 0 and merged by key. Results must equal the
single-device index bit for bit (each (query, row) score is position-independent). Needs >= 2 GPUs."""
import numpy as np
import pytest

from parity import check_topk

pytestmark = pytest.mark.gpu

MARGIN = 8


@pytest.fixture(scope="module")
def cs():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import codesearch_b200 as m
    m.load_library()
    return m


def _pair(cs, rows, dtype="fp32", prefilter=False):
    out = []
    for devs in ([0], [0, 1]):
        st = cs.VectorStore.new(None, rows.shape[1], devices=devs, dtype=dtype)
        st.append_rows(rows, np.arange(rows.shape[0], dtype=np.uint32))
        if prefilter:
            st.set_tensor_prefilter(True)
        st.build_index()
        out.append(st)
    return out


@pytest.mark.parametrize("mode", ["simt", "prefilter", "bf16"])
def test_multidevice_batch_equals_single_device(cs, oracle, mode):
    rng = np.random.default_rng({"simt": 1, "prefilter": 2, "bf16": 3}[mode])
    n, d, b, k = 60_000, 384, 130, 50
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[[5, 40_000]] = 0.0 if mode != "bf16" else rows[[5, 40_000]]     # zero-norm rows (fp32 index only)
    one, two = _pair(cs, rows, dtype="bf16" if mode == "bf16" else "fp32", prefilter=mode == "prefilter")
    assert two.device_stats().n_devices == 2 and min(two.device_stats().rows_per_device[:2]) > 0
    qs = rng.standard_normal((b, d)).astype(np.float32)
    a = one.search_batch_ids(qs, k)
    c = two.search_batch_ids(qs, k)
    assert np.array_equal(a[2], c[2])
    assert np.array_equal(a[0], c[0]) and np.array_equal(a[1].view(np.uint32), c[1].view(np.uint32))
    if mode != "bf16":
        for j in (0, b - 1):
            ri, rd, r64 = oracle.np_search(rows, qs[j], k + MARGIN)
            check_topk(c[0][j], c[1][j], ri, rd, r64, k)
        gi, gd = two.search_ids(qs[3], k)                                    # single-query path of the 2-device index
        assert np.array_equal(gi, c[0][3])
    else:
        gi, gd = two.search_ids(qs[3], k)                                    # bf16 single query = a batch of one
        assert np.array_equal(gi, c[0][3]) and np.array_equal(gd, c[1][3])


@pytest.mark.parametrize("b,k", [(9, 200), (3, 10), (16, 100)])
def test_multidevice_search_variants_equals_single_device(cs, oracle, b, k):
    """csgpu_search_variants on a 2-device index: per-shard dedup + k-way merge == the single-device dedup, bit for bit,
    and == the oracle's restatement of the HashMap/BinaryHeap pass (search/mod.rs:513-590)."""
    rng = np.random.default_rng(b * 1000 + k)
    n, d = 40_000, 384
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[7] = 0.0
    one, two = _pair(cs, rows)
    base = rng.standard_normal(d).astype(np.float32)
    qs = np.stack([base] + [base + np.float32(0.3) * rng.standard_normal(d).astype(np.float32) for _ in range(b - 1)])
    ai, ad = one.search_variants_ids(qs, k)
    ci, cd = two.search_variants_ids(qs, k)
    assert np.array_equal(ai, ci) and np.array_equal(ad.view(np.uint32), cd.view(np.uint32))
    lists = [one.search_ids(q, k) for q in qs]
    wi, wd = oracle.dedup_variants(lists, k)
    assert np.array_equal(ci, wi) and np.array_equal(cd, wd)


def test_multidevice_fused_single_query_on_two_real_gpus(cs, oracle):
    """Round 2: csgpu_search / _filtered / _tagged on a 2-device index are two scan launches with the gather exchange fused
    into their tails (peer stores from device 1 into device 0's HBM over NVLink, device 0 merges and writes mapped host
    memory). Bit-identical to the single-device index; exactly one launch per device; concurrent host threads."""
    import threading
    from codesearch_b200 import _lib
    from codesearch_b200 import tags as T
    lib = _lib.load()
    rng = np.random.default_rng(2024)
    n, d = 120_000, 384
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[[9, 100_000]] = 0.0
    tg = T.synth_tags(0, n)
    pair = []
    for devs in ([0], [0, 1]):
        st = cs.VectorStore.new(None, d, devices=devs)
        st.append_rows(rows, np.arange(n, dtype=np.uint32), tg)
        st.build_index()
        pair.append(st)
    one, two = pair
    qs = rng.standard_normal((8, d)).astype(np.float32)
    two.search_ids(qs[0], 10)
    for k in (1, 10, 100, 1000):
        l0 = lib.csgpu_kernel_launches()
        g = two.search_ids(qs[1], k)
        assert lib.csgpu_kernel_launches() - l0 == 2
        h = one.search_ids(qs[1], k)
        assert np.array_equal(g[0], h[0]) and np.array_equal(g[1].view(np.uint32), h[1].view(np.uint32)), k
    ri, rd, r64 = oracle.np_search(rows, qs[1], 100 + MARGIN)
    g = two.search_ids(qs[1], 100)
    check_topk(g[0], g[1], ri, rd, r64, 100)
    flt = cs.RowFilter.from_mask(np.arange(n) % 4 != 0)
    a, b = two.search_ids(qs[2], 50, flt), one.search_ids(qs[2], 50, flt)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    pred = T.TagPredicate(lang_mask=0x00FF, file_lo=5, file_hi=2500)
    a, b = two.search_tagged_ids(qs[3], 200, pred), one.search_tagged_ids(qs[3], 200, pred)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    want = [one.search_ids(q, 20) for q in qs]
    bad = []

    def worker(t):
        for rep in range(20):
            j = (t + rep) % 8
            g = two.search_ids(qs[j], 20)
            if not (np.array_equal(g[0], want[j][0]) and np.array_equal(g[1], want[j][1])):
                bad.append((t, rep))

    ths = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert not bad
