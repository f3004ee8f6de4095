"""Byte prefilter on the fp32 index (csgpu_set_byte_prefilter, csrc/scan_i8.cuh): a single query streams a
1-byte-per-element shadow of the corpus as a FILTER with a proven per-row bound; the rows it cannot exclude are rescored
from the fp32 rows with the single-query kernel's arithmetic inside the same launch. The bar is the tensor
prefilter's: ids AND distances bit-identical to the default csgpu_search on every query, oracle parity like every other
fp32 path, and a correct answer (through the fp32 scan kernel) whenever the filter cannot bound a query. GPU box only.
"""
import os

import numpy as np
import pytest

from parity import check_topk

pytestmark = pytest.mark.gpu

MARGIN = 8
os.environ.setdefault("CSGPU_I8_MIN_ROWS", "4096")   # read once by the library: lets small corpora take the int8 route


@pytest.fixture(scope="module")
def cs():
    import codesearch_b200 as m
    m.load_library()
    return m


def _pair(cs, rows, ids=None):
    """(store with the byte prefilter, plain store) over the same rows."""
    out = []
    for on in (True, False):
        st = cs.VectorStore.new(None, rows.shape[1])
        st.append_rows(rows, np.arange(rows.shape[0], dtype=np.uint32) if ids is None else ids)
        if on:
            st.set_byte_prefilter(True)
        st.build_index()
        out.append(st)
    return out


def _assert_same(fast, plain, q, k):
    gi, gd = fast.search_ids(q, k)
    ri, rd = plain.search_ids(q, k)
    assert np.array_equal(gi, ri), (k, gi[:8], ri[:8])
    assert np.array_equal(gd.view(np.uint32), rd.view(np.uint32)), k     # bit-identical distances
    return gi, gd


@pytest.mark.parametrize("n,d,ks", [
    (300_000, 384, (1, 10, 32, 100, 128, 256)),     # the headline shape
    (60_000, 128, (10, 33)),                   # V = 1: 32 rows per warp iteration
    (40_000, 320, (10, 100)),                  # dim4 = 80: padded shadow lines, predicated fp32 lanes
    (30_000, 768, (10, 200)),                  # V = 6 (BASELINE configs[4] width)
    (20_000, 1024, (5, 64)),                   # V = 8
    (9_000, 100, (7,)),                        # dim % 4 == 0 only
    (5_000, 30, (3,)),                         # dim not a multiple of 4 (row padding)
])
def test_byte_prefilter_bit_identical(cs, oracle, n, d, ks):
    rng = np.random.default_rng(n + d)
    rows = rng.standard_normal((n, d)).astype(np.float32)
    fast, plain = _pair(cs, rows)
    s = fast.device_stats()
    assert s.byte_shadow_bytes == n * (128 * ((d + 127) // 128) + 4)
    qs = rng.standard_normal((6, d)).astype(np.float32)
    for k in ks:
        before = fast.device_stats()
        for j in range(qs.shape[0]):
            gi, gd = _assert_same(fast, plain, qs[j], k)
            if j == 0:
                ri, rd, r64 = oracle.np_search(rows, qs[j], k + MARGIN)
                check_topk(gi, gd, ri, rd, r64, k)
        after = fast.device_stats()
        assert after.byte_searches - before.byte_searches == qs.shape[0]     # the int8 route really ran ...
        handed_back = after.byte_fallbacks - before.byte_fallbacks           # ... and answered by itself (a grid that starts
        assert handed_back <= 1, (k, handed_back)                            # staggered may legitimately hand one query back)
        if handed_back == 0:
            assert 0 < after.byte_rescored == after.byte_candidates          # every survivor of the filter is rescored
    # a query taken from the corpus (distance ~ 0 at rank 0), and a scaled query (normalised inside)
    _assert_same(fast, plain, rows[123], ks[0])
    _assert_same(fast, plain, 1e-3 * qs[0], ks[0])
    _assert_same(fast, plain, 1e4 * qs[1], ks[0])


def test_byte_prefilter_routes_what_it_does_not_cover(cs):
    rng = np.random.default_rng(5)
    rows = rng.standard_normal((20_000, 384)).astype(np.float32)
    fast, plain = _pair(cs, rows)
    q = rng.standard_normal(384).astype(np.float32)
    b0 = fast.device_stats().byte_searches
    _assert_same(fast, plain, q, 257)                                   # k above the int8 route's limit
    _assert_same(fast, plain, q, 1000)
    flt = cs.RowFilter.from_mask(np.arange(rows.shape[0]) % 3 == 0)     # k above the limit under a filter: fp32 filtered scan
    fi, fd = fast.search_ids(q, 300, flt)
    pi, pd = plain.search_ids(q, 300, flt)
    assert np.array_equal(fi, pi) and np.array_equal(fd.view(np.uint32), pd.view(np.uint32))
    assert fast.device_stats().byte_searches == b0
    # zero-norm query: the launch reports it, the fp32 kernel answers (every distance 0.0, ascending ids)
    z = np.zeros(384, dtype=np.float32)
    gi, gd = _assert_same(fast, plain, z, 10)
    assert gi.tolist() == list(range(10)) and not gd.any()
    s = fast.device_stats()
    assert s.byte_searches == b0 + 1 and s.byte_fallbacks >= 1
    # |q|^2 overflows fp32 (finite elements): the scan kernel's 1/sqrt(inf) = 0 scores every row 0.5 — same answer here
    f0 = s.byte_fallbacks
    big = (q * np.float32(1e19)).astype(np.float32)
    assert np.isfinite(big).all()
    _assert_same(fast, plain, big, 10)
    assert fast.device_stats().byte_fallbacks == f0 + 1
    _assert_same(fast, plain, (q * np.float32(1e-22)).astype(np.float32), 10)    # tiny but normalisable: the int8 route
    assert fast.device_stats().byte_fallbacks <= f0 + 2                          # (normally f0 + 1: answered by itself)


@pytest.mark.parametrize("n,d", [(200_000, 384), (60_000, 768), (30_000, 100)])
def test_byte_prefilter_under_filters_bit_identical(cs, oracle, n, d):
    """Round 2: filtered and tagged searches take the int8 route too (scan_i8_kernel FILT: only row groups the filter allows
    are streamed, only allowed rows publish bounds or become candidates). Same bar: ids AND distances bit-identical to the
    fp32 filtered scan — id bitmaps at several densities, tag predicates (language mask, file range, per-file bitmap),
    zero-norm rows inside and outside the filter, k up to 256, and an empty filter."""
    from codesearch_b200 import tags as T
    rng = np.random.default_rng(n * 3 + d)
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[[11, 5000, n - 3]] = 0.0
    ids = np.arange(n, dtype=np.uint32)
    tg = T.synth_tags(0, n)
    pair = []
    for on in (True, False):
        st = cs.VectorStore.new(None, d)
        st.append_rows(rows, ids, tg)
        if on:
            st.set_byte_prefilter(True)
        st.build_index()
        pair.append(st)
    fast, plain = pair
    qs = rng.standard_normal((4, d)).astype(np.float32)

    def same(a, b, what):
        assert np.array_equal(a[0], b[0]), (what, a[0][:6], b[0][:6])
        assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)), what

    b0 = fast.device_stats()
    n_calls = 0
    for dens in (0.9, 0.3, 0.02):
        mask = rng.random(n) < dens
        mask[11] = True; mask[5000] = False                      # one zero-norm row allowed, one masked
        flt = cs.RowFilter.from_mask(mask)
        for k in (10, 100, 256):
            got = fast.search_ids(qs[0], k, flt)
            same(got, plain.search_ids(qs[0], k, flt), ("bitmap", dens, k))
            n_calls += 1
            if k == 10:
                ri, rd, r64 = oracle.search(rows, qs[0], k + MARGIN, bitmap=flt.bitmap, n_bits=flt.n_bits)
                check_topk(got[0], got[1], ri, rd, r64, k)
    n_files = (n + 36) // 37
    fbm = np.zeros((n_files + 63) // 64, dtype=np.uint64)
    for f in range(0, n_files, 3):
        fbm[f >> 6] |= np.uint64(1) << np.uint64(f & 63)
    preds = [T.TagPredicate(lang_mask=0x0000FFFF), T.TagPredicate(file_lo=100, file_hi=n_files // 2),
             T.TagPredicate(lang_mask=0x7F0, file_lo=7, file_hi=n_files - 9, file_bitmap=fbm, n_file_bits=n_files),
             T.TagPredicate(file_bitmap=fbm, n_file_bits=n_files), T.TagPredicate(file_lo=3, file_hi=3)]
    for pi_, pred in enumerate(preds):
        for k in (10, 200):
            same(fast.search_tagged_ids(qs[1], k, pred), plain.search_tagged_ids(qs[1], k, pred), ("tags", pi_, k))
            n_calls += 1
    s1 = fast.device_stats()
    assert s1.byte_searches - b0.byte_searches == n_calls          # the int8 route really ran for every one of them ...
    assert s1.byte_fallbacks - b0.byte_fallbacks <= 3               # ... and (almost) always answered by itself
    ei, ed = fast.search_ids(qs[2], 10, cs.RowFilter(np.zeros(1, dtype=np.uint64), 0))   # an empty filter allows nothing
    assert len(ei) == 0
    z = np.zeros(d, np.float32)                                      # zero-norm query under a filter: handed to the fp32 kernel
    flt = cs.RowFilter.from_mask(np.arange(n) % 5 == 0)
    same(fast.search_ids(z, 10, flt), plain.search_ids(z, 10, flt), "zero query")


def test_byte_prefilter_adversarial_inputs(cs, oracle):
    """Inputs the filter cannot prune: near-duplicate clusters far inside the int8 bound, exact duplicates (ties by id),
    and a corpus sorted so that every row beats all earlier ones. Whatever the route, the answer is the exact one."""
    rng = np.random.default_rng(11)
    d = 384
    centres = rng.standard_normal((30, d)).astype(np.float32)
    rows = (np.repeat(centres, 700, axis=0) + 0.02 * rng.standard_normal((30 * 700, d))).astype(np.float32)
    rows[100] = rows[99]
    rows[5000] = rows[4999]
    rows = rows[rng.permutation(rows.shape[0])]
    fast, plain = _pair(cs, rows)
    for c in (0, 7, 29):
        q = (centres[c] + 0.01 * rng.standard_normal(d)).astype(np.float32)
        for k in (10, 100):
            gi, gd = _assert_same(fast, plain, q, k)
            ri, rd, r64 = oracle.np_search(rows, q, k + MARGIN)
            check_topk(gi, gd, ri, rd, r64, k)
    # every row identical: an n-way tie, ids decide. 21 000 candidates are all rescored and selected in the launch;
    # 90 000 overflow the candidate list -> the fp32 kernel answers
    for n_same, falls_back in ((21_000, False), (90_000, True)):
        same = np.tile(rng.standard_normal((1, d)).astype(np.float32), (n_same, 1))
        fast2, plain2 = _pair(cs, same)
        gi, _ = _assert_same(fast2, plain2, same[0], 10)
        assert gi.tolist() == list(range(10))
        if falls_back:     # (the smaller case normally answers by itself, but a staggered grid may legitimately hand it back)
            assert fast2.device_stats().byte_fallbacks >= 1, n_same
    # rows ordered from the farthest to the nearest: the running threshold never helps
    q = rng.standard_normal(d).astype(np.float32)
    base = rng.standard_normal((30_000, d)).astype(np.float32)
    cosv = (base @ q) / np.linalg.norm(base, axis=1)
    srt = base[np.argsort(cosv)]
    fast3, plain3 = _pair(cs, srt)
    _assert_same(fast3, plain3, q, 10)
    _assert_same(fast3, plain3, q, 256)


def test_byte_prefilter_zero_rows_updates_snapshot_toggle(cs, oracle, tmp_path):
    rng = np.random.default_rng(3)
    n, d = 12_000, 384
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[10] = 0.0                    # zero-norm rows: distance 0.0, injected by the tail
    rows[7777] = 0.0
    q = rng.standard_normal(d).astype(np.float32)
    db = str(tmp_path / "db")
    os.makedirs(db)
    fast = cs.VectorStore.new(db, d)
    fast.append_rows(rows, np.arange(n, dtype=np.uint32))
    fast.build_index()
    fast.set_byte_prefilter(True)                      # enabling after build creates the shadow now
    assert fast.device_stats().byte_shadow_bytes == (n - 2) * (384 + 4)
    plain = cs.VectorStore.new(None, d)
    plain.append_rows(rows, np.arange(n, dtype=np.uint32))
    plain.build_index()
    gi, gd = _assert_same(fast, plain, q, 10)
    assert gi[:2].tolist() == [10, 7777] and gd[0] == 0.0 and gd[1] == 0.0
    # delete the current best rows + append new ones, rebuild: the shadow follows
    kill = gi[2:6].tolist()
    extra = rng.standard_normal((500, d)).astype(np.float32)
    for st in (fast, plain):
        st.delete_chunks(kill)
        st.append_rows(extra, np.arange(n, n + 500, dtype=np.uint32))
        st.build_index()
    gi2, _ = _assert_same(fast, plain, q, 10)
    assert not set(kill) & set(gi2.tolist())
    assert fast.device_stats().byte_searches >= 2
    # reopen from the snapshot (hydrated in the constructor), then switch the prefilter on
    re = cs.VectorStore.new(db, d)
    re.set_byte_prefilter(True)
    b0 = re.device_stats().byte_searches
    if not re.is_indexed():
        pytest.skip("snapshot hydrate not available")
    _assert_same(re, plain, q, 10)
    assert re.device_stats().byte_shadow_bytes > 0
    assert re.device_stats().byte_searches == b0 + 1
    # toggle off: shadow dropped, same answers from the fp32 kernel
    fast.set_byte_prefilter(False)
    assert fast.device_stats().byte_shadow_bytes == 0
    b1 = fast.device_stats().byte_searches
    _assert_same(fast, plain, q, 10)
    assert fast.device_stats().byte_searches == b1
    # bf16 index: refused
    bf = cs.VectorStore.new(None, 384, dtype="bf16")
    with pytest.raises(cs.CsgpuError):
        bf.set_byte_prefilter(True)


def test_byte_prefilter_concurrent_searches(cs):
    """csgpu_search is re-entrant (&self in the reference): searches from several host threads each take their own
    context (scratch + status word) and all return the exact answer."""
    import threading
    rng = np.random.default_rng(21)
    rows = rng.standard_normal((50_000, 384)).astype(np.float32)
    fast, plain = _pair(cs, rows)
    qs = rng.standard_normal((16, 384)).astype(np.float32)
    want = [plain.search_ids(q, 10) for q in qs]
    errs = []

    def work(t):
        try:
            for rep in range(5):
                for j in range(t, len(qs), 4):
                    gi, gd = fast.search_ids(qs[j], 10)
                    assert np.array_equal(gi, want[j][0]) and np.array_equal(gd.view(np.uint32), want[j][1].view(np.uint32))
        except Exception as e:   # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs


@pytest.mark.parametrize("which", [(0, 1, 2), (1,)])
def test_byte_prefilter_device_entry_points_and_fused_exchange(cs, oracle, which):
    """`which` = the ranks with the byte prefilter on. Mixed ranks launch DIFFERENT kernels for the same query: a rank
    already spinning in the exchange must not stall a peer's first (lazily loaded) launch — every kernel an exchange
    search can use is loaded when the exchange / the prefilter is set up (preload_exchange_kernels).
    The device-resident entry points (rank-per-GPU sharding) take the int8 route too: the fp32 scan is enqueued
    behind it as a launch that only runs if the int8 kernel raised its device-side status word, then the exchange runs
    as a launch of its own. Three "ranks" on one GPU (three streams), compared with the unsharded plain index."""
    import ctypes
    import torch
    from codesearch_b200 import _lib
    from codesearch_b200.sharded import decode_keys
    lib = _lib.load()
    rng = np.random.default_rng(41)
    n, d = 60_000, 384
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[5] = 0.0
    bounds = [0, 15_000, 38_000, n]
    W = 3
    stores = []
    for a, b in zip(bounds, bounds[1:]):
        st = cs.VectorStore.new(None, d)
        st.append_rows(rows[a:b], np.arange(a, b, dtype=np.uint32))
        if len(stores) in which:
            st.set_byte_prefilter(True)
        st.build_index()
        stores.append(st)
    whole = cs.VectorStore.new(None, d)
    whole.append_rows(rows, np.arange(n, dtype=np.uint32))
    whole.build_index()
    qs = rng.standard_normal((6, d)).astype(np.float32)
    qs[4] = 0.0                                                  # zero-norm query: the conditional fp32 scan answers
    # local keys of one rank through csgpu_search_keys_device == host search of the same shard, bit for bit
    for qi, k in enumerate([10, 100, 128, 200, 10]):
        qd = torch.from_numpy(qs[qi]).cuda()
        out = torch.empty(k, dtype=torch.int64, device="cuda")
        _lib.check(lib.csgpu_search_keys_device(stores[1].handle, qd.data_ptr(), k, out.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        ids, dist = decode_keys(out.cpu().numpy())
        stores[1].set_byte_prefilter(False)
        hi, hd = stores[1].search_ids(qs[qi], k)
        stores[1].set_byte_prefilter(True)
        assert np.array_equal(ids, hi) and np.array_equal(dist.view(np.uint32), hd.view(np.uint32)), (qi, k)
    assert stores[1].device_stats().byte_searches >= 4
    # fused exchange
    for r, st in enumerate(stores):
        h = (ctypes.c_ubyte * 64)()
        _lib.check(lib.csgpu_exchange_create(st.handle, W, r, h))
    peers = (ctypes.c_void_p * W)(*[st.handle for st in stores])
    for st in stores:
        _lib.check(lib.csgpu_exchange_connect_local(st.handle, peers))
    streams = [torch.cuda.Stream() for _ in range(W)]
    for qi, k in enumerate([10, 10, 32, 100, 10, 1000]):
        qd = torch.from_numpy(qs[qi]).cuda()
        outs = [torch.empty(k, dtype=torch.int64, device="cuda") for _ in range(W)]
        torch.cuda.synchronize()
        for r, st in enumerate(stores):
            _lib.check(lib.csgpu_search_keys_exchange_device(st.handle, qd.data_ptr(), k, outs[r].data_ptr(),
                                                             streams[r].cuda_stream))
        torch.cuda.synchronize()
        gi, gd = whole.search_ids(qs[qi], k)
        for r in range(W):
            ids, dist = decode_keys(outs[r].cpu().numpy())
            assert np.array_equal(ids, gi) and np.array_equal(dist.view(np.uint32), gd.view(np.uint32)), (qi, k, r)
        if qi != 4:
            oi, od, o64 = oracle.np_search(rows, qs[qi], k + MARGIN)
            check_topk(gi, gd, oi, od, o64, min(k, n))
    for st in stores:
        t = ctypes.c_uint32(7)
        _lib.check(lib.csgpu_exchange_status(st.handle, ctypes.byref(t)))
        assert t.value == 0
        lib.csgpu_exchange_destroy(st.handle)


def test_byte_prefilter_small_groups_prefer_the_int8_kernel(cs):
    """Host routing: with the prefilter on, up to 4 queries of a batch (or of a coalesced group) run one after the other
    through the int8 kernel — faster than one fp32 multi-query pass — and larger groups keep the multi-query scan."""
    rng = np.random.default_rng(9)
    rows = rng.standard_normal((30_000, 384)).astype(np.float32)
    fast, plain = _pair(cs, rows)
    qs = rng.standard_normal((8, 384)).astype(np.float32)
    for b, via_int8 in ((3, True), (4, True), (8, False)):
        s0 = fast.device_stats().byte_searches
        a = fast.search_batch_ids(qs[:b], 10)
        c = plain.search_batch_ids(qs[:b], 10)
        assert np.array_equal(a[0], c[0]) and np.array_equal(a[1].view(np.uint32), c[1].view(np.uint32)) and np.array_equal(a[2], c[2])
        assert fast.device_stats().byte_searches - s0 == (b if via_int8 else 0), b
