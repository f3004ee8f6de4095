"""Row-tag filter columns: csgpu_append_tagged / csgpu_search_tagged / csgpu_search_tagged_keys_device.

The predicate (language mask, file range, per-file bitmap) is evaluated on the device from the packed tag column
before a row is read. Parity: identical ids/distances to (i) the oracle restricted to the rows the oracle's own
predicate restatement allows and (ii) csgpu_search_filtered with the equivalent id bitmap (bit-identical).
GPU box only.
"""
import ctypes
import os

import numpy as np
import pytest

from parity import check_topk

pytestmark = pytest.mark.gpu
os.environ.setdefault("CSGPU_I8_MIN_ROWS", "4096")   # read once by the library: lets small shards take the int8 route

MARGIN = 8


@pytest.fixture(scope="module")
def cs():
    import codesearch_b200 as m
    m.load_library()
    return m


def _bitmap(mask):
    n = mask.size
    bm = np.zeros((n + 63) // 64, dtype=np.uint64)
    for f in np.nonzero(mask)[0]:
        bm[f >> 6] |= np.uint64(1) << np.uint64(f & 63)
    return bm


def _cases(rng, n_files):
    from codesearch_b200.tags import TagPredicate
    fmask = rng.random(n_files) < 0.2
    bm = _bitmap(fmask)
    return [
        TagPredicate(),
        TagPredicate(lang_mask=(1 << 0) | (1 << 1) | (1 << 13)),
        TagPredicate(file_lo=n_files // 4, file_hi=n_files // 2),
        TagPredicate(lang_mask=0x7FFFFF & ~(1 << 3), file_bitmap=bm, n_file_bits=n_files),
        TagPredicate(file_bitmap=bm, n_file_bits=n_files // 3),
        TagPredicate(lang_mask=1 << 7, file_lo=5, file_hi=5),
        TagPredicate(lang_mask=0),
    ]


@pytest.mark.parametrize("d", [384, 768, 100])
def test_tagged_parity(cs, oracle, d):
    rng = np.random.default_rng(100 + d)
    n = 30000
    rows = rng.standard_normal((n, d)).astype(np.float32)
    ids = rng.permutation(2 * n)[:n].astype(np.uint32)
    tags = oracle.synth_tags(0, n)
    tags[rng.random(n) < 0.02] = 0xFFFFFFFF                       # some untagged rows
    rows[[7, 700, 7000]] = 0.0                                     # zero-norm rows keep their tags on the side list
    st = cs.VectorStore.new(None, d)
    st.append_rows(rows, ids, tags)
    st.build_index()
    assert np.array_equal(st.get_tags(ids[:500]), tags[:500])
    assert np.array_equal(st.get_tags(ids[[7, 700, 7000]]), tags[[7, 700, 7000]])
    n_files = n // 37 + 1
    for ci, p in enumerate(_cases(rng, n_files)):
        row_ok = oracle.tag_predicate_mask(tags, p.lang_mask, p.file_lo, p.file_hi, p.file_bitmap, p.n_file_bits)
        allowed = np.zeros(2 * n, dtype=bool)
        allowed[ids[row_ok]] = True
        flt = cs.RowFilter.from_mask(allowed)
        for k in (10, 200):
            q = rng.standard_normal(d).astype(np.float32)
            gi, gd = st.search_tagged_ids(q, k, p)
            fi, fd = st.search_ids(q, k, flt)
            assert np.array_equal(gi, fi) and np.array_equal(gd, fd), (ci, k)   # same kernel arithmetic, other filter source
            k_eff = min(k, int(row_ok.sum()))
            assert len(gi) == k_eff
            oi, od, o64 = oracle.np_search(rows, q, k + MARGIN, ids=ids, allowed=allowed)
            check_topk(gi, gd, oi, od, o64, k_eff)


@pytest.mark.parametrize("n,d,devices", [(30_000, 384, None), (30_000, 100, None), (24_000, 384, [0, 0]), (700_000, 384, None)])
def test_variants_under_a_tag_predicate(cs, oracle, n, d, devices):
    """csgpu_search_variants_tagged = the reference's hybrid search with a language / path filter as ONE call: b variants,
    each searched over the rows that pass the predicate, deduplicated by chunk id. Must equal the dedup of the b
    csgpu_search_tagged lists bit for bit — on the multi-query route (small corpora: the predicate is applied to the results
    that beat a threshold), on the per-variant filtered-scan route (700k x 384 = 1.08 GB > the 768 MB crossover; also
    dim % 128 != 0), and on a two-shard index."""
    rng = np.random.default_rng(n + d)
    tags = oracle.synth_tags(0, n)
    st = cs.VectorStore.new(None, d, devices=devices)
    if n >= 500_000:
        st.append_synthetic(55, 0, n, 0, tagged=True)
    else:
        rows = rng.standard_normal((n, d)).astype(np.float32)
        rows[[7, 700, 7000]] = 0.0                                 # zero-norm rows keep their tags on the side list
        tags[rng.random(n) < 0.02] = 0xFFFFFFFF                    # some untagged rows
        st.append_rows(rows, np.arange(n, dtype=np.uint32), tags)
    st.build_index()
    n_files = n // 37 + 1
    base = rng.standard_normal(d).astype(np.float32)
    qs = np.stack([base] + [base + np.float32(0.3) * rng.standard_normal(d).astype(np.float32) for _ in range(15)])
    qs[4] = 0.0                                                    # a zero-norm variant: distance 0.0 to every allowed row
    lib = cs._lib.load()
    for ci, p in enumerate(_cases(rng, n_files)):
        for b, k in ((9, 200), (2, 10), (16, 32), (5, 33), (1, 10)):
            l0 = lib.csgpu_kernel_launches()
            gi, gd = st.search_variants_tagged_ids(qs[:b], k, p)
            launches = lib.csgpu_kernel_launches() - l0
            lists = [st.search_tagged_ids(q, k, p) for q in qs[:b]]
            wi, wd = oracle.dedup_variants(lists, k)
            assert np.array_equal(gi, wi) and np.array_equal(gd.view(np.uint32), wd.view(np.uint32)), (ci, b, k)
            assert len(set(gi.tolist())) == len(gi)
            if devices is None and d % 128 == 0 and n < 500_000 and b >= 2:
                assert launches == 2, (launches, b, k)             # one multi-query pass + the dedup kernel
    # no predicate restriction at all == the unfiltered variants call
    from codesearch_b200.tags import TagPredicate
    if n < 500_000:
        ai, ad = st.search_variants_tagged_ids(qs[:9], 50, TagPredicate())
        ui, ud = st.search_variants_ids(qs[:9], 50)
        untagged_hit = (st.get_tags(ui) == 0xFFFFFFFF).any()
        if not untagged_hit:                                       # (an untagged row passes no language mask)
            assert np.array_equal(ai, ui) and np.array_equal(ad.view(np.uint32), ud.view(np.uint32))
    with pytest.raises(cs.CsgpuError):
        st.search_variants_tagged_ids(qs[:9, : d - 1], 10, TagPredicate())


def test_tags_follow_rows_through_delete_and_rebuild(cs, oracle):
    rng = np.random.default_rng(77)
    n, d = 6000, 384
    rows = rng.standard_normal((n, d)).astype(np.float32)
    ids = np.arange(n, dtype=np.uint32)
    tags = oracle.synth_tags(0, n)
    st = cs.VectorStore.new(None, d)
    st.append_rows(rows[:4000], ids[:4000], tags[:4000])
    st.build_index()
    dead = rng.permutation(4000)[:1500].astype(np.uint32)
    st.delete_chunks(dead)
    st.append_rows(rows[4000:], ids[4000:], tags[4000:])
    st.append_rows(rows[:3] * 2.0, ids[:3], np.array([5, 6, 7], np.uint32))   # replace: new tag wins
    st.build_index()
    live = np.ones(n, dtype=bool)
    live[dead] = False
    live[:3] = True
    want = tags.copy()
    want[:3] = [5, 6, 7]
    got = st.get_tags(ids)
    assert np.array_equal(got[live], want[live])
    assert (got[~live] == 0xFFFFFFFF).all()
    from codesearch_b200.tags import TagPredicate
    p = TagPredicate(lang_mask=(1 << 2) | (1 << 9) | (1 << 0))
    row_ok = oracle.tag_predicate_mask(want, p.lang_mask) & live
    q = rng.standard_normal(d).astype(np.float32)
    gi, gd = st.search_tagged_ids(q, 50, p)
    allowed = np.zeros(n, dtype=bool)
    allowed[row_ok] = True
    oi, od, o64 = oracle.np_search(rows[live], q, 50 + MARGIN, ids=ids[live], allowed=allowed)
    check_topk(gi, gd, oi, od, o64, 50)


def test_synthetic_tags_and_snapshot(cs, oracle, tmp_path):
    n, d = 20000, 384
    st = cs.VectorStore.new(str(tmp_path / "db"), d)
    st.append_synthetic(1234, 1000, n, 0, tagged=True)
    st.build_index()                                               # writes the snapshot, tags.u32 included
    ids = np.arange(1000, 1000 + n, dtype=np.uint32)
    want = oracle.synth_tags(1000, n)
    assert np.array_equal(st.get_tags(ids), want)
    from codesearch_b200.tags import TagPredicate
    p = TagPredicate(lang_mask=(1 << 4) | (1 << 11), file_lo=100, file_hi=400)
    q = oracle.synth_rows(4321, 3, 1, d)[0]
    a = st.search_tagged_ids(q, 100, p)
    st2 = cs.VectorStore.open_readonly(str(tmp_path / "db"), d)
    assert np.array_equal(st2.get_tags(ids), want)
    b = st2.search_tagged_ids(q, 100, p)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    rows = oracle.synth_rows(1234, 1000, n, d)
    ok = oracle.tag_predicate_mask(want, p.lang_mask, p.file_lo, p.file_hi)
    allowed = np.zeros(1000 + n, dtype=bool)
    allowed[ids[ok]] = True
    oi, od, o64 = oracle.np_search(rows, q, 100 + MARGIN, ids=ids, allowed=allowed)
    check_topk(a[0], a[1], oi, od, o64, min(100, int(ok.sum())))


def test_search_tagged_by_path_and_language(cs):
    rng = np.random.default_rng(3)
    d = 64
    paths = ["/r/src/auth.rs", "/r/src/math.py", "/r/docs/readme.md", "/r/src/db/store.rs", "/r/Makefile"]
    chunks = []
    for i in range(200):
        p = paths[i % len(paths)]
        chunks.append(cs.EmbeddedChunk(cs.Chunk(f"chunk {i}", i, i + 1, "Function", p), rng.standard_normal(d).astype(np.float32)))
    st = cs.VectorStore.new(None, d)
    st.insert_chunks(chunks)
    st.build_index()
    q = rng.standard_normal(d).astype(np.float32)
    full = st.search(q, 200)
    r = st.search_tagged(q, 10, languages=["Rust"])
    assert len(r) == 10 and all(x.path.endswith(".rs") for x in r)
    assert [x.id for x in r] == [x.id for x in full if x.path.endswith(".rs")][:10]      # == host post-filter order
    r = st.search_tagged(q, 10, path_prefix="src/", project_root="/r")                     # src/search/mod.rs:727-737
    assert [x.id for x in r] == [x.id for x in full if x.path.startswith("/r/src/")][:10]
    r = st.search_tagged(q, 10, path_contains="db")                                        # src/server/mod.rs:553-559
    assert [x.id for x in r] == [x.id for x in full if "db" in x.path][:10]
    r = st.search_tagged(q, 500, languages=["Shell", "Markdown"])
    assert len(r) == 80
    assert st.search_tagged(q, 10, languages=["Go"]) == []
    # hybrid search under the same filters: query variants -> one list, every hit inside the filter, `limit` of them survive
    qs = np.stack([q] + [q + np.float32(0.3) * rng.standard_normal(d).astype(np.float32) for _ in range(8)])
    r = st.search_variants_tagged(qs, 10, languages=["Rust"])
    assert len(r) == 10 and all(x.path.endswith(".rs") for x in r) and len({x.id for x in r}) == 10
    per_variant = [st.search_tagged(v, 10, languages=["Rust"]) for v in qs]
    best = {}
    for lst in per_variant:
        for x in lst:
            if x.id not in best or x.distance < best[x.id]:
                best[x.id] = x.distance
    want = sorted(best.items(), key=lambda t: (t[1], t[0]))[:10]
    assert [(x.id, x.distance) for x in r] == want
    assert st.search_variants_tagged(qs, 10, languages=["Go"]) == []


@pytest.mark.parametrize("byte_prefilter", [False, True])
def test_tagged_device_entry_with_fused_exchange(cs, oracle, byte_prefilter):
    """BASELINE configs[4] shape on one GPU: 3 'ranks', predicate filter + fused exchange in the same kernel. With the byte
    prefilter on (round 2): int8 FILT kernel -> conditional fp32 filtered scan (no-op) -> exchange_keys_kernel per rank."""
    import torch
    from codesearch_b200 import _lib
    from codesearch_b200.sharded import decode_keys
    lib = _lib.load()
    rng = np.random.default_rng(41)
    n, d, W = 30000, 768, 3
    rows = rng.standard_normal((n, d)).astype(np.float32)
    ids = np.arange(n, dtype=np.uint32)
    tags = oracle.synth_tags(0, n)
    bounds = [0, 8000, 21000, n]
    stores = []
    for a, b in zip(bounds, bounds[1:]):
        st = cs.VectorStore.new(None, d)
        st.append_rows(rows[a:b], ids[a:b], tags[a:b])
        if byte_prefilter:
            st.set_byte_prefilter(True)
        st.build_index()
        stores.append(st)
    for r, st in enumerate(stores):
        h = (ctypes.c_ubyte * 64)()
        _lib.check(lib.csgpu_exchange_create(st.handle, W, r, h))
    peers = (ctypes.c_void_p * W)(*[st.handle for st in stores])
    for st in stores:
        _lib.check(lib.csgpu_exchange_connect_local(st.handle, peers))
    searches0 = sum(st.device_stats().byte_searches for st in stores)
    n_files = n // 37 + 1
    fmask = rng.random(n_files) < 0.25
    bm_dev = torch.from_numpy(_bitmap(fmask).view(np.int64)).cuda()
    streams = [torch.cuda.Stream() for _ in range(W)]
    for qi, (k, use_bm, lang_mask) in enumerate([(200, True, 0xFFFFFFFF), (10, False, (1 << 1) | (1 << 6)), (200, True, 0x3FF)]):
        pred = _lib.Predicate(lang_mask, 0, 0xFFFFFFFF, 0, bm_dev.data_ptr() if use_bm else None, n_files if use_bm else 0)
        q = rng.standard_normal(d).astype(np.float32)
        qd = torch.from_numpy(q).cuda()
        outs = [torch.empty(k, dtype=torch.int64, device="cuda") for _ in range(W)]
        torch.cuda.synchronize()
        for r, st in enumerate(stores):
            _lib.check(lib.csgpu_search_tagged_keys_device(st.handle, qd.data_ptr(), k, ctypes.byref(pred), 1,
                                                           outs[r].data_ptr(), streams[r].cuda_stream))
        torch.cuda.synchronize()
        ok = oracle.tag_predicate_mask(tags, lang_mask, 0, 0xFFFFFFFF, _bitmap(fmask) if use_bm else None, n_files if use_bm else 0)
        allowed = np.zeros(n, dtype=bool)
        allowed[ok] = True
        oi, od, o64 = oracle.np_search(rows, q, k + MARGIN, ids=ids, allowed=allowed)
        first = decode_keys(outs[0].cpu().numpy())
        for r in range(1, W):
            ri, rd = decode_keys(outs[r].cpu().numpy())
            assert np.array_equal(ri, first[0]) and np.array_equal(rd, first[1])
        check_topk(first[0], first[1], oi, od, o64, min(k, int(ok.sum())))
    assert sum(st.device_stats().byte_searches for st in stores) - searches0 == (9 if byte_prefilter else 0)
    for st in stores:
        lib.csgpu_exchange_destroy(st.handle)


def test_filtered_scan_work_counter_split_matches_fixed_stride(cs):
    """The work-counter row split of the filtered scan (BlockCursor, scan.cuh) is off by default — measured not faster,
    profiles/r02_static_vs_dynamic_filtered_multi.txt — but it ships behind CSGPU_SCAN_DYNAMIC_ALL=1, so it is tested: a
    child process with the switch on must return, bit for bit, what this process (fixed stride) returns, for the id bitmap
    and the tag predicate, with the per-warp selector (k <= 32) and the CTA-shared one (k > 32), and for the 8-query pass."""
    import json, os, subprocess, sys
    code = r'''
import json, sys, numpy as np
sys.path.insert(0, %r)
import codesearch_b200 as cs
from codesearch_b200.tags import TagPredicate
st = cs.VectorStore.new(None, 128)
st.append_synthetic(77, 0, 90_000, 0, tagged=True)
st.build_index()
rng = np.random.default_rng(5)
qs = rng.standard_normal((8, 128)).astype(np.float32)
flt = cs.RowFilter.from_mask(np.arange(90_000) %% 3 != 1)
pred = TagPredicate(lang_mask=0x0FFF, file_lo=100, file_hi=2000)
out = []
for k in (10, 100):
    for r in (st.search_ids(qs[0], k, flt), st.search_tagged_ids(qs[1], k, pred)):
        out.append([r[0].tolist(), r[1].view(np.uint32).tolist()])
    b = st.search_batch_ids(qs, k)
    out.append([b[0].tolist(), b[1].view(np.uint32).tolist()])
print("RESULT" + json.dumps(out))
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = []
    for dyn in ("0", "1"):
        env = dict(os.environ, CSGPU_SCAN_DYNAMIC_ALL=dyn)
        p = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr[-2000:]
        res.append(json.loads(p.stdout.split("RESULT", 1)[1]))
    assert res[0] == res[1]
