#!/usr/bin/env python
"""Small end-to-end exercise of every kernel family, meant to run under compute-sanitizer
(tools/sanitize.sh: memcheck, racecheck, synccheck). Sizes are tiny: the sanitizer slows kernels 10-100x."""
import os, sys
os.environ.setdefault("CSGPU_I8_MIN_ROWS", "1024")          # let the tiny corpora below take the byte-prefilter route
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import codesearch_b200 as cs
from codesearch_b200.tags import TagPredicate, synth_tags

rng = np.random.default_rng(1)
n, d = 6000, 384
rows = rng.standard_normal((n, d)).astype(np.float32)
rows[17] = 0.0
ids = np.arange(n, dtype=np.uint32)
tags = synth_tags(0, n)

st = cs.VectorStore.new(None, d)
st.append_rows(rows, ids, tags)
st.build_index()
q = rng.standard_normal(d).astype(np.float32)
for k in (10, 100, 1000):                                   # single-query scan: register selector, CTA buffer, column merge
    i, _ = st.search_ids(q, k)
    assert len(i) == k
st.search_ids(q, 50, cs.RowFilter.from_mask(rng.random(n) < 0.3))                      # id-bitmap filter
st.search_tagged_ids(q, 50, TagPredicate(lang_mask=0x3F, file_lo=3, file_hi=120))     # row-tag predicate
qs = rng.standard_normal((70, d)).astype(np.float32)
for kk in (10, 100):                                        # variants under a tag predicate: multi-query kernels with the predicate
    vt = st.search_variants_tagged_ids(qs[:9], kk, TagPredicate(lang_mask=0x3F, file_lo=3, file_hi=120))
    assert len(set(vt[0].tolist())) == len(vt[0])
st.search_batch_ids(qs[:8], 10)                             # multi-query scan, per-warp lists
st.search_batch_ids(qs[:8], 100)                            # multi-query scan, CTA buffers
st.search_variants_ids(qs[:5], 20)                          # device-side dedup of variants
st.search_batch_ids(qs[:12], 10)                            # 9..16 queries in one pass (two query groups), per-warp lists
st.search_batch_ids(qs[:16], 200)                           # ... and CTA buffers
st.search_variants_ids(qs[:9], 200)                         # the reference's hybrid shape: 9 variants x limit 200
os.environ["CSGPU_BATCH_SIMT"] = "1"
st.search_batch_ids(qs[:40], 20)                            # fp32 SIMT GEMM, 64-query tile (setmaxnreg warp specialisation)
st.search_batch_ids(qs, 20)                                 # fp32 SIMT GEMM, 128-query tile + select
assert st.device_stats().batch_route == 1
os.environ["CSGPU_BATCH_SIMT"] = "0"
os.environ["CSGPU_GEMM_MIN_BATCH"] = "2"
for nb, kk in ((70, 20), (5, 10), (40, 200)):               # tcgen05 kind::tf32 filter off the fp32 rows + rescoring select
    t = st.search_batch_ids(qs[:nb], kk)                    # (70: 128 query rows per chunk; 5 / 40: the small TMA box + zeroed rows)
    assert st.device_stats().batch_route == 3
    for j in (0, nb - 1):
        gi, gd = st.search_ids(qs[j], kk)
        assert np.array_equal(t[0][j], gi) and np.array_equal(t[1][j], gd)
v1 = st.search_variants_ids(qs[:9], 50)                     # variants as one tf32 batch + host dedup ...
os.environ["CSGPU_GEMM_MIN_BATCH"] = "100000"
v2 = st.search_variants_ids(qs[:9], 50)                     # ... equal the multi-query scan + device dedup
assert np.array_equal(v1[0], v2[0]) and np.array_equal(v1[1], v2[1])
os.environ.pop("CSGPU_GEMM_MIN_BATCH")
st.set_tensor_prefilter(True)
a = st.search_batch_ids(qs, 20)                             # tcgen05 filter + rescoring select
for j in (0, 69):
    gi, gd = st.search_ids(qs[j], 20)
    assert np.array_equal(a[0][j], gi) and np.array_equal(a[1][j], gd)
st.delete_chunks(ids[::7])
st.build_index()                                            # compaction + shadow rebuild
st.search_batch_ids(qs[:3], 5)
st.set_byte_prefilter(True)                                 # int8 shadow build + scan_i8_kernel (helper warp, CTA-end rescoring, tail)
for k in (10, 100, 256):
    gi, gd = st.search_ids(q, k)
    st.set_byte_prefilter(False)
    hi, hd = st.search_ids(q, k)
    st.set_byte_prefilter(True)
    assert np.array_equal(gi, hi) and np.array_equal(gd, hd)
flt8 = cs.RowFilter.from_mask(rng.random(st.device_stats().live_rows + 2000) < 0.4)     # scan_i8_kernel FILT: id bitmap ...
for k in (10, 100):
    gi, gd = st.search_ids(q, k, flt8)
    ti, td = st.search_tagged_ids(q, k, TagPredicate(lang_mask=0x3FF, file_lo=2, file_hi=140))   # ... and tag predicate
    st.set_byte_prefilter(False)
    hi, hd = st.search_ids(q, k, flt8)
    ui, ud = st.search_tagged_ids(q, k, TagPredicate(lang_mask=0x3FF, file_lo=2, file_hi=140))
    st.set_byte_prefilter(True)
    assert np.array_equal(gi, hi) and np.array_equal(gd, hd) and np.array_equal(ti, ui) and np.array_equal(td, ud)
st.search_ids(np.zeros(d, np.float32), 10)                  # zero-norm query: status word -> fp32 kernel
assert st.device_stats().byte_searches >= 4
st.set_byte_prefilter(False)
bf = cs.VectorStore.new(None, d, dtype="bf16")
bf.append_rows(rows, ids)
bf.build_index()
bf.search_batch_ids(qs, 20)                                 # bf16 index on tcgen05
zq = qs[:3].copy(); zq[1] = 0.0
zi, zd, zn = bf.search_batch_ids(zq, 7)                     # zero-norm query on the bf16 index: zero_query_ids_kernel
assert np.array_equal(zi[1], np.arange(7, dtype=np.uint32)) and not zd[1].any()
# in-process multi-device index, both shards on this device: N scan launches with the gather exchange fused into their tails
two = cs.VectorStore.new(None, d, devices=[0, 0])
two.append_rows(rows, ids, tags)
two.build_index()
for k in (10, 100):
    a2 = two.search_ids(q, k)
    assert len(a2[0]) == k and len(set(a2[0].tolist())) == k
two.search_ids(q, 50, cs.RowFilter.from_mask(rng.random(n) < 0.3))
two.search_tagged_ids(q, 50, TagPredicate(lang_mask=0x3F, file_lo=3, file_hi=120))
# fused cross-GPU exchange, three "ranks" on one device (three streams), with and without the tag predicate
import ctypes
import torch
from codesearch_b200 import _lib
from codesearch_b200.sharded import decode_keys
lib = _lib.load()
W, bounds = 3, [0, 1500, 4000, n]
stores = []
for a0, b0 in zip(bounds, bounds[1:]):
    s_ = cs.VectorStore.new(None, d)
    s_.append_rows(rows[a0:b0], ids[a0:b0], tags[a0:b0])
    s_.build_index()
    stores.append(s_)
stores[1].set_byte_prefilter(True)                          # one rank on the int8 route (+ conditional scan + exchange_keys_kernel)
for r, s_ in enumerate(stores):
    h = (ctypes.c_ubyte * 64)()
    _lib.check(lib.csgpu_exchange_create(s_.handle, W, r, h))
peers = (ctypes.c_void_p * W)(*[s_.handle for s_ in stores])
for s_ in stores:
    _lib.check(lib.csgpu_exchange_connect_local(s_.handle, peers))
streams = [torch.cuda.Stream() for _ in range(W)]
qd = torch.from_numpy(q).cuda()
pred = _lib.Predicate(0xFFFF, 0, 0xFFFFFFFF, 0, None, 0)
for k, tagged in ((10, False), (100, False), (50, True)):
    outs = [torch.empty(k, dtype=torch.int64, device="cuda") for _ in range(W)]
    torch.cuda.synchronize()
    for r, s_ in enumerate(stores):
        if tagged:
            _lib.check(lib.csgpu_search_tagged_keys_device(s_.handle, qd.data_ptr(), k, ctypes.byref(pred), 1, outs[r].data_ptr(), streams[r].cuda_stream))
        else:
            _lib.check(lib.csgpu_search_keys_exchange_device(s_.handle, qd.data_ptr(), k, outs[r].data_ptr(), streams[r].cuda_stream))
    torch.cuda.synchronize()
    first = decode_keys(outs[0].cpu().numpy())
    for r in range(1, W):
        other = decode_keys(outs[r].cpu().numpy())
        assert np.array_equal(first[0], other[0])
print("sanitize driver ok")
