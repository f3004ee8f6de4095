#!/usr/bin/env python
"""Hybrid search under a language / path filter (the caller shape of BASELINE configs[4]): 9 query variants x limit 200 under
a row-tag predicate as ONE csgpu_search_variants_tagged call, beside 9 csgpu_search_tagged calls + host dedup.
  python tools/bench_variants_tagged.py [rows dim ...]   (default: 100000 384  5000000 768)"""
import os, sys, time, json
sys.path.insert(0, os.getcwd())
import numpy as np
import codesearch_b200 as cs
from codesearch_b200 import _lib
from codesearch_b200.tags import TagPredicate
from oracle import oracle as O   # host-side dedup restatement only (the checker, never timed as the product)
args = [int(x) for x in sys.argv[1:]] or [100_000, 384, 5_000_000, 768]
for n, d in zip(args[0::2], args[1::2]):
    st = cs.VectorStore.new(None, d); st.reserve(n); st.append_synthetic(1234, 0, n, 0, tagged=True); st.build_index()
    qs = np.empty((9, d), np.float32)
    _lib.check(_lib.load().csgpu_synth_rows_host(st.handle, 4321, 0, 9, qs.ctypes.data_as(_lib._f32p)))
    n_files = n // 37 + 1
    for name, p in (("density 1.0", TagPredicate()), ("density ~0.25 (6 of 23 languages)", TagPredicate(lang_mask=0x3F)),
                    ("density ~0.01 (a path prefix: 1 % of the files)", TagPredicate(file_lo=n_files // 2, file_hi=n_files // 2 + n_files // 100))):
        for k in (10, 200):
            for _ in range(3): st.search_variants_tagged_ids(qs, k, p)
            t0 = time.perf_counter()
            for _ in range(20): g = st.search_variants_tagged_ids(qs, k, p)
            one = (time.perf_counter() - t0) / 20 * 1e3
            t0 = time.perf_counter()
            for _ in range(5): w = O.dedup_variants([st.search_tagged_ids(q, k, p) for q in qs], k)
            many = (time.perf_counter() - t0) / 5 * 1e3
            same = bool(np.array_equal(g[0], w[0]) and np.array_equal(g[1].view(np.uint32), w[1].view(np.uint32)))
            print(json.dumps({"rows": n, "dim": d, "filter": name, "k": k, "one_call_ms": round(one, 3),
                              "nine_tagged_searches_plus_host_dedup_ms": round(many, 3), "identical": same}), flush=True)
    st.close()
