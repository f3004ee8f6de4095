#!/usr/bin/env python
"""N3 (query variants, src/search/mod.rs:508-590): 9 variants of one query over 10M x 384 through csgpu_search_variants,
multi-query-scan + device dedup vs one tensor-prefilter batch + host dedup; results must be identical."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import codesearch_b200 as cs
from codesearch_b200 import _lib
n, d = 10_000_000, 384
st = cs.VectorStore.new(None, d); st.reserve(n); st.append_synthetic(1234, 0, n); st.build_index()
qs = np.empty((9, d), np.float32)
_lib.check(_lib.load().csgpu_synth_rows_host(st.handle, 4321, 0, 9, qs.ctypes.data_as(_lib._f32p)))
def t(k, reps=10):
    st.search_variants_ids(qs, k); t0 = time.perf_counter()
    for _ in range(reps): r = st.search_variants_ids(qs, k)
    return (time.perf_counter() - t0) / reps * 1e3, r
for k in (10, 200):
    a, ra = t(k)
    st.set_tensor_prefilter(True)
    b, rb = t(k)
    st.set_tensor_prefilter(False)
    print(f"9 variants x top-{k}, 10M x 384: scan route {a:.3f} ms, tensor-prefilter route {b:.3f} ms, identical: {np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1].view(np.uint32), rb[1].view(np.uint32))}", flush=True)
