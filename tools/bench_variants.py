#!/usr/bin/env python
"""N3 (query variants, src/search/mod.rs:508-590): 9 variants of one query through csgpu_search_variants at the reference's
own scale (100k rows) and at 10M x 384, three routes: multi-query scan + device dedup (CSGPU_GEMM_MIN_BATCH=100000), the
default routing (cost model: tf32 tensor-core batch + host dedup where it is faster), and the opt-in bf16 tensor prefilter.
Results must be identical bit for bit.   python tools/bench_variants.py [rows ...]"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import codesearch_b200 as cs
from codesearch_b200 import _lib
d = 384
for n in [int(x) for x in sys.argv[1:]] or [100_000, 10_000_000]:
    st = cs.VectorStore.new(None, d); st.reserve(n); st.append_synthetic(1234, 0, n); st.build_index()
    qs = np.empty((9, d), np.float32)
    _lib.check(_lib.load().csgpu_synth_rows_host(st.handle, 4321, 0, 9, qs.ctypes.data_as(_lib._f32p)))
    def t(k, reps=20):
        st.search_variants_ids(qs, k); st.search_variants_ids(qs, k); t0 = time.perf_counter()
        for _ in range(reps): r = st.search_variants_ids(qs, k)
        return (time.perf_counter() - t0) / reps * 1e3, r
    same = lambda x, y: bool(np.array_equal(x[0], y[0]) and np.array_equal(x[1].view(np.uint32), y[1].view(np.uint32)))
    for k in (10, 50, 200):
        os.environ["CSGPU_GEMM_MIN_BATCH"] = "100000"
        a, ra = t(k)
        os.environ.pop("CSGPU_GEMM_MIN_BATCH")
        c, rc = t(k)
        route = st.device_stats().batch_route
        os.environ["CSGPU_GEMM_MIN_BATCH"] = "2"
        f, rf = t(k)
        os.environ.pop("CSGPU_GEMM_MIN_BATCH")
        st.set_tensor_prefilter(True)
        b, rb = t(k)
        st.set_tensor_prefilter(False)
        print(f"9 variants x top-{k}, {n} x {d}: scan route {a:.3f} ms | default routing {c:.3f} ms | tf32 batch forced {f:.3f} ms | "
              f"bf16 tensor-prefilter route {b:.3f} ms | identical: {same(ra, rc) and same(ra, rf) and same(ra, rb)}", flush=True)
    st.close()
