#!/usr/bin/env python
"""A/B of the multi-query scan kernels on ONE box: device time of csgpu_search_batch (scan route pinned with
CSGPU_GEMM_MIN_BATCH) at 100k and 10M rows, for this build or for another library given as argv[1] — e.g. the same objects
with an older scan_multi.o linked in (how profiles/r02_multi_tail_ab.txt was made):
    nvcc ... -c <old>/scan_multi.cu -o /tmp/old/scan_multi.o
    nvcc -shared -cudart static build/obj/{csgpu,scan_filtered,gemm_topk,snapshot,scan_i8}.o /tmp/old/scan_multi.o -o build/ab/libcsgpu_oldmulti.so
    python tools/ab_multi.py build/ab/libcsgpu_oldmulti.so; python tools/ab_multi.py
A diagnostic, not a benchmark line."""
import sys, os, json, time
sys.path.insert(0, os.getcwd())
import numpy as np
from codesearch_b200 import _lib
if len(sys.argv) > 1: _lib.LIB_PATH = sys.argv[1]
import codesearch_b200 as cs
lib = _lib.load()
os.environ["CSGPU_GEMM_MIN_BATCH"] = "100000"
for rows in (100000, 10000000):
    st = cs.VectorStore.new(None, 384); st.reserve(rows); st.append_synthetic(1234, 0, rows); st.build_index()
    qs = np.empty((16, 384), np.float32)
    _lib.check(lib.csgpu_synth_rows_host(st.handle, 4321, 0, 16, qs.ctypes.data_as(_lib._f32p)))
    for b, k in ((2, 10), (8, 10), (16, 10), (8, 32)):
        for _ in range(3): st.search_batch_ids(qs[:b], k)
        ds = []
        for _ in range(10):
            st.search_batch_ids(qs[:b], k); ds.append(st.device_stats().last_search_us)
        print(os.path.basename(_lib.LIB_PATH), rows, b, k, "device_us", round(sorted(ds)[len(ds)//2], 1), flush=True)
    st.close()
