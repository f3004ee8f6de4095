import sys, os, json, statistics, time
sys.path.insert(0, os.getcwd())
import numpy as np
import codesearch_b200 as cs
from codesearch_b200 import _lib
lib = _lib.load()
d = 384
for rows in (12500, 25000, 50000, 65000, 80000, 100000, 150000, 200000, 400000):
    st = cs.VectorStore.new(None, d)
    st.append_synthetic(1234, 0, rows)
    st.build_index()
    qs = np.empty((64, d), np.float32)
    _lib.check(lib.csgpu_synth_rows_host(st.handle, 4321, 0, 64, qs.ctypes.data_as(_lib._f32p)))
    for i in range(20): st.search_ids(qs[i], 10)
    dev = []
    t0 = time.perf_counter()
    for i in range(300):
        st.search_ids(qs[i % 64], 10)
    e2e = (time.perf_counter() - t0) / 300
    for i in range(100):
        st.search_ids(qs[i % 64], 10); dev.append(st.device_stats().last_search_us)
    dm = statistics.median(dev)
    print(json.dumps({"rows": rows, "MB": round(rows*d*4/1e6,1), "device_us": round(dm,1), "e2e_us": round(e2e*1e6,1), "GBps": round(rows*d*4/dm/1e3,1)}), flush=True)
    st.close()
