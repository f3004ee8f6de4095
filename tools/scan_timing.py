#!/usr/bin/env python
"""Where a single fp32 query's time goes: CSGPU_SCAN_TIMING=1 makes scan_topk_kernel stamp %globaltimer per CTA and the
library print a one-line timeline per search (a diagnostic, never a benchmark).   python tools/scan_timing.py <rows> <k>"""
import os, sys
os.environ["CSGPU_SCAN_TIMING"] = "1"
sys.path.insert(0, os.getcwd())
import numpy as np
import codesearch_b200 as cs
from codesearch_b200 import _lib
n = int(sys.argv[1]); k = int(sys.argv[2])
st = cs.VectorStore.new(None, 384); st.reserve(n); st.append_synthetic(1234, 0, n); st.build_index()
qs = np.empty((8, 384), np.float32)
_lib.check(_lib.load().csgpu_synth_rows_host(st.handle, 4321, 0, 8, qs.ctypes.data_as(_lib._f32p)))
for i in range(8): st.search_ids(qs[i], k)
print("device_us", st.device_stats().last_search_us)
