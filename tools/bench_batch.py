#!/usr/bin/env python
"""Throughput of csgpu_search_batch (multi-query single-pass scan) on the 10M x 384 corpus, one GPU.
Not the headline bench (bench.py is); numbers land in profiles/."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import codesearch_b200 as cs
from codesearch_b200 import _lib

p = argparse.ArgumentParser()
p.add_argument("--rows", type=int, default=10_000_000)
p.add_argument("--dim", type=int, default=384)
p.add_argument("--reps", type=int, default=10)
args = p.parse_args()
lib = _lib.load()
st = cs.VectorStore.new(None, args.dim)
st.reserve(args.rows)
st.append_synthetic(1234, 0, args.rows)
st.build_index()
qs = np.empty((1024, args.dim), np.float32)
_lib.check(lib.csgpu_synth_rows_host(st.handle, 4321, 0, 1024, qs.ctypes.data_as(_lib._f32p)))
out = []
for b, k in [(1, 10), (2, 10), (4, 10), (8, 10), (9, 200), (8, 100), (16, 100), (64, 100), (1024, 100)]:
    reps = max(1, args.reps // max(1, b // 16))
    st.search_batch_ids(qs[:b], k)
    t0 = time.perf_counter()
    for _ in range(reps):
        st.search_batch_ids(qs[:b], k)
    dt = (time.perf_counter() - t0) / reps
    passes = (b + 7) // 8
    out.append({"batch": b, "k": k, "ms": round(dt * 1e3, 3), "qps": round(b / dt, 1),
                "ms_per_pass": round(dt * 1e3 / passes, 3),
                "scanned_GBps_per_pass": round(args.rows * args.dim * 4 / (dt / passes) / 1e9, 1)})
    print(json.dumps(out[-1]), flush=True)
