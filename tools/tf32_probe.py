#!/usr/bin/env python
"""The tf32 batch route on the default fp32 index (gemm_tf32.cuh): bit-identity with the single-query kernel and timings
over batch sizes, against the fp32 SIMT route and the multi-query scan.   python tools/tf32_probe.py [rows] [dim]
A diagnostic (numbers land in profiles/), not the headline bench."""
import json, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import codesearch_b200 as cs
from codesearch_b200 import _lib

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 384
cases = sys.argv[3] if len(sys.argv) > 3 else "2:10,8:10,16:10,16:100,32:100,64:100,128:100,256:100,1024:10,1024:100"
lib = _lib.load()
st = cs.VectorStore.new(None, dim)
st.reserve(rows)
st.append_synthetic(1234, 0, rows)
st.build_index()
qs = np.empty((1024, dim), np.float32)
_lib.check(lib.csgpu_synth_rows_host(st.handle, 4321, 0, 1024, qs.ctypes.data_as(_lib._f32p)))
ROUTES = {0: "none", 1: "simt_f32", 2: "tc_bf16", 3: "tc_tf32"}


def run(b, k, reps):
    st.search_batch_ids(qs[:b], k)
    l0 = lib.csgpu_kernel_launches()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = st.search_batch_ids(qs[:b], k)
    dt = (time.perf_counter() - t0) / reps
    s = st.device_stats()
    return out, {"ms": round(dt * 1e3, 3), "device_ms": round(s.last_search_us / 1e3, 3), "launches": int((lib.csgpu_kernel_launches() - l0) // reps),
                 "route": ROUTES.get(s.batch_route, "?"), "rescored_per_query": round(s.prefilter_rescored / b, 1), "filter_max_err": float(s.filter_max_err)}


for case in cases.split(","):
    b, k = (int(x) for x in case.split(":"))
    reps = 5 if b <= 128 else 3
    os.environ["CSGPU_GEMM_MIN_BATCH"] = "2"      # every batch >= 2 takes the GEMM-shaped route
    os.environ["CSGPU_BATCH_SIMT"] = "0"
    (oi, od, on), tf = run(b, k, reps)
    bad = 0
    step = max(1, b // 16)
    for j in range(0, b, step):
        gi, gd = st.search_ids(qs[j], k)
        bad += int(not (np.array_equal(oi[j], gi) and np.array_equal(od[j].view(np.uint32), gd.view(np.uint32))))
    rec = {"rows": rows, "dim": dim, "batch": b, "k": k, "tf32": tf, "tf32_vs_single_query_kernel": "bit-identical" if bad == 0 else f"{bad} MISMATCHES of {len(range(0, b, step))}"}
    if b <= 32:
        os.environ["CSGPU_GEMM_MIN_BATCH"] = "100000"   # multi-query scan
        _, mq = run(b, k, reps)
        rec["multi_query_scan"] = {"ms": mq["ms"], "device_ms": mq["device_ms"], "launches": mq["launches"]}
    elif b <= 128 or os.environ.get("PROBE_SIMT_ALL"):
        os.environ["CSGPU_GEMM_MIN_BATCH"] = "2"
        os.environ["CSGPU_BATCH_SIMT"] = "1"
        _, sm = run(b, k, 1)
        rec["simt"] = {"ms": sm["ms"], "device_ms": sm["device_ms"], "route": sm["route"]}
    print(json.dumps(rec), flush=True)
