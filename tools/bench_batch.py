#!/usr/bin/env python
"""Throughput of csgpu_search_batch on the 10M x 384 corpus, one GPU (BASELINE configs[2]):
fp32 index (multi-query scan / SIMT batched kernel) and the opt-in bf16 index (tcgen05 kernel).
Not the headline bench (bench.py is); numbers land in profiles/.

  python tools/bench_batch.py --dtype fp32 --cases 1:10,8:10,1024:100
  python tools/bench_batch.py --dtype bf16 --cases 1024:100 --recall
"""
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
p.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"])
p.add_argument("--cases", default="1:10,2:10,4:10,8:10,9:200,8:100,16:100,64:100,1024:100")
p.add_argument("--prefilter", action="store_true", help="fp32 index + tensor prefilter (bf16 shadow on tcgen05, exact fp32 rescoring)")
p.add_argument("--recall", action="store_true", help="bf16: recall@k vs the fp32 index on the same rows (first 64 queries)")
args = p.parse_args()
lib = _lib.load()
st = cs.VectorStore.new(None, args.dim, dtype=args.dtype)
st.reserve(args.rows)
st.append_synthetic(1234, 0, args.rows)
if args.prefilter:
    st.set_tensor_prefilter(True)
st.build_index()
qs = np.empty((1024, args.dim), np.float32)
_lib.check(lib.csgpu_synth_rows_host(st.handle, 4321, 0, 1024, qs.ctypes.data_as(_lib._f32p)))
flop_per_q = 2.0 * args.rows * args.dim
last = None
for case in args.cases.split(","):
    b, k = (int(x) for x in case.split(":"))
    reps = max(1, args.reps // max(1, b // 16))
    last = st.search_batch_ids(qs[:b], k)
    l0 = lib.csgpu_kernel_launches()
    t0 = time.perf_counter()
    for _ in range(reps):
        st.search_batch_ids(qs[:b], k)
    dt = (time.perf_counter() - t0) / reps
    launches = (lib.csgpu_kernel_launches() - l0) // reps
    rec = {"dtype": args.dtype + ("+tensor-prefilter" if args.prefilter else ""), "batch": b, "k": k, "ms": round(dt * 1e3, 3), "qps": round(b / dt, 1),
           "TFLOPs": round(b * flop_per_q / dt / 1e12, 2), "device_ms": round(st.device_stats().last_search_us / 1e3, 3),
           "launches": int(launches)}
    if args.prefilter:
        rec["rescored_rows_per_query"] = round(st.device_stats().prefilter_rescored / max(1, min(b, 1024)), 1)
    print(json.dumps(rec), flush=True)

if args.prefilter:   # exactness at full size: the batch result must be bit-identical to the single-query kernel
    b, k = 1024, 100
    oi, od, on = st.search_batch_ids(qs[:b], k)
    bad = 0
    for j in range(0, b, 16):
        gi, gd = st.search_ids(qs[j], k)
        bad += int(not (np.array_equal(oi[j], gi) and np.array_equal(od[j].view(np.uint32), gd.view(np.uint32))))
    print(json.dumps({"prefilter_vs_single_query_kernel": "bit-identical" if bad == 0 else f"{bad} MISMATCHES", "queries_checked": b // 16,
                      "rows": args.rows, "k": k}), flush=True)

if args.recall and args.dtype == "bf16":
    b, k = 64, 100
    oi, od, on = st.search_batch_ids(qs[:b], k)
    st.close()
    ref = cs.VectorStore.new(None, args.dim)
    ref.reserve(args.rows)
    ref.append_synthetic(1234, 0, args.rows)
    ref.build_index()
    ri, rd, rn = ref.search_batch_ids(qs[:b], k)
    recall, err = [], []
    for j in range(b):
        f = dict(zip(ri[j].tolist(), rd[j].tolist()))
        recall.append(len(set(ri[j].tolist()) & set(oi[j].tolist())) / k)
        err += [abs(d - f[i]) for i, d in zip(oi[j].tolist(), od[j].tolist()) if i in f]
    print(json.dumps({"bf16_recall_at_100_vs_fp32_exact": round(float(np.mean(recall)), 4), "min": float(np.min(recall)),
                      "abs_dist_err_mean": float(np.mean(err)), "abs_dist_err_max": float(np.max(err)),
                      "queries": b, "rows": args.rows}), flush=True)
