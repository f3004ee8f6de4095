#!/usr/bin/env python
"""Single-query scan across k, dim and filter density on one GPU (BASELINE configs[1] and the per-GPU part of
configs[4]: 768-d, k=200, file/language filter mask). Device time = csgpu_stats.last_search_us (CUDA events
around the scan inside csgpu_search*), median over --reps queries. Numbers land in profiles/.

  python tools/bench_single.py --dim 384 --rows 10000000 --ks 10,100,200
  python tools/bench_single.py --dim 768 --rows 5000000 --ks 200 --densities 1.0,0.25,0.01
"""
import argparse, json, os, statistics, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import codesearch_b200 as cs
from codesearch_b200 import _lib

p = argparse.ArgumentParser()
p.add_argument("--rows", type=int, default=10_000_000)
p.add_argument("--dim", type=int, default=384)
p.add_argument("--ks", default="10,100,200")
p.add_argument("--densities", default="", help="comma list of filter densities; mask = synthetic file/language predicate")
p.add_argument("--reps", type=int, default=30)
p.add_argument("--byte-prefilter", action="store_true",
               help="csgpu_set_byte_prefilter: int8 shadow + exact rescoring; every timed query is first checked bit-identical to the fp32 scan")
args = p.parse_args()
lib = _lib.load()
n, d = args.rows, args.dim
st = cs.VectorStore.new(None, d)
st.reserve(n)
st.append_synthetic(1234, 0, n)
st.build_index()
qs = np.empty((64, d), np.float32)
_lib.check(lib.csgpu_synth_rows_host(st.handle, 4321, 0, 64, qs.ctypes.data_as(_lib._f32p)))
bytes_scan = n * d * 4
want = {}
if args.byte_prefilter:
    for k in (int(x) for x in args.ks.split(",")):
        want[k] = [st.search_ids(qs[i], k) for i in range(64)]          # fp32 scan kernel, before the shadow exists
    t0 = time.perf_counter()
    st.set_byte_prefilter(True)
    print(json.dumps({"byte_shadow_build_s": round(time.perf_counter() - t0, 3),
                      "byte_shadow_bytes": int(st.device_stats().byte_shadow_bytes)}), flush=True)


def run(k, flt, label):
    dev, wall = [], []
    for i in range(args.reps + 3):
        t0 = time.perf_counter()
        st.search_ids(qs[i % 64], k, flt)
        t1 = time.perf_counter()
        if i >= 3:
            wall.append((t1 - t0) * 1e3)
            dev.append(st.device_stats().last_search_us / 1e3)
    dm, wm = statistics.median(dev), statistics.median(wall)
    rec = {"rows": n, "dim": d, "k": k, "filter": label, "device_ms": round(dm, 4), "e2e_ms": round(wm, 4),
           "scanned_GBps": round(bytes_scan / dm / 1e6, 1), "qps_e2e": round(1e3 / wm, 1)}
    if args.byte_prefilter and flt is None:
        s0 = st.device_stats()
        same = 0
        for i in range(64):
            gi, gd = st.search_ids(qs[i], k)
            same += int(np.array_equal(gi, want[k][i][0]) and np.array_equal(gd.view(np.uint32), want[k][i][1].view(np.uint32)))
        s1 = st.device_stats()
        shadow = int(s1.byte_shadow_bytes)
        rec.update({"route": "byte prefilter (int8 shadow + fp32 rescoring)", "bit_identical_to_fp32_scan": f"{same}/64",
                    "int8_searches": int(s1.byte_searches - s0.byte_searches), "fallbacks": int(s1.byte_fallbacks - s0.byte_fallbacks),
                    "last_candidates": int(s1.byte_candidates), "last_rows_rescored": int(s1.byte_rescored),
                    "shadow_GBps": round(shadow / dm / 1e6, 1), "equivalent_fp32_GBps": rec.pop("scanned_GBps")})
    print(json.dumps(rec), flush=True)


for k in (int(x) for x in args.ks.split(",")):
    run(k, None, "none")
    for dens in (float(x) for x in args.densities.split(",") if x):
        # file_id = row / 37, lang_id = hash(file_id) % 23; predicate "lang in S and file_id in range"
        row = np.arange(n, dtype=np.int64)
        file_id = row // 37
        lang = (file_id * 2654435761 % (1 << 32)) % 23
        if dens >= 1.0:
            mask = np.ones(n, bool)
        else:
            n_lang = max(1, round(23 * min(1.0, dens * 4)))          # languages allowed
            frac_files = dens / (n_lang / 23)                        # leading fraction of files allowed
            mask = (lang < n_lang) & (file_id < frac_files * (n // 37 + 1))
        flt = cs.RowFilter.from_mask(mask)
        run(k, flt, f"density {mask.mean():.4f} (target {dens})")
