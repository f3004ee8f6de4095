#!/usr/bin/env python
"""bench.py — the hot path's headline benchmark (BASELINE.json: queries/sec & scanned HBM GB/s,
10M x 384 fp32, single-query top-10; 1/2/4/8 GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun ... bench.py --gpus N ...      (one rank per GPU; NCCL only bootstraps)

A "step" is ONE single-query search (top-10) over the whole corpus. Weak scaling: every GPU holds
`--rows-per-gpu` (10M) rows of the synthetic corpus, so at N GPUs one query scans N x 10M rows (row-sharded; one fused
scan + top-k + cross-GPU exchange + merge kernel per rank). `value` = whole-job scanned GB/s (N x 15.36 GB / step
time); `qps` rides along. The corpus (15.36 GB per GPU) is >100x the 126 MB L2, so no L2 flush is needed between steps.

What the JSON line carries besides the contract keys (all measured in this run, outside the headline's timed region):
  parity       every rank's answers for PARITY_QUERIES queries against the streaming f64 oracle over ITS 10M-row shard,
               merged by key across ranks (N > 1: also fused == NCCL arm bit for bit, identical on every rank, exchange
               status clean). A mismatch makes the run exit non-zero.
  skew         N > 1: per-step wait of each rank's exchange tail for its slowest peer (globaltimer stamps in the kernel)
  extras       N = 1: BASELINE configs[2] (fp32 SIMT batch, bf16 tcgen05 index, tensor prefilter), configs[4]'s per-GPU
               leg (768-d, top-200, tag predicate at three densities), the 50M-row shard of configs[3], GPU latency at
               100k / 200k rows, the opt-in byte prefilter.
               N > 1: configs[3] (50M rows per GPU: 100M / 200M / 400M rows) with planted-neighbour checks, configs[4]
               sharded, and the SAME job through the in-process C ABI (one process, csgpu_create(devices, N) +
               csgpu_search: what a one-process Rust host binds).
  cpu_baseline N = 1: the reference's exact scan restated in C (oracle/), un-scaled 10M rows all threads + single
               thread, and BASELINE configs[0] (100k rows) single thread / all threads.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED_CORPUS, SEED_QUERY = 1234, 4321
N_QUERIES = 64
PARITY_QUERIES = 2
SCORE_TOL = 1e-5          # north-star tolerance for fp32 distances (tests/parity.py)
TIE_EPS = 4e-7            # f64 gap below which fp32 cannot be expected to order two rows (tests/parity.py)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=200)
    p.add_argument("--warmup", type=int, default=10)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--rows-per-gpu", type=int, default=10_000_000)
    p.add_argument("--dim", type=int, default=384)
    p.add_argument("--k", type=int, default=10)
    p.add_argument("--cpu-max-rows", type=int, default=40_000_000,
                   help="most rows the CPU arms materialise in host RAM (40M x 384 fp32 = 61 GB); above it the arm scans "
                        "this many and scales by bytes, and says so")
    p.add_argument("--cpu-seconds", type=float, default=8.0)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-parity", action="store_true")
    p.add_argument("--no-extras", action="store_true", help="skip everything reported under 'extras'")
    p.add_argument("--extras", default="all",
                   help="comma list of extras to run: batch,bf16,prefilter,byte,c5,c4,small,inproc (default all that apply)")
    p.add_argument("--c4-rows-per-gpu", type=int, default=50_000_000)
    p.add_argument("--byte-prefilter", action="store_true",
                   help="NOT the headline: the same job with csgpu_set_byte_prefilter on every rank (int8 shadow streamed as a "
                        "filter + exact fp32 rescoring, bit-identical results); the line is re-labelled and its roofline counts "
                        "the shadow bytes")
    p.add_argument("--exchange", default="fused", choices=["fused", "nccl"],
                   help="N>1: fused = scan kernel writes its keys into the peers' HBM and merges in its tail (1 launch); "
                        "nccl = scan kernel + NCCL all-gather + merge kernel")
    return p.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            j = json.load(f)
            return j, float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return {}, 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:  # noqa: BLE001
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arms (the oracle is loaded ONLY here and in the parity check: bench.py's cpu_baseline / reference legs)
# ------------------------------------------------------------------------------------------------
def _time_cpu_scan(O, rows, qs, k, seconds, max_iters, warm=True):
    if warm:
        O.cpu_baseline_search(rows, qs[0], k)
    t0 = time.perf_counter()
    it = 0
    while True:
        O.cpu_baseline_search(rows, qs[it % len(qs)], k)
        it += 1
        if time.perf_counter() - t0 >= seconds or it >= max_iters:
            break
    return (time.perf_counter() - t0) / it, it


def cpu_baseline(args, gpu_small=None):
    """The reference's exact scan (examples/benchmark_models.rs:155-165,323-328) generalised to top-k, restated in C
    (oracle/oracle.c cs_cpu_baseline_search). Un-scaled: the 10M x 384 corpus is materialised in host RAM (15.36 GB,
    same generator/seed as the GPU corpus) and scanned whole — all host threads, and single-thread (faithful to the
    reference's sequential loop). Plus BASELINE configs[0] (100k x 384), the reference's real operating scale
    (src/constants.rs:93-95), single-thread and all threads."""
    from oracle import oracle as O
    O.build()
    all_threads = len(os.sched_getaffinity(0))   # torchrun exports OMP_NUM_THREADS=1; the baseline is ALL threads this process may use
    O.set_threads(all_threads)
    d, k = args.dim, args.k
    qs = O.synth_rows(SEED_QUERY, 0, N_QUERIES, d)
    out = {"unit": "GB/s", "kind": "port", "cores": all_threads,
           "what": "C restatement of the reference's exact scan (oracle/oracle.c: f32 dot + norms per row, OpenMP over rows, "
                   "per-thread heaps); NOT arroy+LMDB, which cannot be built here (no cargo/rustc)"}
    # ---- configs[0]: 100k x 384
    small = O.synth_rows(SEED_CORPUS, 0, 100_000, d)
    c0 = {}
    for name, nt in (("single_thread", 1), ("all_threads", all_threads)):
        O.set_threads(nt)
        dt, it = _time_cpu_scan(O, small, qs, k, 1.5, 400)
        c0[name] = {"cores": nt, "ms_per_query": round(dt * 1e3, 3), "qps": round(1 / dt, 1),
                    "GBps": round(small.nbytes / dt / 1e9, 2), "queries": it}
    out["configs0_100k_x_384"] = c0
    del small
    # ---- configs[1]: 10M x 384, un-scaled when the host has the RAM
    n = args.rows_per_gpu
    scaled = False
    try:
        avail_kb = next(int(l.split()[1]) for l in open("/proc/meminfo") if l.startswith("MemAvailable"))
    except Exception:  # noqa: BLE001
        avail_kb = 0
    need = n * d * 4
    if avail_kb and avail_kb * 1024 < need * 1.25:
        n = max(1_000_000, int(avail_kb * 1024 * 0.5 // (d * 4)))
        scaled = True
    O.set_threads(all_threads)
    t0 = time.perf_counter()
    try:
        rows = O.synth_rows(SEED_CORPUS, 0, n, d)
    except MemoryError:
        n, scaled = 1_000_000, True
        rows = O.synth_rows(SEED_CORPUS, 0, n, d)
    gen_s = time.perf_counter() - t0
    dt, it = _time_cpu_scan(O, rows, qs, k, args.cpu_seconds, 400)
    gbs = rows.nbytes / dt / 1e9
    O.set_threads(1)
    dt1, it1 = _time_cpu_scan(O, rows, qs, k, 4.0, 3, warm=False)
    O.set_threads(all_threads)
    ids_cpu, _ = O.cpu_baseline_search(rows, qs[0], k)
    del rows
    out.update({
        "value": round(gbs, 3), "scaled": scaled, "rows_scanned": n,
        "sample": f"{n} x {d} fp32 rows (same generator/seed as the GPU corpus; generated in {gen_s:.1f} s), top-{k}, "
                  f"{it} queries on {all_threads} threads, {dt * 1e3:.1f} ms/query"
                  + ("" if not scaled else f"; host RAM too small for {args.rows_per_gpu} rows: scaled by bytes"),
        "ms_per_query": round(dt * 1e3, 2), "qps": round(1 / dt * n / args.rows_per_gpu, 4),
        "single_thread": {"cores": 1, "ms_per_query": round(dt1 * 1e3, 1), "GBps": round(n * d * 4 / dt1 / 1e9, 2),
                          "qps": round(1 / dt1 * n / args.rows_per_gpu, 4), "queries": it1},
        "top_ids_query0": [int(i) for i in ids_cpu],
    })
    if gpu_small:
        out["gpu_beside_it"] = gpu_small
    return out


def run_reference(args):
    """`--impl reference`: the reference arm = the CPU exact scan on this box's host cores (kind "port": the real arroy +
    LMDB path is unbuildable here), all threads, on this arm's config: N x rows-per-gpu rows in host RAM, un-scaled up to
    --cpu-max-rows (above that: that many rows, scaled by bytes, flagged). A step = one query over the rows held."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    nt = len(os.sched_getaffinity(0))
    O.set_threads(nt)
    world, d, k = args.gpus, args.dim, args.k
    total = args.rows_per_gpu * world
    n = min(total, args.cpu_max_rows)
    try:
        avail_kb = next(int(l.split()[1]) for l in open("/proc/meminfo") if l.startswith("MemAvailable"))
        while n > 1_000_000 and avail_kb * 1024 < n * d * 4 * 1.25:
            n //= 2
    except Exception:  # noqa: BLE001
        pass
    scaled = n < total
    t0 = time.perf_counter()
    rows = O.synth_rows(SEED_CORPUS, 0, n, d)
    gen_s = time.perf_counter() - t0
    qs = O.synth_rows(SEED_QUERY, 0, N_QUERIES, d)
    for i in range(args.warmup):
        O.cpu_baseline_search(rows, qs[i % N_QUERIES], k)
    t0 = time.perf_counter()
    for i in range(args.steps):
        O.cpu_baseline_search(rows, qs[i % N_QUERIES], k)
    dt = (time.perf_counter() - t0) / args.steps
    gbs = n * d * 4 / dt / 1e9
    ms_full = dt * 1e3 * total / n
    sample = (f"{n} x {d} fp32 rows in host RAM (generated in {gen_s:.1f} s), top-{k}, {args.steps} timed queries after "
              f"{args.warmup} warm-ups, {dt * 1e3:.1f} ms per pass"
              + ("" if not scaled else f"; the job is {total} rows: time scaled by bytes x{total / n:.2f}"))
    cb = {"value": round(gbs, 3), "unit": "GB/s", "cores": nt, "kind": "port", "scaled": scaled, "rows_scanned": n, "sample": sample}
    line = {
        "impl": "reference", "metric": "scanned_GBps_single_query_top10_fp32", "value": round(gbs, 3), "unit": "GB/s",
        "qps": round(1e3 / ms_full, 5),
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_full, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{total}x{d} fp32 corpus, single-query top-{k} (BASELINE configs[1] per GPU); CPU exact scan "
                               f"(reference restatement, {nt} threads) over {n} rows" + (" — scaled" if scaled else " — un-scaled"),
                   "rows_per_gpu": args.rows_per_gpu, "dim": d, "k": k},
        "cpu_baseline": cb,
        "e2e": {"value": round(gbs, 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# parity (checker: the oracle; never inside a timed region)
# ------------------------------------------------------------------------------------------------
def merge_by_key(lists, k):
    """k-way merge of per-shard (ids, d32, d64) oracle lists by (distance f32, id) — the order of the API."""
    ids = np.concatenate([l[0] for l in lists]); d32 = np.concatenate([l[1] for l in lists]); d64 = np.concatenate([l[2] for l in lists])
    order = np.lexsort((ids, d32))
    order = order[:k]
    return ids[order], d32[order], d64[order]


def compare_with_oracle(g_ids, g_dist, o_ids, o_d32, o_d64, k):
    """tests/parity.py's rule: ids bit-exact in (distance, id) order and |d - oracle| <= 1e-5; an id swap is accepted only
    inside a group the oracle's own f64 distances cannot separate (< 4e-7). Returns (ok, swaps, max_abs_err, why)."""
    g_ids = np.asarray(g_ids); g_dist = np.asarray(g_dist, dtype=np.float32)
    if len(g_ids) != k:
        return False, 0, None, f"expected {k} results, got {len(g_ids)}"
    if len(set(g_ids.tolist())) != k:
        return False, 0, None, "duplicate ids"
    for i in range(1, k):
        if not (g_dist[i - 1], g_ids[i - 1]) < (g_dist[i], g_ids[i]):
            return False, 0, None, f"not ascending at {i}"
    d64 = {int(i): float(x) for i, x in zip(o_ids, o_d64)}
    swaps, err = 0, 0.0
    for i in range(k):
        g = int(g_ids[i])
        if g not in d64:
            return False, swaps, err, f"rank {i}: id {g} is not in the oracle's top-{len(o_ids)}"
        err = max(err, abs(float(g_dist[i]) - d64[g]))
        if g != int(o_ids[i]):
            if abs(d64[g] - float(o_d64[i])) >= TIE_EPS:
                return False, swaps, err, f"rank {i}: id {g} vs oracle id {int(o_ids[i])} is not a near-tie"
            swaps += 1
    if err > SCORE_TOL:
        return False, swaps, err, f"distance error {err:.3g} > {SCORE_TOL}"
    return True, swaps, err, ""


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
    torch.cuda.set_device(local_rank)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        from datetime import timedelta
        cpu_group = dist.new_group(backend="gloo", timeout=timedelta(seconds=600))   # host-side barriers / gathers that must not put kernels on the GPUs
    assert world == args.gpus or world == 1 and args.gpus == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    import codesearch_b200 as cs
    from codesearch_b200 import _lib
    from codesearch_b200.sharded import ShardedSearcher, decode_keys
    lib = _lib.load()  # raises if libcsgpu.so is missing: no fallback

    want_extras = set() if args.no_extras else set(
        "batch,bf16,prefilter,byte,c5,c4,small,inproc".split(",") if args.extras == "all" else args.extras.split(","))
    n, d, k = args.rows_per_gpu, args.dim, args.k
    peaks_json, peak, peak_src = peaks()

    def cpu_barrier():
        if world > 1:
            dist.barrier(group=cpu_group)

    def gather_obj(obj):
        if world == 1:
            return [obj]
        out = [None] * world
        dist.all_gather_object(out, obj, group=cpu_group)
        return out

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return float(x)
        return max(gather_obj(float(x)))

    store = cs.VectorStore.new(None, d, devices=[local_rank])
    store.reserve(n)
    store.append_synthetic(SEED_CORPUS, rank * n, n, 0)   # chunk id = global row index
    torch.cuda.synchronize()
    t_build = time.perf_counter()
    store.build_index()                                   # normalise to unit length (f64 norms), drop dead / zero / non-finite rows
    torch.cuda.synchronize()
    build_ms = (time.perf_counter() - t_build) * 1e3      # what replaces the reference's arroy tree build (store.rs:386-430)
    if args.byte_prefilter:
        store.set_byte_prefilter(True)
    searcher = ShardedSearcher(store, k_max=max(k, 16), exchange=args.exchange)

    # queries: same generator, different seed; produced by the device generator, kept on host AND device
    q_host = np.empty((N_QUERIES, d), dtype=np.float32)
    _lib.check(lib.csgpu_synth_rows_host(store.handle, SEED_QUERY, 0, N_QUERIES, q_host.ctypes.data_as(_lib._f32p)))
    q_dev = torch.from_numpy(q_host).cuda()

    # ---------------- device-resident timed region: `value` -----------------------------------
    for i in range(args.warmup):
        searcher.search_keys_device(q_dev[i % N_QUERIES], k)
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.csgpu_kernel_launches()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    scan_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    stream = torch.cuda.current_stream().cuda_stream
    local = searcher.local[:k]
    ev[0].record()
    for i in range(args.steps):
        q = q_dev[i % N_QUERIES]
        if world == 1:
            scan_ev[i][0].record()
            searcher.search_keys_device(q, k)
            scan_ev[i][1].record()
        elif args.exchange == "fused":
            scan_ev[i][0].record()     # one kernel: scan + peer stores over NVLink + flag wait + global merge
            _lib.check(lib.csgpu_search_keys_exchange_device(store.handle, q.data_ptr(), k,
                                                             searcher.out[:k].data_ptr(), stream))
            scan_ev[i][1].record()
        else:
            scan_ev[i][0].record()
            _lib.check(lib.csgpu_search_keys_device(store.handle, q.data_ptr(), k, local.data_ptr(), stream))
            scan_ev[i][1].record()
            gathered = searcher.gathered[: world * k]
            dist.all_gather_into_tensor(gathered, local)
            _lib.check(lib.csgpu_merge_keys_device(store.handle, gathered.data_ptr(), world, k,
                                                   searcher.out[:k].data_ptr(), stream))
    ev[1].record()
    sync_all()
    launches = lib.csgpu_kernel_launches() - launches0
    elapsed_ms = ev[0].elapsed_time(ev[1])
    scan_ms = [a.elapsed_time(b) for a, b in scan_ev]
    elapsed_ms = max_over_ranks(elapsed_ms)
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- per-step cross-rank skew (N > 1, fused): measured inside the kernels -----------------
    skew = None
    if world > 1 and args.exchange == "fused":
        try:
            nmax = min(args.steps, 128)
            ns = np.zeros(nmax * world, dtype=np.uint64)
            nq = ctypes.c_uint32(0)
            _lib.check(lib.csgpu_exchange_wait_stats(store.handle, ns.ctypes.data_as(_lib._u64p), nmax, ctypes.byref(nq)))
            mine = ns[: nq.value * world].reshape(nq.value, world).max(axis=1).astype(np.float64) / 1e3   # us this rank waited, per step
            allw = np.array(gather_obj(mine.tolist()))                                                   # [world][steps]
            if rank == 0:
                per_step = allw.max(axis=0)          # the fastest rank's wait = spread between first and last rank of that step
                slowest = allw.argmin(axis=0)        # the rank that waited least arrived last
                skew = {"what": "per step, the longest any rank's exchange tail waited for a peer's keys after publishing its own "
                                "(globaltimer stamps in the kernel; = finish-time spread between the fastest and the slowest GPU)",
                        "steps": int(nq.value), "us_mean": round(float(per_step.mean()), 1), "us_p50": round(float(np.percentile(per_step, 50)), 1),
                        "us_p90": round(float(np.percentile(per_step, 90)), 1), "us_max": round(float(per_step.max()), 1),
                        "slowest_rank_histogram": np.bincount(slowest, minlength=world).tolist(),
                        "mean_wait_us_per_rank": [round(float(x), 1) for x in allw.mean(axis=1)]}
        except Exception as e:  # noqa: BLE001
            skew = {"error": repr(e)}

    # ---------------- end to end through the public API: `e2e` --------------------------------
    e2e = None
    if not args.no_e2e:
        e2e_steps = args.steps
        if world == 1:
            def one(i):
                return store.search_ids(q_host[i % N_QUERIES], k)      # C ABI, host pointers in/out
        else:
            def one(i):
                return searcher.search(q_host[i % N_QUERIES], k)       # pinned H2D, fused scan+exchange+merge, D2H
        for i in range(max(3, args.warmup)):
            one(i)
        sync_all()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            ids, dd = one(i)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e_ms = dt / e2e_steps * 1e3
        e2e = {"value": round(world * n * d * 4 / (e2e_ms * 1e-3) / 1e9, 2), "unit": "GB/s",
               "qps": round(1e3 / e2e_ms, 3), "ms_per_step": round(e2e_ms, 4),
               "h2d_bytes_per_step": d * 4, "d2h_bytes_per_step": k * 8,
               "api": "VectorStore.search_ids -> csgpu_search (host pointers)" if world == 1 else
                      ("ShardedSearcher.search -> csgpu_search_exchange (host pointers: pinned H2D, ONE fused launch = scan + peer-memory exchange + "
                       "merge, its last CTA writes the keys into mapped host memory)"
                       if args.exchange == "fused" else
                       "ShardedSearcher.search (pinned H2D, csgpu_search_keys_device, NCCL all-gather, csgpu_merge_keys_device, D2H)")}

    # ---------------- parity: every rank, against the oracle over its own shard ---------------------------
    parity = None
    fused_answers = []    # (ids, dist) of the first queries, for the in-process C-ABI comparison below
    if not args.no_parity:
        parity = check_parity(args, cs, _lib, lib, dist, store, searcher, q_host, rank, world, gather_obj, fused_answers)

    line = None
    if rank == 0:
        ms_per_step = elapsed_ms / args.steps
        total_bytes = world * n * d * 4
        value = total_bytes / (ms_per_step * 1e-3) / 1e9
        kern_ms = statistics.mean(scan_ms)
        achieved = n * d * 4 / (kern_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                traffic = json.load(f).get(f"scan_topk_kernel<3,true,4,false>|rows={n}|dim={d}|k={k}", {}).get("traffic_bytes")
        except Exception:  # noqa: BLE001
            pass
        line = {
            "metric": "scanned_GBps_single_query_top10_fp32", "value": round(value, 2), "unit": "GB/s",
            "qps": round(1e3 / ms_per_step, 3),
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{n * world}x{d} fp32 unit-vector corpus ({n} rows per GPU, row-sharded), "
                                   f"single-query exact cosine top-{k} (BASELINE configs[1] per GPU)",
                       "rows_per_gpu": n, "dim": d, "k": k, "queries": N_QUERIES,
                       "l2": "no flush needed: 15.36 GB scanned per GPU per step >> 126 MB L2",
                       "parallelism": f"row-shard x{world}" + ("" if world == 1 else (
                           " + exchange fused into the scan kernel (peer stores over NVLink, flags, in-kernel merge)"
                           if args.exchange == "fused" else " + NCCL all-gather of k keys + merge kernel"))},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": traffic,
                         "kernel": "scan_topk_kernel<3,true,4,false>", "kernel_ms": round(kern_ms, 4),
                         "algorithmic_bytes_per_launch": n * d * 4, "peak_source": peak_src,
                         "frac_of_nominal_8TBps": round(achieved / 8000.0, 4)},
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "parity": parity,
            "build_index_ms": round(build_ms, 2),   # csgpu_build of this rank's rows (the reference: arroy tree build, 10-30 s on
                                                    # large sets by its own comment, src/index/mod.rs:773); not part of any timing above
        }
        if skew is not None:
            line["skew"] = skew
        if args.byte_prefilter:
            relabel_byte_prefilter(line, store, world, args, kern_ms, achieved, peak, e2e, n, d, k)

    # ---------------- extras (outside every headline timing) --------------------------------------------
    extras = {}
    if want_extras and not args.byte_prefilter:
        if world == 1:
            if "batch" in want_extras or "prefilter" in want_extras or "byte" in want_extras:
                extras.update(extras_batch(store, lib, n, d, peaks_json, want_extras))
        # everything below needs the HBM the headline store holds
        del searcher
        store.close()
        torch.cuda.empty_cache()
        cpu_barrier()
        if world == 1:
            if "bf16" in want_extras:
                extras["batch_bf16_tcgen05_index"] = guarded(lambda: extras_bf16(cs, _lib, lib, n, d, peaks_json))
            if "small" in want_extras:
                extras["small_corpus_latency"] = guarded(lambda: extras_small(cs, _lib, lib, d, k))
        else:
            if "inproc" in want_extras:
                res = None
                if rank == 0:
                    res = guarded(lambda: extras_in_process(cs, _lib, lib, world, n, d, k, args, q_host, fused_answers, e2e))
                cpu_barrier()   # the other ranks keep their GPUs idle (gloo barrier: nothing is launched) while rank 0 drives all of them
                if rank == 0:
                    extras["in_process_c_abi"] = res
        if "c5" in want_extras:
            r = guarded(lambda: extras_c5(cs, _lib, lib, torch, dist, rank, world, local_rank, gather_obj, sync_all))
            if rank == 0:
                extras["c5_hybrid_vector_leg_768d_top200_filtered"] = r
        if "c4" in want_extras:
            r = guarded(lambda: extras_c4(cs, _lib, lib, torch, dist, rank, world, local_rank, args, gather_obj, sync_all, max_over_ranks))
            if rank == 0:
                extras["c4_50M_rows_per_gpu"] = r
    if rank == 0:
        if extras:
            line["extras"] = extras
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = guarded(lambda: cpu_baseline(args, extras.get("small_corpus_latency")))
        print(json.dumps(line), flush=True)
    bad = parity is not None and not parity.get("ok", False)
    if world > 1:
        bad = any(gather_obj(bool(bad)))
        dist.barrier()
        dist.destroy_process_group()
    if bad:
        sys.exit(3)


def guarded(fn):
    try:
        return fn()
    except Exception as e:  # noqa: BLE001 — the headline line must not depend on a secondary measurement
        import traceback
        return {"error": repr(e), "where": traceback.format_exc().strip().splitlines()[-3:]}


def relabel_byte_prefilter(line, store, world, args, kern_ms, achieved, peak, e2e, n, d, k):
    shadow = int(store.device_stats().byte_shadow_bytes)
    line["metric"] = "queries_per_s_single_query_top10_exact_via_byte_prefilter"
    line["value"], line["unit"] = line["qps"], "queries/s"
    line["dtype"] = "s8 filter + f32 rescoring"
    line["config"]["workload"] += "; byte prefilter on (csgpu_set_byte_prefilter): results bit-identical to the fp32 scan"
    line["config"]["l2"] = f"no flush needed: {shadow / 1e9:.2f} GB of shadow scanned per GPU per step >> 126 MB L2"
    if world > 1 and args.exchange == "fused":
        line["config"]["parallelism"] = (f"row-shard x{world} + int8 kernel, conditional fp32 scan (no-op), then the exchange as "
                                         "its own one-CTA launch (peer stores over NVLink, flags, merge)")
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f).get(f"scan_i8_kernel<3,true,6>|rows={n}|dim={d}|k={k}", {}).get("traffic_bytes")
    except Exception:  # noqa: BLE001
        traffic = None
    line["roofline"].update({"achieved": round(shadow / (kern_ms * 1e-3) / 1e9, 1),
                             "frac": round(shadow / (kern_ms * 1e-3) / 1e9 / peak, 4), "traffic": traffic,
                             "kernel": "scan_i8_kernel<3,true,6> (+ no-op conditional scan" + (" + exchange_keys_kernel)" if world > 1 and args.exchange == "fused" else ")"),
                             "algorithmic_bytes_per_launch": shadow,
                             "frac_of_nominal_8TBps": round(shadow / (kern_ms * 1e-3) / 1e9 / 8000.0, 4),
                             "equivalent_fp32_GBps_per_gpu": round(achieved, 1)})
    if e2e:
        e2e["value"], e2e["unit"] = e2e.get("qps"), "queries/s"


def check_parity(args, cs, _lib, lib, dist, store, searcher, q_host, rank, world, gather_obj, fused_answers):
    """VERDICT r1 item 1. Every rank: the streaming f64 oracle over ITS shard for PARITY_QUERIES queries (top k + margin),
    lists all-gathered on the host and merged by key = the oracle's global answer; compared on EVERY rank with what the
    fused exchange returned there (ids bit-exact up to oracle-unresolvable near-ties, |d - oracle| <= 1e-5). N > 1 also:
    fused == NCCL arm bit for bit, every rank holds the identical list, csgpu_exchange_status clean. The reason ids must be
    identical everywhere: order is the API (rank = position, /root/reference/src/rerank/mod.rs:57-66)."""
    from oracle import oracle as O
    O.build()
    nt = max(1, len(os.sched_getaffinity(0)) // max(1, world))
    O.set_threads(nt)
    n, d, k = args.rows_per_gpu, args.dim, args.k
    margin = 8
    t0 = time.perf_counter()
    qs = q_host[:PARITY_QUERIES]
    # the generator check: the device-made queries equal the oracle's generator bit for bit
    gen_same = bool(np.array_equal(O.synth_rows(SEED_QUERY, 0, PARITY_QUERIES, d).view(np.uint32), qs.view(np.uint32)))
    got = [searcher.search(qs[j], k) if world > 1 else store.search_ids(qs[j], k) for j in range(PARITY_QUERIES)]
    fused_answers.extend(got)
    oi, od, o64, on = O.search_synth(SEED_CORPUS, rank * n, n, d, qs, k + margin)
    mine = [(oi[j, : on[j]], od[j, : on[j]], o64[j, : on[j]]) for j in range(PARITY_QUERIES)]
    all_lists = gather_obj(mine)                                     # [world][query] -> per-shard oracle lists
    ok, swaps, err, why = True, 0, 0.0, []
    for j in range(PARITY_QUERIES):
        g_ids, g_d32, g_d64 = merge_by_key([all_lists[r][j] for r in range(world)], k + margin)
        o, s, e, w = compare_with_oracle(got[j][0], got[j][1], g_ids, g_d32, g_d64, k)
        ok = ok and o; swaps += s; err = max(err, e or 0.0)
        if w:
            why.append(f"rank {rank} query {j}: {w}")
    res = {"queries": PARITY_QUERIES, "k": k, "vs_oracle": "exact" if ok and swaps == 0 else ("exact up to f32 near-ties" if ok else "MISMATCH"),
           "near_tie_swaps": swaps, "max_abs_distance_err": float(err), "tolerance": SCORE_TOL,
           "oracle": f"oracle/oracle.c cs_oracle_search_synth (f64) over each rank's own {n}-row shard, {nt} threads per rank, "
                     f"lists merged by (distance, id) on the host; {time.perf_counter() - t0:.1f} s",
           "queries_generator_bit_identical": gen_same}
    ok = ok and gen_same
    if world > 1:
        nccl_same = None
        if args.exchange == "fused":
            nccl = ShardedSearcherNccl(searcher)
            nccl_same = True
            for j in range(PARITY_QUERIES):
                ni, nd = nccl.search(qs[j], k)
                nccl_same = nccl_same and bool(np.array_equal(ni, got[j][0]) and np.array_equal(nd.view(np.uint32), got[j][1].view(np.uint32)))
            ok = ok and nccl_same
        every = gather_obj([(g[0].tolist(), g[1].view(np.uint32).tolist()) for g in got])
        identical = all(e == every[0] for e in every)
        t = ctypes.c_uint32(0)
        status = None
        if args.exchange == "fused":
            _lib.check(lib.csgpu_exchange_status(store.handle, ctypes.byref(t)))
            status = int(t.value)
        oks = gather_obj((bool(ok), why, status))
        ok = all(o[0] for o in oks) and identical and all((o[2] or 0) == 0 for o in oks)
        why = [w for o in oks for w in o[1]]
        res.update({"ranks_identical": bool(identical), "fused_equals_nccl_arm": nccl_same,
                    "exchange_status": [o[2] for o in oks], "checked_on_ranks": world})
    res["ok"] = bool(ok)
    if why:
        res["why"] = why[:8]
    return res


class ShardedSearcherNccl:
    """The comparison arm on the SAME store and buffers: local scan kernel -> NCCL all-gather of k keys -> merge kernel."""

    def __init__(self, fused):
        self.s = fused

    def search(self, q, k):
        import torch
        import torch.distributed as dist
        from codesearch_b200 import _lib
        from codesearch_b200.sharded import decode_keys
        s = self.s
        stream = torch.cuda.current_stream().cuda_stream
        s.q_pin[: q.size].copy_(torch.from_numpy(np.ascontiguousarray(q, dtype=np.float32)))
        s.q_dev.copy_(s.q_pin, non_blocking=True)
        local = s.local[:k]
        _lib.check(s.lib.csgpu_search_keys_device(s.store.handle, s.q_dev.data_ptr(), k, local.data_ptr(), stream))
        gathered = s.gathered[: s.world * k]
        dist.all_gather_into_tensor(gathered, local)
        out = torch.empty(k, dtype=torch.int64, device=s.dev)
        _lib.check(s.lib.csgpu_merge_keys_device(s.store.handle, gathered.data_ptr(), s.world, k, out.data_ptr(), stream))
        return decode_keys(out.cpu().numpy())


# ------------------------------------------------------------------------------------------------
# extras
# ------------------------------------------------------------------------------------------------
def _synth_queries(_lib, lib, store, b, d, seed=SEED_QUERY):
    qs = np.empty((b, d), dtype=np.float32)
    _lib.check(lib.csgpu_synth_rows_host(store.handle, seed, 0, b, qs.ctypes.data_as(_lib._f32p)))
    return qs


def _time_batches(store, lib, qs, k, reps):
    store.search_batch_ids(qs, k)
    l0 = lib.csgpu_kernel_launches()
    t0 = time.perf_counter()
    dev = []
    for _ in range(reps):
        out = store.search_batch_ids(qs, k)
        dev.append(store.device_stats().last_search_us)
    dt = (time.perf_counter() - t0) / reps
    return dt, float(np.median(dev)) / 1e3, int((lib.csgpu_kernel_launches() - l0) // reps), out


def extras_batch(store, lib, n, d, peaks_json, want):
    """Secondary, not the headline: BASELINE configs[2] on the SAME fp32 index through csgpu_search_batch (host buffers in and
    out): (i) the default route, the register-tiled fp32 SIMT kernel (bound: FP32 FMA pipe, 148 SM x 128 lanes x 2 x max
    clock); (ii) the opt-in tensor prefilter (tcgen05 on a bf16 shadow as a filter + exact fp32 rescoring, csrc/rescore.cuh),
    re-checked bit for bit against the single-query kernel; (iii) the opt-in byte prefilter for single queries."""
    from codesearch_b200 import _lib
    out = {}
    b, k = 1024, 100
    qs = _synth_queries(_lib, lib, store, b, d)
    sm_max = float(peaks_json.get("sm_max_mhz", 1965.0))
    fp32_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12
    flop = 2.0 * n * d * b
    if "batch" in want:
        def tf32():
            os.environ.pop("CSGPU_BATCH_SIMT", None)
            smp = ClockSampler(0).start()
            dt, dev_ms, launches, res = _time_batches(store, lib, qs, k, 5)
            clk = smp.stop()
            st = store.device_stats()
            same = True
            for j in range(0, b, 64):
                gi, gd = store.search_ids(qs[j], k)
                same = same and np.array_equal(res[0][j], gi) and np.array_equal(res[1][j].view(np.uint32), gd.view(np.uint32))
            small = {}
            for bb in (2, 8, 16, 32, 64, 128, 256):
                dts, dvs, _, _ = _time_batches(store, lib, qs[:bb], k, 5)
                small[str(bb)] = {"ms": round(dts * 1e3, 3), "device_ms": round(dvs, 3)}
            tf_peak = float(peaks_json.get("bf16_tflops", 0) or 0) / 2.0
            return {"workload": f"{n}x{d} fp32 index, batch {b} x top-{k} (BASELINE configs[2]), csgpu_search_batch on the DEFAULT index: "
                                "tcgen05 kind::tf32 straight off the fp32 rows as a filter + exact fp32 rescoring (csrc/gemm_tf32.cuh), host buffers",
                    "route": {1: "simt_f32", 2: "tc_bf16", 3: "tc_tf32"}.get(int(st.batch_route), str(st.batch_route)),
                    "ms_per_batch": round(dt * 1e3, 3), "device_ms": round(dev_ms, 3), "qps": round(b / dt, 1),
                    "TFLOPs_device": round(flop / (dev_ms * 1e-3) / 1e12, 1),
                    "bound": "tensor (tf32 pipe: half the bf16 rate; ncu of the main phase: tensor pipe 80 % active, its memory side 88 %) above 128 queries; hbm up to 128 queries (ncu: 85 % of DRAM peak)",
                    "frac_of_tf32_tensor_peak": (round(flop / (dev_ms * 1e-3) / 1e12 / tf_peak, 4) if tf_peak else None),
                    "tf32_peak_TFLOPs": (round(tf_peak, 1) if tf_peak else None),
                    "peak_source": "half of MEASURED_PEAKS.json bf16_tflops (tf32 runs at half the bf16 rate)",
                    "shadow_bytes": int(st.shadow_bytes), "gpu_launches_per_batch": launches,
                    "fp32_rows_rescored_per_query": round(int(st.prefilter_rescored) / b, 1),
                    "filter_max_err": float(st.filter_max_err), "filter_margin": 1.1e-3,
                    "bit_identical_to_single_query_kernel": bool(same), "queries_checked": b // 64,
                    "by_batch_size_top100": small, "clocks": clk}
        out["batch_fp32_default_tf32_filter"] = guarded(tf32)

        def simt():
            os.environ["CSGPU_BATCH_SIMT"] = "1"     # the SIMT kernel: the route for dim > 1024, kept measured
            try:
                smp = ClockSampler(0).start()
                dt, dev_ms, launches, res = _time_batches(store, lib, qs, k, 2)
                clk = smp.stop()
                st = store.device_stats()
                same = True
                for j in range(0, b, 128):
                    gi, gd = store.search_ids(qs[j], k)
                    same = same and np.array_equal(res[0][j], gi) and bool(np.abs(res[1][j] - gd).max() <= 1e-6)
            finally:
                os.environ.pop("CSGPU_BATCH_SIMT", None)
            return {"workload": f"{n}x{d} fp32 index, batch {b} x top-{k} (BASELINE configs[2]), register-tiled fp32 SIMT kernel (CSGPU_BATCH_SIMT=1; the route for dim > 1024), host buffers",
                    "route": {1: "simt_f32", 2: "tc_bf16", 3: "tc_tf32"}.get(int(st.batch_route), str(st.batch_route)),
                    "ms_per_batch": round(dt * 1e3, 2), "device_ms": round(dev_ms, 2), "qps": round(b / dt, 1),
                    "TFLOPs": round(flop / dt / 1e12, 2), "bound": "fp32 FMA pipe",
                    "peak_TFLOPs": round(fp32_peak, 1), "frac": round(flop / (dev_ms * 1e-3) / 1e12 / fp32_peak, 4),
                    "peak_source": f"148 SM x 128 lanes x 2 flop x {sm_max:.0f} MHz (MEASURED_PEAKS.json sm_max_mhz)",
                    "gpu_launches_per_batch": launches, "ids_equal_single_query_kernel": bool(same), "queries_checked": b // 128,
                    "clocks": clk}
        out["batch_fp32_simt"] = guarded(simt)
    if "batch" in want:
        def serving():
            """Many concurrent callers of csgpu_search (the reference's server shape: `&self` from many threads,
            src/server/mod.rs:545-548) with the micro-batcher on: a group of up to 128 callers rides one tf32 tensor-core batch."""
            T, R, kk = 64, 12, 10
            want_res = [store.search_ids(qs[j], kk) for j in range(T)]
            t0 = time.perf_counter()
            for j in range(16):
                store.search_ids(qs[j], kk)
            one = (time.perf_counter() - t0) / 16
            store.set_coalescing(True)
            try:
                s0 = store.device_stats()
                bad = []
                start = threading.Barrier(T + 1)

                def work(j):
                    start.wait()
                    for _ in range(R):
                        gi, gd = store.search_ids(qs[j], kk)
                        if not (np.array_equal(gi, want_res[j][0]) and np.array_equal(gd.view(np.uint32), want_res[j][1].view(np.uint32))):
                            bad.append(j)
                ts = [threading.Thread(target=work, args=(j,)) for j in range(T)]
                [t.start() for t in ts]
                start.wait()
                t0 = time.perf_counter()
                [t.join() for t in ts]
                dt = time.perf_counter() - t0
                s1 = store.device_stats()
            finally:
                store.set_coalescing(False)
            return {"workload": f"{n}x{d} fp32 index, {T} host threads x {R} csgpu_search calls each (top-{kk}, host buffers), csgpu_set_coalescing on",
                    "qps": round(T * R / dt, 1), "single_caller_qps": round(1.0 / one, 1),
                    "searches": int(s1.coalesced_queries - s0.coalesced_queries), "corpus_passes": int(s1.coalesced_passes - s0.coalesced_passes),
                    "every_answer_bit_identical_to_its_uncoalesced_search": not bad}
        out["coalesced_serving_64_callers"] = guarded(serving)
    if "prefilter" in want:
        def pref():
            store.set_tensor_prefilter(True)
            try:
                dt, dev_ms, launches, res = _time_batches(store, lib, qs, k, 5)
                same = True
                for j in range(0, b, 128):
                    gi, gd = store.search_ids(qs[j], k)
                    same = same and np.array_equal(res[0][j], gi) and np.array_equal(res[1][j].view(np.uint32), gd.view(np.uint32))
                for i in range(3):
                    store.search_batch_ids(qs[i:i + 1], 10)
                t0 = time.perf_counter()
                for i in range(50):
                    s_i, s_d, _ = store.search_batch_ids(qs[i:i + 1], 10)
                dt1 = (time.perf_counter() - t0) / 50
                gi, gd = store.search_ids(qs[49], 10)
                same1 = bool(np.array_equal(s_i[0], gi) and np.array_equal(s_d[0].view(np.uint32), gd.view(np.uint32)))
            finally:
                store.set_tensor_prefilter(False)
            return {"workload": f"{n}x{d} fp32 index, batch {b} x top-{k} (BASELINE configs[2]), csgpu_search_batch with the tensor prefilter on, host buffers",
                    "ms_per_batch": round(dt * 1e3, 3), "device_ms": round(dev_ms, 3), "qps": round(b / dt, 1), "TFLOPs": round(flop / dt / 1e12, 1),
                    "gpu_launches_per_batch": launches, "bit_identical_to_single_query_kernel": bool(same), "queries_checked": b // 128,
                    "single_query_top10": {"ms_per_query": round(dt1 * 1e3, 3), "qps": round(1.0 / dt1, 1), "bit_identical": same1}}
        out["batch_fp32_tensor_prefilter"] = guarded(pref)
    if "byte" in want:
        out["single_query_top10_via_byte_prefilter"] = guarded(lambda: extras_byte_prefilter(store, qs, lib, n, d))
    return out


def extras_byte_prefilter(store, qs, lib, n, d):
    """Secondary: the headline query (one query, top-10, host buffers in and out) through VectorStore.search_ids with the
    opt-in byte prefilter (csgpu_set_byte_prefilter: int8 shadow streamed as a filter with a proven bound + exact fp32
    rescoring in the same launch, csrc/scan_i8.cuh). Every timed query is compared bit for bit with the fp32 scan kernel."""
    k, nq = 10, 64
    want = [store.search_ids(qs[i], k) for i in range(nq)]       # fp32 scan kernel (the shadow does not exist yet)
    store.set_byte_prefilter(True)
    try:
        st0 = store.device_stats()
        got = [store.search_ids(qs[i], k) for i in range(nq)]
        first_us = None
        same = sum(int(np.array_equal(g[0], w[0]) and np.array_equal(g[1].view(np.uint32), w[1].view(np.uint32)))
                   for g, w in zip(got, want))
        t0 = time.perf_counter()
        reps = 200
        dev = []
        for i in range(reps):
            store.search_ids(qs[i % nq], k)
            dev.append(store.device_stats().last_search_us)
        dt = (time.perf_counter() - t0) / reps                        # includes the stats call: an upper bound
        st1 = store.device_stats()
        shadow = int(st1.byte_shadow_bytes)
        dev_ms = float(np.median(dev)) / 1e3
    finally:
        store.set_byte_prefilter(False)
    return {"workload": f"{n}x{d} fp32 index, one query per call, top-{k}, csgpu_search with the byte prefilter on (host buffers)",
            "ms_per_query": round(dt * 1e3, 4), "qps": round(1.0 / dt, 1), "device_ms": round(dev_ms, 4),
            "shadow_bytes": shadow, "shadow_GBps": round(shadow / dev_ms / 1e6, 1),
            "equivalent_fp32_GBps": round(n * d * 4 / dev_ms / 1e6, 1),
            "bit_identical_to_fp32_scan_kernel": f"{same}/{nq}",
            "int8_searches": int(st1.byte_searches - st0.byte_searches), "answered_by_fp32_scan_instead": int(st1.byte_fallbacks - st0.byte_fallbacks),
            "candidates_last_query": int(st1.byte_candidates), "fp32_rows_rescored_last_query": int(st1.byte_rescored)}


def extras_bf16(cs, _lib, lib, n, d, peaks_json):
    """BASELINE configs[2], "opt-in bf16 tcgen05 index": a bf16-stored index of the same rows, batch 1024 x top-100 through
    csgpu_search_batch (host buffers). Bound: tensor cores; recall@100 and |d_bf16 - d_fp32| against the exact fp32 index on
    the same rows (64 queries). Tolerance stated by SURVEY §8a A9: |d_bf16 - d_fp32| <= 2e-3."""
    b, k = 1024, 100
    st = cs.VectorStore.new(None, d, dtype="bf16")
    st.reserve(n)
    st.append_synthetic(SEED_CORPUS, 0, n)
    st.build_index()
    qs = _synth_queries(_lib, lib, st, b, d)
    smp = ClockSampler(0).start()
    dt, dev_ms, launches, res = _time_batches(st, lib, qs, k, 5)
    clk = smp.stop()
    dt10, dev10, _, _ = _time_batches(st, lib, qs, 10, 5)
    z = np.zeros((1, d), np.float32)                              # zero-norm query: distance 0.0, ascending id (contract, §8b)
    zi, zd, zn = st.search_batch_ids(z, 5)
    zero_ok = bool(zn[0] == 5 and np.array_equal(zi[0, :5], np.arange(5, dtype=np.uint32)) and not zd[0, :5].any())
    st.close()
    ref = cs.VectorStore.new(None, d)
    ref.reserve(n)
    ref.append_synthetic(SEED_CORPUS, 0, n)
    ref.set_tensor_prefilter(True)                                # exact fp32 answers at tensor speed (bit-identical to csgpu_search)
    ref.build_index()
    ri, rd, rn = ref.search_batch_ids(qs[:64], k)
    ref.close()
    recall, err = [], []
    for j in range(64):
        f = dict(zip(ri[j].tolist(), rd[j].tolist()))
        recall.append(len(set(ri[j].tolist()) & set(res[0][j].tolist())) / k)
        err += [abs(x - f[i]) for i, x in zip(res[0][j].tolist(), res[1][j].tolist()) if i in f]
    flop = 2.0 * n * d * b
    burst = float(peaks_json.get("bf16_tflops", 1662.2)); sust = float(peaks_json.get("bf16_tflops_sustained", 1344.8))
    return {"workload": f"{n}x{d} bf16 index (opt-in), batch {b} x top-{k}, csgpu_search_batch, host buffers",
            "ms_per_batch": round(dt * 1e3, 3), "device_ms": round(dev_ms, 3), "qps": round(b / dt, 1),
            "TFLOPs_device": round(flop / (dev_ms * 1e-3) / 1e12, 1), "bound": "tensor",
            "frac_of_burst_peak": round(flop / (dev_ms * 1e-3) / 1e12 / burst, 4), "frac_of_sustained_peak": round(flop / (dev_ms * 1e-3) / 1e12 / sust, 4),
            "peak_TFLOPs": {"burst": burst, "sustained": sust, "source": "MEASURED_PEAKS.json"},
            "top10": {"ms_per_batch": round(dt10 * 1e3, 3), "device_ms": round(dev10, 3)},
            "gpu_launches_per_batch": launches,
            "recall_at_100_vs_fp32_exact": round(float(np.mean(recall)), 4), "recall_min": float(np.min(recall)),
            "abs_dist_err_max": float(np.max(err)), "abs_dist_err_mean": float(np.mean(err)), "tolerance": 2e-3,
            "zero_norm_query_contract": zero_ok, "clocks": clk}


def extras_small(cs, _lib, lib, d, k):
    """GPU latency where a real codesearch index lives (src/constants.rs:93-95: ~100k chunks): end to end through csgpu_search
    with host buffers, beside cpu_baseline.configs0_100k_x_384."""
    out = {}
    for rows in (100_000, 200_000):
        st = cs.VectorStore.new(None, d)
        st.append_synthetic(SEED_CORPUS, 0, rows)
        st.build_index()
        qs = _synth_queries(_lib, lib, st, N_QUERIES, d)
        for i in range(10):
            st.search_ids(qs[i], k)
        dev = []
        t0 = time.perf_counter()
        for i in range(300):
            st.search_ids(qs[i % N_QUERIES], k)
        dt = (time.perf_counter() - t0) / 300
        for i in range(50):
            st.search_ids(qs[i % N_QUERIES], k)
            dev.append(st.device_stats().last_search_us)
        variants = {}
        for nv in (2, 4, 9, 16):      # query variants of one search (search/mod.rs:498-511), one csgpu_search_batch call
            for i in range(5):
                st.search_batch_ids(qs[:nv], k)
            t0 = time.perf_counter()
            for i in range(100):
                st.search_batch_ids(qs[:nv], k)
            variants[str(nv)] = round((time.perf_counter() - t0) / 100 * 1e3, 4)
        # the reference's hybrid default: <= 9 query variants x retrieval limit 200, best distance per chunk id, one list
        # (src/search/mod.rs:498-511, :513-590) = ONE csgpu_search_variants call
        for i in range(5):
            st.search_variants_ids(qs[:9], 200)
        t0 = time.perf_counter()
        for i in range(100):
            st.search_variants_ids(qs[i % 8: i % 8 + 9], 200)
        hybrid_ms = (time.perf_counter() - t0) / 100 * 1e3
        variants["9_variants_x_limit_200_deduplicated (csgpu_search_variants)"] = round(hybrid_ms, 4)
        out[f"{rows}_rows"] = {"e2e_ms_per_query": round(dt * 1e3, 4), "qps": round(1 / dt, 1),
                               "device_us": round(float(np.median(dev)), 1), "top_ids_query0": [int(i) for i in st.search_ids(qs[0], k)[0]],
                               "e2e_ms_per_batch_of_query_variants": variants}
        st.close()
    out["api"] = "VectorStore.search_ids -> csgpu_search, host pointers in and out, one query per call"
    # the same calls without Python in the loop: a plain-C program over include/csgpu.h (tools/bench_c_abi.c, built by build())
    exe = os.path.join(ROOT, "build", "bench_c_abi")
    if os.path.exists(exe):
        try:
            r = subprocess.run([exe, "100000", "200000"], capture_output=True, text=True, timeout=120)
            out["c_abi_latency_plain_c_host"] = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")] or {"error": r.stderr[-300:]}
        except Exception as e:  # noqa: BLE001
            out["c_abi_latency_plain_c_host"] = {"error": repr(e)[:200]}
    return out


def extras_in_process(cs, _lib, lib, world, n, d, k, args, q_host, fused_answers, e2e):
    """the SAME job through what a one-process host binds — csgpu_create(devices, N) + csgpu_search with
    host pointers — run by rank 0 alone while the other ranks idle (their stores are freed). N concurrent scan launches, the
    gather exchange fused into their tails, device 0's last CTA writes the global top-k into mapped host memory."""
    st = cs.VectorStore.new(None, d, devices=list(range(world)))
    st.reserve(n * world)
    st.append_synthetic(SEED_CORPUS, 0, n * world, 0)        # split evenly: shard g = rows [g n, (g+1) n), like the ranks
    st.build_index()
    per = list(st.device_stats().rows_per_device[:world])
    for i in range(max(5, args.warmup)):
        st.search_ids(q_host[i % N_QUERIES], k)
    l0 = lib.csgpu_kernel_launches()
    t0 = time.perf_counter()
    for i in range(args.steps):
        st.search_ids(q_host[i % N_QUERIES], k)
    dt = (time.perf_counter() - t0) / args.steps
    launches = (lib.csgpu_kernel_launches() - l0) / args.steps
    dev = []
    for i in range(20):
        st.search_ids(q_host[i % N_QUERIES], k)
        dev.append(st.device_stats().last_search_us)
    same = None
    if fused_answers:
        same = all(np.array_equal(g[0], w[0]) and np.array_equal(g[1].view(np.uint32), w[1].view(np.uint32))
                   for g, w in ((st.search_ids(q_host[j], k), fused_answers[j]) for j in range(len(fused_answers))))
    st.close()
    r = {"api": f"one process: csgpu_create(devices=[0..{world - 1}]) + csgpu_search (host pointers in and out)",
         "rows_per_device": per, "ms_per_query": round(dt * 1e3, 4), "qps": round(1 / dt, 2),
         "GBps": round(world * n * d * 4 / dt / 1e9, 1), "device_ms_root": round(float(np.median(dev)) / 1e3, 4),
         "gpu_launches_per_query": launches, "bit_identical_to_rank_per_gpu": bool(same) if same is not None else None}
    if e2e:
        r["rank_per_gpu_e2e_ms"] = e2e["ms_per_step"]
        r["ratio_to_rank_per_gpu_e2e"] = round(dt * 1e3 / e2e["ms_per_step"], 4)
    return r


def extras_c5(cs, _lib, lib, torch, dist, rank, world, local_rank, gather_obj, sync_all):
    """BASELINE configs[4], the vector leg: 5M rows per GPU x 768-d fp32, top-200, row-tag predicate
    ("lang in S and file_id < F") at densities 1.0 / 0.25 / 0.01 evaluated on the device before a row is read; N > 1: rows
    dealt in blocks of 1024 files round-robin, fused exchange. Device time per query (CUDA events, max over ranks)."""
    from codesearch_b200.sharded import ShardedSearcher
    n, d, k, reps = 5_000_000, 768, 200, 20
    N = n * world
    BLOCK = 37 * 1024
    n_blocks = (N + BLOCK - 1) // BLOCK
    my_blocks = [b for b in range(n_blocks) if b % world == rank]
    st = cs.VectorStore.new(None, d, devices=[local_rank])
    st.reserve(sum(min(BLOCK, N - b * BLOCK) for b in my_blocks))
    if world == 1:
        st.append_synthetic(SEED_CORPUS, 0, N, 0, tagged=True)
    else:
        for b in my_blocks:
            st.append_synthetic(SEED_CORPUS, b * BLOCK, min(BLOCK, N - b * BLOCK), 0, tagged=True)
    st.build_index()
    searcher = ShardedSearcher(st, exchange="fused")
    qs = _synth_queries(_lib, lib, st, 16, d)
    qd = torch.from_numpy(qs).cuda()
    n_files = (N + 36) // 37
    out = {"rows_total": N, "dim": d, "k": k, "n_gpus": world, "densities": {}}
    preds, answers = [], []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for dens in (1.0, 0.25, 0.01):
        if dens >= 1.0:
            pred = _lib.Predicate(0xFFFFFFFF, 0, 0xFFFFFFFF, 0, None, 0)
        else:
            n_lang = max(1, round(23 * min(1.0, dens * 4)))
            file_hi = max(1, int(dens / (n_lang / 23) * n_files))
            pred = _lib.Predicate((1 << n_lang) - 1, 0, file_hi - 1, 0, None, 0)
        for i in range(3):
            searcher.search_keys_device(qd[i], k, pred)
        sync_all()
        ev0.record()
        for i in range(reps):
            keys = searcher.search_keys_device(qd[i % 16], k, pred)
        ev1.record()
        sync_all()
        ms = max(gather_obj(ev0.elapsed_time(ev1) / reps))
        out["densities"][str(dens)] = {"device_ms": round(ms, 4), "dense_GBps": round(N * (d * 4 + 4) / ms / 1e6, 1)}
        preds.append(pred)
        answers.append(searcher.search_keys_device(qd[0], k, pred).cpu().numpy().copy())
    if world == 1:
        # the caller shape of configs[4]: HYBRID search under the mask — 9 query variants x limit 200, each searched over the
        # rows that pass the predicate, deduplicated by chunk id, as ONE csgpu_search_variants_tagged call (host buffers in and
        # out); checked against the dedup of the nine single tagged searches
        from codesearch_b200.tags import TagPredicate
        hyb = {}
        try:
            for dens, pred in zip((1.0, 0.25, 0.01), preds):
                tp = TagPredicate(lang_mask=int(pred.lang_mask), file_lo=int(pred.file_lo), file_hi=int(pred.file_hi))
                for i in range(2):
                    st.search_variants_tagged_ids(qs[:9], k, tp)
                t0 = time.perf_counter()
                for i in range(10):
                    gi, gd = st.search_variants_tagged_ids(qs[i % 7: i % 7 + 9], k, tp)
                ms = (time.perf_counter() - t0) / 10 * 1e3
                gi, gd = st.search_variants_tagged_ids(qs[:9], k, tp)
                best = {}
                for q in qs[:9]:
                    for i_, d_ in zip(*st.search_tagged_ids(q, k, tp)):
                        if int(i_) not in best or d_ < best[int(i_)]:
                            best[int(i_)] = d_
                want_ids = [i_ for i_, _ in sorted(best.items(), key=lambda t: (t[1], t[0]))[:k]]
                hyb[str(dens)] = {"e2e_ms": round(ms, 3), "equals_dedup_of_nine_tagged_searches": bool(gi.tolist() == want_ids)}
        except Exception as e:  # noqa: BLE001
            hyb = {"error": repr(e)[:200]}
        out["hybrid_9_variants_limit200_under_the_mask"] = hyb
    # the same three searches with the opt-in byte prefilter (round 2: the int8 kernel's FILT instantiation streams only row
    # groups the predicate allows; survivors are rescored in fp32 in the same launch) — every answer must stay bit-identical
    st.set_byte_prefilter(True)
    out["with_byte_prefilter"] = {"what": "csgpu_set_byte_prefilter on every rank: int8 shadow under the same predicate + exact fp32 rescoring"
                                          + ("" if world == 1 else ", then the exchange as its own launch"), "densities": {}}
    same = True
    for dens, pred, want in zip((1.0, 0.25, 0.01), preds, answers):
        for i in range(3):
            searcher.search_keys_device(qd[i], k, pred)
        sync_all()
        ev0.record()
        for i in range(reps):
            searcher.search_keys_device(qd[i % 16], k, pred)
        ev1.record()
        sync_all()
        ms = max(gather_obj(ev0.elapsed_time(ev1) / reps))
        got = searcher.search_keys_device(qd[0], k, pred).cpu().numpy()
        same = same and bool(np.array_equal(got, want))
        out["with_byte_prefilter"]["densities"][str(dens)] = {"device_ms": round(ms, 4)}
    sd = st.device_stats()
    out["with_byte_prefilter"]["bit_identical_to_fp32_filtered_scan"] = all(gather_obj(same))
    out["with_byte_prefilter"]["answered_by_fp32_scan_instead"] = int(sum(gather_obj(int(sd.byte_fallbacks))))
    st.close()
    return out


def extras_c4(cs, _lib, lib, torch, dist, rank, world, local_rank, args, gather_obj, sync_all, max_over_ranks):
    """BASELINE configs[3]: 50M rows per GPU (76.8 GB of fp32 rows per GPU; 100M / 200M / 400M rows at 2 / 4 / 8 GPUs),
    single-query top-10, fused exchange. Reports the step time, aggregate GB/s and — as the N = 1 reference measured on the
    same GPUs in the same run — each rank's LOCAL scan of its 50M rows without the exchange (efficiency = local / fused).
    Checks: planted neighbours (a corpus row from every shard as the query must come back first with distance ~0), the
    returned distances against the f64 oracle on the returned rows, every rank identical, fused == host merge of the ranks'
    local lists."""
    from codesearch_b200.sharded import ShardedSearcher, decode_keys
    n, d, k = args.c4_rows_per_gpu, args.dim, args.k
    steps = max(20, min(args.steps, 60))
    st = cs.VectorStore.new(None, d, devices=[local_rank])
    st.reserve(n)
    st.append_synthetic(SEED_CORPUS, rank * n, n, 0)
    st.build_index()
    searcher = ShardedSearcher(st, k_max=16, exchange="fused")
    qs = _synth_queries(_lib, lib, st, N_QUERIES, d)
    qd = torch.from_numpy(qs).cuda()
    stream = torch.cuda.current_stream().cuda_stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    # local scans (no exchange): the single-GPU number on this very GPU
    for i in range(3):
        _lib.check(lib.csgpu_search_keys_device(st.handle, qd[i].data_ptr(), k, searcher.local[:k].data_ptr(), stream))
    sync_all()
    ev[0].record()
    for i in range(steps):
        _lib.check(lib.csgpu_search_keys_device(st.handle, qd[i % N_QUERIES].data_ptr(), k, searcher.local[:k].data_ptr(), stream))
    ev[1].record()
    sync_all()
    local_ms = ev[0].elapsed_time(ev[1]) / steps
    local_all = gather_obj(local_ms)
    smp = ClockSampler(local_rank).start() if rank == 0 else None
    for i in range(3):
        searcher.search_keys_device(qd[i], k)
    sync_all()
    ev[2].record()
    for i in range(steps):
        searcher.search_keys_device(qd[i % N_QUERIES], k)
    ev[3].record()
    sync_all()
    fused_ms = max_over_ranks(ev[2].elapsed_time(ev[3]) / steps)
    clk = smp.stop() if smp else None
    sync_all()   # the ranks leave the sampler / gathers at different times: start the host-timed loop together
    t0 = time.perf_counter()
    for i in range(steps):
        searcher.search(qs[i % N_QUERIES], k)
    e2e_ms = max_over_ranks((time.perf_counter() - t0) / steps * 1e3)
    # ---- checks
    from oracle import oracle as O
    O.build()
    O.set_threads(max(1, len(os.sched_getaffinity(0)) // world))
    ok, notes = True, []
    planted = [g * n + off for g in range(world) for off in (0, n - 1)]          # first and last row of every shard
    for r_id in planted:
        q = O.synth_rows(SEED_CORPUS, r_id, 1, d)[0]
        ids, dd = searcher.search(q, k) if world > 1 else st.search_ids(q, k)
        if not (len(ids) == k and int(ids[0]) == r_id and float(dd[0]) <= 1e-6):
            ok = False; notes.append(f"planted row {r_id} not first: {ids[:3].tolist()} {dd[:3].tolist()}")
    for j in range(2):
        ids, dd = searcher.search(qs[j], k) if world > 1 else st.search_ids(qs[j], k)
        for i_, d_ in zip(ids.tolist(), dd.tolist()):                              # returned rows re-scored by the f64 oracle
            row = O.synth_rows(SEED_CORPUS, i_, 1, d)
            oi, od, o64 = O.search(row, qs[j], 1)
            if abs(float(o64[0]) - d_) > SCORE_TOL:
                ok = False; notes.append(f"query {j} id {i_}: distance {d_} vs oracle {float(o64[0])}")
        if not all((dd[i] < dd[i + 1]) or (dd[i] == dd[i + 1] and ids[i] < ids[i + 1]) for i in range(k - 1)):
            ok = False; notes.append(f"query {j}: not ascending")
        # fused == merge of the ranks' local lists (host)
        _lib.check(lib.csgpu_search_keys_device(st.handle, qd[j].data_ptr(), k, searcher.local[:k].data_ptr(), stream))
        torch.cuda.synchronize()
        loc = gather_obj(searcher.local[:k].cpu().numpy().view(np.uint64).tolist())
        merged = np.array(sorted(x for l in loc for x in l)[:k], dtype=np.uint64)
        mi, md = decode_keys(merged.view(np.int64))
        if not (np.array_equal(mi, ids) and np.array_equal(md.view(np.uint32), dd.view(np.uint32))):
            ok = False; notes.append(f"query {j}: fused result differs from the host merge of the local lists")
        every = gather_obj((ids.tolist(), dd.view(np.uint32).tolist()))
        if not all(e == every[0] for e in every):
            ok = False; notes.append(f"query {j}: ranks disagree")
    status = 0
    if world > 1:
        t = ctypes.c_uint32(0)
        _lib.check(lib.csgpu_exchange_status(st.handle, ctypes.byref(t)))
        status = int(t.value)
    oks = gather_obj((ok and status == 0, notes))
    st.close()
    _, peak, _ = peaks()
    tot = world * n * d * 4
    return {"workload": f"{world * n}x{d} fp32 corpus, {n} rows per GPU (BASELINE configs[3]), single-query top-{k}, fused exchange",
            "steps": steps, "ms_per_step": round(fused_ms, 4), "GBps_aggregate": round(tot / fused_ms / 1e6, 1),
            "GBps_per_gpu": round(tot / world / fused_ms / 1e6, 1), "frac_of_measured_hbm_peak": round(tot / world / fused_ms / 1e6 / peak, 4),
            "qps": round(1e3 / fused_ms, 2), "e2e_ms_per_step": round(e2e_ms, 4),
            "local_scan_ms_per_rank": [round(x, 4) for x in local_all],
            "efficiency_vs_local_scan": round(max(local_all) / fused_ms, 4) if world > 1 else 1.0,
            "efficiency_vs_mean_local_scan": round(float(np.mean(local_all)) / fused_ms, 4) if world > 1 else 1.0,
            "checks": {"ok": all(o[0] for o in oks), "planted_rows": planted, "notes": [x for o in oks for x in o[1]][:6],
                       "what": "planted neighbours first with d<=1e-6; returned rows re-scored by the f64 oracle within 1e-5; ascending; "
                               "fused == host merge of local lists; ranks identical; exchange status clean"},
            "clocks": clk}


if __name__ == "__main__":
    main()
