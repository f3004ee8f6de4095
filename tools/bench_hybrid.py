#!/usr/bin/env python
"""BASELINE configs[4]: hybrid search's vector leg — 768-d top-200 with a file/language filter,
row-sharded over N GPUs (one rank per GPU, fused exchange), fed into a host RRF stand-in.

  python tools/bench_hybrid.py                                  # 1 GPU, 5M x 768
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_hybrid.py

Per GPU: rows-per-gpu x 768 fp32 synthetic rows with synthetic tags (file_id = row / 37, lang_id = fmix32(file_id) % 23).
Rows are dealt to the ranks in blocks of 1024 files (37,888 rows) round-robin, so that a path-prefix-like filter (a
contiguous file range) loads every GPU equally instead of landing on one contiguous row shard.
The filter is the device-side row-tag predicate (csgpu_search_tagged_keys_device): "lang in S and file_id in range",
chosen to hit the target densities. Masked rows are never read, so two byte counts are reported:
dense (N x D x 4 + 4 B tag per row) and effective (passing rows x D x 4 + 4 B tag per row).
The RRF stand-in (rerank/mod.rs:57-66: score = 1/(k + rank + 1), k = 60) fuses the vector list with a synthetic
second list on the host, inside the e2e timing. One JSON line per density, rank 0.
"""
import argparse, ctypes, json, os, statistics, sys, time
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import codesearch_b200 as cs
from codesearch_b200 import _lib
from codesearch_b200.sharded import ShardedSearcher, decode_keys

p = argparse.ArgumentParser()
p.add_argument("--rows-per-gpu", type=int, default=5_000_000)
p.add_argument("--dim", type=int, default=768)
p.add_argument("--k", type=int, default=200)
p.add_argument("--densities", default="1.0,0.25,0.01")
p.add_argument("--reps", type=int, default=40)
p.add_argument("--exchange", default="fused")
args = p.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = _lib.load()
n, d, k = args.rows_per_gpu, args.dim, args.k
N = n * world
BLOCK = 37 * 1024                                     # rows per block = 1024 whole files
n_blocks = (N + BLOCK - 1) // BLOCK
my_blocks = [b for b in range(n_blocks) if b % world == rank]
st = cs.VectorStore.new(None, d, devices=[local])
st.reserve(sum(min(BLOCK, N - b * BLOCK) for b in my_blocks))
for b in my_blocks:
    st.append_synthetic(1234, b * BLOCK, min(BLOCK, N - b * BLOCK), 0, tagged=True)   # chunk id = global row index
st.build_index()
searcher = ShardedSearcher(st, exchange=args.exchange)
qs = np.empty((64, d), np.float32)
_lib.check(lib.csgpu_synth_rows_host(st.handle, 4321, 0, 64, qs.ctypes.data_as(_lib._f32p)))
n_files = (N + 36) // 37


def fmix32(h):
    h = h.astype(np.uint32)
    h ^= h >> np.uint32(16); h *= np.uint32(0x85EBCA6B); h ^= h >> np.uint32(13); h *= np.uint32(0xC2B2AE35); h ^= h >> np.uint32(16)
    return h


def rrf(vec_ids, other_ids, k_rrf=60.0):
    score = {}
    for r, i in enumerate(vec_ids.tolist()):
        score[i] = score.get(i, 0.0) + 1.0 / (k_rrf + r + 1.0)
    for r, i in enumerate(other_ids.tolist()):
        score[i] = score.get(i, 0.0) + 1.0 / (k_rrf + r + 1.0)
    return sorted(score.items(), key=lambda t: -t[1])[:k]


ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
rng = np.random.default_rng(9)
for dens in (float(x) for x in args.densities.split(",")):
    if dens >= 1.0:
        pred = _lib.Predicate(0xFFFFFFFF, 0, 0xFFFFFFFF, 0, None, 0)
        lang_mask, file_hi = 0xFFFFFFFF, n_files
    else:
        n_lang = max(1, round(23 * min(1.0, dens * 4)))       # languages allowed: S = {0 .. n_lang-1}
        frac_files = dens / (n_lang / 23)                     # leading fraction of files allowed
        lang_mask, file_hi = (1 << n_lang) - 1, max(1, int(frac_files * n_files))
        pred = _lib.Predicate(lang_mask, 0, file_hi - 1, 0, None, 0)
    # passing rows of this rank (host, numpy — not timed)
    passing = 0
    for b in my_blocks:
        r0, r1 = b * BLOCK, min(N, (b + 1) * BLOCK)
        files = np.arange(r0 // 37, (r1 - 1) // 37 + 1, dtype=np.uint64)
        f_ok = ((np.uint64(lang_mask) >> (fmix32(files.astype(np.uint32)) % np.uint32(23)).astype(np.uint64)) & np.uint64(1)).astype(bool) & (files < file_hi)
        lo = np.maximum(files * 37, r0); hi = np.minimum(files * 37 + 37, r1)
        passing += int(((hi - lo) * f_ok).sum())
    dev_ms, e2e_ms = [], []
    for i in range(args.reps + 3):
        q = qs[i % 64]
        other = rng.integers(0, N, size=k).astype(np.uint32)      # synthetic BM25 list
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        searcher.q_pin[:d].copy_(torch.from_numpy(q))
        searcher.q_dev.copy_(searcher.q_pin, non_blocking=True)
        ev0.record()
        keys = searcher.search_keys_device(searcher.q_dev, k, pred)
        ev1.record()
        searcher.out_pin[:k].copy_(keys, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        ids, dd = decode_keys(searcher.out_pin[:k].numpy())
        fused = rrf(ids, other)
        t1 = time.perf_counter()
        if i >= 3:
            dev_ms.append(ev0.elapsed_time(ev1)); e2e_ms.append((t1 - t0) * 1e3)
    dm, em = statistics.median(dev_ms), statistics.median(e2e_ms)
    stats = torch.tensor([dm, em, float(passing)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dm, em, passing_all = mx[0].item(), mx[1].item(), int(sm[2].item())
    else:
        passing_all = passing
    if rank == 0:
        dense = N * (d * 4 + 4)
        eff = passing_all * d * 4 + N * 4
        print(json.dumps({"config": "C5 hybrid vector leg", "n_gpus": world, "rows_total": N, "dim": d, "k": k,
                          "exchange": args.exchange if world > 1 else "none",
                          "filter": "none" if dens >= 1.0 else f"lang<{bin(lang_mask).count('1')} & file<{file_hi}",
                          "density": round(passing_all / N, 5), "device_ms": round(dm, 4), "e2e_ms_with_rrf": round(em, 4),
                          "dense_GBps": round(dense / dm / 1e6, 1), "effective_GBps": round(eff / dm / 1e6, 1),
                          "qps_e2e": round(1e3 / em, 1), "results": len(ids), "fused_len": len(fused)}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
