"""N>1 host logic on CPU: world_size-2 (and 3) gloo process groups exercise the shard partition and
the path's single exchange step (all-gather of k sorted keys). The per-shard top-k comes from the
oracle here (no GPU in this container); the GPU tests cover the same flow with the CUDA kernels
(tests/test_gpu_parity.py::test_device_entry_points_and_merge). Property checked: the global top-k
is contained in the union of per-shard top-k lists, and merging by (distance, id) key reproduces the
single-index answer exactly — which is what makes row-sharding exact."""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _okey(d32: np.ndarray) -> np.ndarray:
    u = d32.astype(np.float32).view(np.uint32).astype(np.uint64)
    neg = (u >> np.uint64(31)) != 0
    return np.where(neg, u ^ np.uint64(0xFFFFFFFF), u ^ np.uint64(0x80000000))


def encode_keys(ids, dist32, k):
    keys = np.full(k, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
    keys[: len(ids)] = (_okey(np.asarray(dist32)) << np.uint64(32)) | np.asarray(ids, dtype=np.uint64)
    return keys


def _worker(rank, world, port, n, d, k, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from codesearch_b200 import _lib
        from codesearch_b200.sharded import allgather_keys, shard_range
        O.set_threads(1)
        first, cnt = shard_range(n, rank, world)
        rows = O.synth_rows(1234, first, cnt, d)
        ids = np.arange(first, first + cnt, dtype=np.uint32)
        plant = shard_range(n, 0, world)[1] > 5
        if rank == 0 and plant:   # duplicates across shards must tie-break by id after the merge
            rows[5] = O.synth_rows(1234, n - 1, 1, d)[0]
        q = O.synth_rows(4321, 3, 1, d)[0]
        li, ld, _ = O.search(rows, q, k, ids=ids)
        local = torch.from_numpy(encode_keys(li, ld, k).view(np.int64))
        gathered = allgather_keys(local, world)
        assert gathered.shape == (world * k,)
        merged = np.sort(gathered.numpy().view(np.uint64))[:k]
        lib = _lib.load()
        oi = np.zeros(k, np.uint32); od = np.zeros(k, np.float32); on = ctypes.c_uint32()
        lib.csgpu_decode_keys(merged.ctypes.data_as(_lib._u64p), k, oi.ctypes.data_as(_lib._u32p),
                              od.ctypes.data_as(_lib._f32p), ctypes.byref(on))
        # every rank must hold the identical global answer
        all_rows = O.synth_rows(1234, 0, n, d)
        if plant:
            all_rows[5] = all_rows[n - 1]
        gi, gd, _ = O.search(all_rows, q, k)
        ok = on.value == min(k, n) and np.array_equal(oi[: on.value], gi) and np.array_equal(od[: on.value], gd)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,n,k", [(2, 4001, 10), (3, 1000, 64), (2, 7, 10)])
def test_sharded_exchange_gloo(world, n, k):
    from oracle import oracle as O
    O.build()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, 64, k, ret), nprocs=world, join=True)
    assert dict(ret) == {r: True for r in range(world)}


def test_shard_range_partitions_exactly():
    sys.path.insert(0, ROOT)
    from codesearch_b200.sharded import shard_range
    for n in (0, 1, 7, 10_000_000, 400_000_000):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _worker_tagged(rank, world, port, n, d, k, ret):
    """BASELINE configs[4] host logic: rows dealt to the ranks in blocks of whole files round-robin (tools/bench_hybrid.py),
    a language + file-range predicate applied per shard before the top-k, one all-gather, merge by key."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from codesearch_b200.sharded import allgather_keys
        O.set_threads(1)
        block = 37 * 8
        n_blocks = (n + block - 1) // block
        mine = [b for b in range(n_blocks) if b % world == rank]
        ids = np.concatenate([np.arange(b * block, min(n, (b + 1) * block), dtype=np.uint32) for b in mine])
        rows = np.concatenate([O.synth_rows(1234, b * block, min(block, n - b * block), d) for b in mine])
        tags = np.concatenate([O.synth_tags(b * block, min(block, n - b * block)) for b in mine])
        lang_mask, file_lo, file_hi = (1 << 0) | (1 << 3) | (1 << 7) | (1 << 11) | (1 << 20), 5, (n // 37) // 2
        ok_rows = O.tag_predicate_mask(tags, lang_mask, file_lo, file_hi)
        q = O.synth_rows(4321, 1, 1, d)[0]
        li, ld, _ = O.search(rows[ok_rows], q, k, ids=ids[ok_rows]) if ok_rows.any() else (np.zeros(0, np.uint32), np.zeros(0, np.float32), None)
        local = torch.from_numpy(encode_keys(li, ld, k).view(np.int64))
        merged = np.sort(allgather_keys(local, world).numpy().view(np.uint64))[:k]
        all_rows = O.synth_rows(1234, 0, n, d)
        all_ok = O.tag_predicate_mask(O.synth_tags(0, n), lang_mask, file_lo, file_hi)
        gi, gd, _ = O.search(all_rows[all_ok], q, k, ids=np.arange(n, dtype=np.uint32)[all_ok])
        want = encode_keys(gi, gd, k)
        # balanced: a contiguous file range loads every rank (the point of dealing whole-file blocks round-robin)
        share = torch.tensor([float(ok_rows.sum())])
        shares = [torch.zeros(1) for _ in range(world)]
        dist.all_gather(shares, share)
        tot = sum(s.item() for s in shares)
        ret[rank] = bool(np.array_equal(merged, want) and all(s.item() > 0.5 * tot / world for s in shares))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,k", [(2, 6000, 200), (3, 5000, 10)])
def test_sharded_tagged_filter_gloo(world, n, k):
    from oracle import oracle as O
    O.build()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_tagged, args=(world, _free_port(), n, 64, k, ret), nprocs=world, join=True)
    assert dict(ret) == {r: True for r in range(world)}
