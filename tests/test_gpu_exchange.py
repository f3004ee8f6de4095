"""Fused cross-GPU exchange (csgpu_exchange_*, scan.cuh exchange_and_merge): the scan kernel's last CTA writes the
rank's keys into every peer's slot block, waits for the peers' flags and merges — one kernel per query per rank.

On the one-GPU test box the "ranks" are three indexes on the same device, wired with csgpu_exchange_connect_local
and launched on three streams (the kernels must be co-resident: each one's tail spins until the others publish).
With >= 2 GPUs the real thing runs: one process per GPU, cudaIpc-mapped peer memory, compared with the NCCL path.
"""
import ctypes
import os

import numpy as np
import pytest

from parity import check_topk

pytestmark = pytest.mark.gpu

MARGIN = 8


@pytest.fixture(scope="module")
def cs():
    import codesearch_b200 as m
    m.load_library()
    return m


def test_fused_exchange_three_ranks_one_gpu(cs, oracle):
    import torch
    from codesearch_b200 import _lib
    from codesearch_b200.sharded import decode_keys
    lib = _lib.load()
    rng = np.random.default_rng(31)
    n, d = 40000, 384
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[5] = 0.0                                               # zero-norm row lives on rank 0
    bounds = [0, 9000, 25000, n]
    W = 3
    stores = []
    for a, b in zip(bounds, bounds[1:]):
        st = cs.VectorStore.new(None, d)
        st.append_rows(rows[a:b], np.arange(a, b, dtype=np.uint32))
        st.build_index()
        stores.append(st)
    for r, st in enumerate(stores):
        h = (ctypes.c_ubyte * 64)()
        _lib.check(lib.csgpu_exchange_create(st.handle, W, r, h))
    peers = (ctypes.c_void_p * W)(*[st.handle for st in stores])
    for st in stores:
        _lib.check(lib.csgpu_exchange_connect_local(st.handle, peers))
    whole = cs.VectorStore.new(None, d)
    whole.append_rows(rows, np.arange(n, dtype=np.uint32))
    whole.build_index()

    streams = [torch.cuda.Stream() for _ in range(W)]
    qs = rng.standard_normal((7, d)).astype(np.float32)
    for qi, k in enumerate([10, 10, 32, 100, 10, 1000, 1]):     # consecutive queries flip the slot parity
        qd = torch.from_numpy(qs[qi]).cuda()
        outs = [torch.empty(k, dtype=torch.int64, device="cuda") for _ in range(W)]
        torch.cuda.synchronize()
        for r, st in enumerate(stores):
            _lib.check(lib.csgpu_search_keys_exchange_device(st.handle, qd.data_ptr(), k, outs[r].data_ptr(),
                                                             streams[r].cuda_stream))
        torch.cuda.synchronize()
        gi, gd = whole.search_ids(qs[qi], k)
        oi, od, o64 = oracle.np_search(rows, qs[qi], k + MARGIN)
        for r in range(W):
            ids, dist = decode_keys(outs[r].cpu().numpy())
            # every rank holds the same global top-k, bit-identical to the unsharded index
            assert np.array_equal(ids, gi) and np.array_equal(dist, gd), (qi, r)
        check_topk(gi, gd, oi, od, o64, min(k, n))
    for st in stores:
        t = ctypes.c_uint32(7)
        _lib.check(lib.csgpu_exchange_status(st.handle, ctypes.byref(t)))
        assert t.value == 0
        lib.csgpu_exchange_destroy(st.handle)


def _rank_main(rank, world, port, n, d, ks, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import codesearch_b200 as cs
    from codesearch_b200.sharded import ShardedSearcher, shard_range
    first, cnt = shard_range(n, rank, world)
    st = cs.VectorStore.new(None, d, devices=[rank])
    st.append_synthetic(1234, first, cnt, 0)
    st.build_index()
    fused = ShardedSearcher(st, exchange="fused")
    nccl = ShardedSearcher(st, exchange="nccl")
    from oracle import oracle as O
    qs = O.synth_rows(4321, 0, len(ks), d)
    out = []
    for q, k in zip(qs, ks):
        fi, fd = fused.search(q, k)
        ni, nd = nccl.search(q, k)
        assert np.array_equal(fi, ni) and np.array_equal(fd, nd)
        out.append((fi.tolist(), fd.tolist()))
    if rank == 0:
        ret.put(out)
    dist.barrier()
    dist.destroy_process_group()


def _rank_batch(rank, world, port, n, d, b, k, mode, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import codesearch_b200 as cs
    from codesearch_b200.sharded import ShardedSearcher, shard_range
    first, cnt = shard_range(n, rank, world)
    st = cs.VectorStore.new(None, d, devices=[rank])
    st.append_synthetic(1234, first, cnt, 0)
    if mode == "prefilter":
        st.set_tensor_prefilter(True)
    st.build_index()
    s = ShardedSearcher(st, exchange="nccl")
    from oracle import oracle as O
    qs = O.synth_rows(4321, 0, b, d)
    ids, dd, nn = s.search_batch(qs, k)
    if rank == 0:
        ret.put((ids.tolist(), dd.tolist(), nn.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode,b", [("scan", 9), ("prefilter", 70)])
def test_sharded_batch_two_gpus(cs, oracle, mode, b):
    """Rank-per-GPU batches: local csgpu_search_batch, one all-gather of [b, k] keys, batched merge kernel."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    n, d, k = 200_000, 384, 20
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_rank_batch, args=(r, 2, 29537, n, d, b, k, mode, ret)) for r in range(2)]
    [p.start() for p in procs]
    ids, dd, nn = ret.get(timeout=300)
    [p.join(timeout=120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    rows = oracle.synth_rows(1234, 0, n, d)
    qs = oracle.synth_rows(4321, 0, b, d)
    for j in (0, b // 2, b - 1):
        oi, od, o64 = oracle.search(rows, qs[j], k + MARGIN)
        assert nn[j] == k
        check_topk(np.array(ids[j], np.uint32), np.array(dd[j], np.float32), oi, od, o64, k)


def test_fused_exchange_two_gpus(cs, oracle):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    n, d, ks = 300_000, 384, [10, 100, 10, 33]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, 29533, n, d, ks, ret)) for r in range(2)]
    [p.start() for p in procs]
    got = ret.get(timeout=300)
    [p.join(timeout=120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    rows = oracle.synth_rows(1234, 0, n, d)
    qs = oracle.synth_rows(4321, 0, len(ks), d)
    for (ids, dist), q, k in zip(got, qs, ks):
        oi, od, o64 = oracle.search(rows, q, k + MARGIN)
        check_topk(np.array(ids, np.uint32), np.array(dist, np.float32), oi, od, o64, k)
