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
os.environ.setdefault("CSGPU_I8_MIN_ROWS", "4096")   # read once by the library: lets small shards take the int8 route

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


def test_host_pointer_exchange_search_three_ranks_one_gpu(cs, oracle):
    """csgpu_search_exchange: the exchange search with HOST pointers (what csgpu_search is to a single index). It blocks until
    the global top-k is back, so the three 'ranks' call it from three host threads, as three processes would; every rank must
    return the unsharded index's answer bit for bit — with the fp32 scan and with the byte prefilter on one of the ranks."""
    import threading
    from codesearch_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(35)
    n, d, W = 36000, 384, 3
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[7] = 0.0
    bounds = [0, 10000, 26000, n]
    stores = []
    for r, (a, b) in enumerate(zip(bounds, bounds[1:])):
        st = cs.VectorStore.new(None, d)
        st.append_rows(rows[a:b], np.arange(a, b, dtype=np.uint32))
        if r == 1:
            st.set_byte_prefilter(True)                          # mixed routes: int8 kernel + exchange_keys_kernel on this rank
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
    qs = rng.standard_normal((6, d)).astype(np.float32)
    ks = [10, 100, 10, 1, 256, 32]
    got = [[None] * len(ks) for _ in range(W)]
    errs = []

    def rank_main(r):
        try:
            for j, k in enumerate(ks):
                oi = np.empty(k, np.uint32); od = np.empty(k, np.float32); on = ctypes.c_uint32(0)
                _lib.check(lib.csgpu_search_exchange(stores[r].handle, qs[j].ctypes.data_as(_lib._f32p), d, k,
                                                     oi.ctypes.data_as(_lib._u32p), od.ctypes.data_as(_lib._f32p), ctypes.byref(on)))
                got[r][j] = (oi[: on.value].copy(), od[: on.value].copy())
        except Exception as e:  # noqa: BLE001
            errs.append((r, repr(e)))

    ths = [threading.Thread(target=rank_main, args=(r,)) for r in range(W)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert not errs, errs
    for j, k in enumerate(ks):
        wi, wd = whole.search_ids(qs[j], k)
        for r in range(W):
            assert np.array_equal(got[r][j][0], wi) and np.array_equal(got[r][j][1].view(np.uint32), wd.view(np.uint32)), (j, r)
    # guard clauses verbatim through this entry point too
    oi = np.empty(4, np.uint32); od = np.empty(4, np.float32); on = ctypes.c_uint32(0)
    rc = lib.csgpu_search_exchange(stores[0].handle, qs[0].ctypes.data_as(_lib._f32p), d - 1, 4, oi.ctypes.data_as(_lib._u32p),
                                   od.ctypes.data_as(_lib._f32p), ctypes.byref(on))
    assert rc == _lib.ERR_DIM and _lib.last_error() == f"Query embedding dimension mismatch: expected {d}, got {d - 1}"
    for st in stores:
        lib.csgpu_exchange_destroy(st.handle)


def test_exchange_missing_rank_is_an_error_not_garbage(cs):
    """SURVEY.md §5 "FFI must return errors (never abort)": only 2 of 3 ranks search. Their in-kernel wait gives up after
    the configured bound, the status word (pinned host memory) is raised, and every later exchange search on those ranks
    returns CSGPU_ERR_NCCL instead of undefined keys — until the exchange is created and connected again."""
    import torch
    from codesearch_b200 import _lib
    from codesearch_b200.sharded import decode_keys
    lib = _lib.load()
    rng = np.random.default_rng(32)
    n, d, W, k = 6000, 64, 3, 10
    rows = rng.standard_normal((n, d)).astype(np.float32)
    stores = []
    for r in range(W):
        st = cs.VectorStore.new(None, d)
        st.append_rows(rows[r * 2000:(r + 1) * 2000], np.arange(r * 2000, (r + 1) * 2000, dtype=np.uint32))
        st.build_index()
        _lib.check(lib.csgpu_exchange_set_timeout_ms(st.handle, 100))
        stores.append(st)

    def connect():
        for r, st in enumerate(stores):
            h = (ctypes.c_ubyte * 64)()
            _lib.check(lib.csgpu_exchange_create(st.handle, W, r, h))
        peers = (ctypes.c_void_p * W)(*[st.handle for st in stores])
        for st in stores:
            _lib.check(lib.csgpu_exchange_connect_local(st.handle, peers))

    connect()
    streams = [torch.cuda.Stream() for _ in range(W)]
    qd = torch.from_numpy(rng.standard_normal(d).astype(np.float32)).cuda()
    outs = [torch.empty(k, dtype=torch.int64, device="cuda") for _ in range(W)]
    torch.cuda.synchronize()
    for r in (0, 1):                                             # rank 2 never launches
        _lib.check(lib.csgpu_search_keys_exchange_device(stores[r].handle, qd.data_ptr(), k, outs[r].data_ptr(), streams[r].cuda_stream))
    torch.cuda.synchronize()                                    # returns after ~0.1 s, not 4 s and not never
    for r in (0, 1):
        t = ctypes.c_uint32(0)
        _lib.check(lib.csgpu_exchange_status(stores[r].handle, ctypes.byref(t)))
        assert t.value == 1
        rc = lib.csgpu_search_keys_exchange_device(stores[r].handle, qd.data_ptr(), k, outs[r].data_ptr(), streams[r].cuda_stream)
        assert rc == _lib.ERR_NCCL and "timed out" in _lib.last_error()
    t = ctypes.c_uint32(9)
    _lib.check(lib.csgpu_exchange_status(stores[2].handle, ctypes.byref(t)))
    assert t.value == 0
    connect()                                                   # a fresh exchange heals all three
    whole = cs.VectorStore.new(None, d)
    whole.append_rows(rows, np.arange(n, dtype=np.uint32))
    whole.build_index()
    for r in range(W):
        _lib.check(lib.csgpu_search_keys_exchange_device(stores[r].handle, qd.data_ptr(), k, outs[r].data_ptr(), streams[r].cuda_stream))
    torch.cuda.synchronize()
    gi, gd = whole.search_ids(qd.cpu().numpy(), k)
    for r in range(W):
        ids, dist = decode_keys(outs[r].cpu().numpy())
        assert np.array_equal(ids, gi) and np.array_equal(dist, gd)
    # skew diagnostic: one query since the reconnect, world entries, all far below the bound
    ns = np.zeros(8 * W, dtype=np.uint64)
    nq = ctypes.c_uint32(0)
    _lib.check(lib.csgpu_exchange_wait_stats(stores[0].handle, ns.ctypes.data_as(_lib._u64p), 8, ctypes.byref(nq)))
    assert nq.value == 1 and int(ns[:W].max()) < 100_000_000
    for st in stores:
        lib.csgpu_exchange_destroy(st.handle)


def test_device_entry_points_two_streams_share_no_scratch(cs):
    """Round-1 advisor finding: the device entry points returned their context to the shared pool while its scratch was in
    flight. They now own a dedicated context and order successive calls with an event, so alternating streams without any
    synchronisation in between — and host-pointer searches from another thread meanwhile — stay bit-identical."""
    import threading
    import torch
    from codesearch_b200 import _lib
    from codesearch_b200.sharded import decode_keys
    lib = _lib.load()
    rng = np.random.default_rng(33)
    n, d, k = 300_000, 128, 10
    st = cs.VectorStore.new(None, d)
    st.append_synthetic(99, 0, n, 0)
    st.build_index()
    qs = rng.standard_normal((16, d)).astype(np.float32)
    want = [st.search_ids(q, k) for q in qs]
    qd = torch.from_numpy(qs).cuda()
    outs = torch.empty((64, k), dtype=torch.int64, device="cuda")
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    stop = threading.Event()
    bad = []

    def host_searches():
        i = 0
        while not stop.is_set():
            g = st.search_ids(qs[i % 16], k)
            if not (np.array_equal(g[0], want[i % 16][0]) and np.array_equal(g[1], want[i % 16][1])):
                bad.append(i)
            i += 1

    th = threading.Thread(target=host_searches)
    th.start()
    for i in range(64):
        _lib.check(lib.csgpu_search_keys_device(st.handle, qd[i % 16].data_ptr(), k, outs[i].data_ptr(), streams[i & 1].cuda_stream))
    torch.cuda.synchronize()
    stop.set()
    th.join()
    assert not bad
    got = outs.cpu().numpy()
    for i in range(64):
        ids, dist = decode_keys(got[i])
        assert np.array_equal(ids, want[i % 16][0]) and np.array_equal(dist, want[i % 16][1]), i


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
