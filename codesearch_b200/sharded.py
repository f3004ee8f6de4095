"""Rank-per-GPU sharded search: one process per GPU under torch.distributed (NCCL over NVLink).

Each rank owns a contiguous row shard in its own single-device csgpu index. A query is scanned by
every rank (local fused scan + top-k, one kernel), the k sortable keys per rank are exchanged with
ONE all-gather (k * 8 B per rank — latency-bound), and every rank runs the same k-way merge kernel,
so all ranks end with the identical global top-k, ties broken by chunk id.

torch is plumbing here (device buffers, the current stream, the process group); the scan and the
merge are csgpu kernels reached through the C ABI's device entry points.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .store import VectorStore


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous row range [first, first+count) owned by `rank` (GPU g owns rows
    [g*N/G, (g+1)*N/G))."""
    first = n_total * rank // world
    return first, n_total * (rank + 1) // world - first


def allgather_keys(local: torch.Tensor, world: int, group=None, out: torch.Tensor | None = None) -> torch.Tensor:
    """The path's one exchange step: every rank contributes its k sorted keys (int64 bit pattern of
    the u64 key), every rank receives [world, k]. Backend-agnostic (NCCL on GPUs, gloo in CPU tests)."""
    k = local.numel()
    if out is None:
        out = torch.empty(world * k, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    return out


def connect_fused_exchange(store: VectorStore, rank: int, world: int, group=None) -> None:
    """One-time setup of the fused exchange (include/csgpu.h): every rank creates its slot block, the 64-byte
    cudaIpc handles are all-gathered on the host (the only use of the process group on this path), and each
    rank maps its peers' blocks. After this, a search is ONE kernel per rank with peer stores over NVLink."""
    lib = _lib.load()
    mine = (ctypes.c_ubyte * _lib.EXCHANGE_HANDLE_BYTES)()
    _lib.check(lib.csgpu_exchange_create(store.handle, world, rank, mine))
    handles = [None] * world
    dist.all_gather_object(handles, bytes(mine), group=group)
    blob = (ctypes.c_ubyte * (_lib.EXCHANGE_HANDLE_BYTES * world)).from_buffer_copy(b"".join(handles))
    _lib.check(lib.csgpu_exchange_connect(store.handle, blob))
    dist.barrier(group=group)   # nobody searches before every rank has mapped every block


class ShardedSearcher:
    """exchange="nccl": local scan kernel -> NCCL all-gather of k keys -> merge kernel (3 launches per query).
    exchange="fused": the scan kernel writes its keys into every peer's HBM and merges in its own tail
    (1 launch per query, no collective library on the data path)."""

    def __init__(self, store: VectorStore, k_max: int = _lib.MAX_K, group=None, exchange: str = "nccl"):
        self.store = store
        self.lib = _lib.load()
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.exchange = exchange if self.world > 1 else "none"
        assert exchange in ("nccl", "fused")
        if self.exchange == "fused":
            connect_fused_exchange(store, dist.get_rank(group), self.world, group)
        self.dev = torch.device("cuda", torch.cuda.current_device())
        d = store.dimensions
        self.d_pad = (d + 3) // 4 * 4
        self.q_pin = torch.zeros(self.d_pad, dtype=torch.float32).pin_memory()
        self.q_dev = torch.zeros(self.d_pad, dtype=torch.float32, device=self.dev)
        self.local = torch.empty(k_max, dtype=torch.int64, device=self.dev)
        self.gathered = torch.empty(self.world * k_max, dtype=torch.int64, device=self.dev)
        self.out = torch.empty(k_max, dtype=torch.int64, device=self.dev)
        self.out_pin = torch.empty(k_max, dtype=torch.int64).pin_memory()

    # -- device-resident: q_dev is a [d_pad] float32 CUDA tensor; returns a view of k keys (int64 bits)
    def search_keys_device(self, q_dev: torch.Tensor, k: int, pred: "_lib.Predicate | None" = None) -> torch.Tensor:
        """pred: row-tag predicate (its file_bitmap, if any, is a DEVICE pointer) -> csgpu_search_tagged_keys_device."""
        stream = torch.cuda.current_stream().cuda_stream
        if pred is not None:
            fused = self.exchange == "fused"
            dst = self.out[:k] if (fused or self.world == 1) else self.local[:k]
            _lib.check(self.lib.csgpu_search_tagged_keys_device(self.store.handle, q_dev.data_ptr(), k, ctypes.byref(pred),
                                                                1 if fused else 0, dst.data_ptr(), stream))
            if fused or self.world == 1:
                return dst
            gathered = allgather_keys(dst, self.world, self.group, self.gathered[: self.world * k])
            out = self.out[:k]
            _lib.check(self.lib.csgpu_merge_keys_device(self.store.handle, gathered.data_ptr(), self.world, k,
                                                        out.data_ptr(), stream))
            return out
        if self.exchange == "fused":
            out = self.out[:k]
            _lib.check(self.lib.csgpu_search_keys_exchange_device(self.store.handle, q_dev.data_ptr(), k,
                                                                  out.data_ptr(), stream))
            return out
        local = self.local[:k]
        _lib.check(self.lib.csgpu_search_keys_device(self.store.handle, q_dev.data_ptr(), k, local.data_ptr(), stream))
        if self.world == 1:
            return local
        gathered = allgather_keys(local, self.world, self.group, self.gathered[: self.world * k])
        out = self.out[:k]
        _lib.check(self.lib.csgpu_merge_keys_device(self.store.handle, gathered.data_ptr(), self.world, k,
                                                    out.data_ptr(), stream))
        return out

    # -- end to end: host query in, host (ids, distances) out
    def search(self, query_embedding, k: int, pred: "_lib.Predicate | None" = None):
        q = np.ascontiguousarray(query_embedding, dtype=np.float32).reshape(-1)
        if q.size != self.store.dimensions:
            raise _lib.CsgpuError(_lib.ERR_DIM, f"Query embedding dimension mismatch: expected "
                                                f"{self.store.dimensions}, got {q.size}")
        if pred is None and self.exchange == "fused":
            # the C ABI's host-pointer form: staging, the one fused launch and the read-back of the k keys (written by the
            # kernel straight into mapped host memory) happen inside the library
            oi = np.empty(max(k, 1), np.uint32); od = np.empty(max(k, 1), np.float32); on = ctypes.c_uint32(0)
            _lib.check(self.lib.csgpu_search_exchange(self.store.handle, q.ctypes.data_as(_lib._f32p), q.size, k,
                                                      oi.ctypes.data_as(_lib._u32p), od.ctypes.data_as(_lib._f32p), ctypes.byref(on)))
            return oi[: on.value], od[: on.value]
        self.q_pin[: q.size].copy_(torch.from_numpy(q))
        self.q_dev.copy_(self.q_pin, non_blocking=True)
        keys = self.search_keys_device(self.q_dev, k, pred)
        self.out_pin[:k].copy_(keys, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self.check_exchange()
        return decode_keys(self.out_pin[:k].numpy())

    def check_exchange(self) -> None:
        """Raises CsgpuError(ERR_NCCL) if an in-kernel wait for a peer's keys ever timed out on this rank: the keys of that
        search (and of every later one) are undefined. The status word is pinned host memory — no device round trip."""
        if self.exchange != "fused":
            return
        t = ctypes.c_uint32(0)
        _lib.check(self.lib.csgpu_exchange_status(self.store.handle, ctypes.byref(t)))
        if t.value:
            raise _lib.CsgpuError(_lib.ERR_NCCL, "cross-GPU exchange timed out: a peer rank never delivered its keys; "
                                                 "results are undefined until the exchange is connected again")


    # -- batches: every rank answers the whole batch over its shard (csgpu_search_batch picks the kernel: multi-query
    #    scan, SIMT GEMM, tcgen05 on a bf16 index or behind the tensor prefilter), ONE all-gather of [b, k] keys, one
    #    batched merge kernel. Returns (ids [b, k], distances [b, k], n [b]) — identical on every rank.
    def search_batch(self, queries, k: int):
        q = np.ascontiguousarray(queries, dtype=np.float32)
        b = q.shape[0]
        oi, od, on = self.store.search_batch_ids(q, k)
        keys = np.empty((b, k), dtype=np.uint64)
        for j in range(b):
            self.lib.csgpu_encode_keys(oi[j].ctypes.data_as(_lib._u32p), od[j].ctypes.data_as(_lib._f32p), int(on[j]), k,
                                       keys[j].ctypes.data_as(_lib._u64p))
        if self.world == 1:
            return oi[:, :k], od[:, :k], on
        local = torch.from_numpy(keys.view(np.int64)).to(self.dev)
        gathered = allgather_keys(local.reshape(-1), self.world, self.group)          # [world][b][k]
        out = torch.empty(b * k, dtype=torch.int64, device=self.dev)
        _lib.check(self.lib.csgpu_merge_keys_batch_device(self.store.handle, gathered.data_ptr(), self.world, b, k,
                                                          out.data_ptr(), torch.cuda.current_stream().cuda_stream))
        merged = out.cpu().numpy().reshape(b, k)
        ids = np.zeros((b, k), np.uint32); dd = np.zeros((b, k), np.float32); n = np.zeros(b, np.uint32)
        for j in range(b):
            i_j, d_j = decode_keys(merged[j])
            n[j] = len(i_j); ids[j, : n[j]] = i_j; dd[j, : n[j]] = d_j
        return ids, dd, n


def decode_keys(keys_i64: np.ndarray):
    lib = _lib.load()
    keys = np.ascontiguousarray(keys_i64).view(np.uint64)
    k = keys.size
    ids = np.empty(k, np.uint32)
    dd = np.empty(k, np.float32)
    n = ctypes.c_uint32()
    lib.csgpu_decode_keys(keys.ctypes.data_as(_lib._u64p), k, ids.ctypes.data_as(_lib._u32p),
                          dd.ctypes.data_as(_lib._f32p), ctypes.byref(n))
    return ids[: n.value], dd[: n.value]
