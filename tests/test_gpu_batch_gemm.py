"""Large-batch search on the default fp32 index: register-tiled fp32 SIMT GEMM + fused threshold filter
(codesearch_b200/csrc/gemm_simt.cuh; BASELINE config C3 "fp32 SIMT"). GPU only, through the C ABI.

Bar: the same as the single-query path — ids identical to the exact-cosine oracle in (distance, id) order
(a swap only inside an f64 near-tie of < 4e-7), |distance - oracle| <= 1e-5. The GEMM kernel sums k = 0..dim-1
in one chain per (query, row), the scan kernels use a lane-strided chain + shuffle tree, so the two fp32
paths may differ in the last ulp of a distance; both are held to the oracle, and to each other within 2e-6.

Round 2: the default batch route of an fp32 index is the tf32 tensor-core filter + exact rescoring
(tests/test_gpu_tf32_batch.py); this file pins the SIMT kernel — the route for dim > 1024 — with CSGPU_BATCH_SIMT=1.
"""
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


@pytest.fixture(autouse=True)
def simt_route():
    old = os.environ.get("CSGPU_BATCH_SIMT")
    os.environ["CSGPU_BATCH_SIMT"] = "1"
    yield
    if old is None:
        os.environ.pop("CSGPU_BATCH_SIMT", None)
    else:
        os.environ["CSGPU_BATCH_SIMT"] = old


def make_store(cs, rows, ids=None):
    st = cs.VectorStore.new(None, rows.shape[1])
    st.append_rows(rows, np.arange(rows.shape[0], dtype=np.uint32) if ids is None else ids)
    st.build_index()
    return st


@pytest.mark.parametrize("n,d,b,k", [
    (20000, 384, 40, 10), (20000, 384, 128, 100), (30011, 384, 130, 100), (9000, 768, 64, 200),
    (5000, 100, 50, 10), (5000, 36, 41, 33), (12345, 1024, 48, 7), (4000, 384, 257, 1000),
    (100, 384, 64, 10), (129, 128, 1024, 5), (70000, 64, 300, 100),
])
def test_gemm_batch_parity(cs, oracle, n, d, b, k):
    from codesearch_b200 import _lib
    rng = np.random.default_rng(n + 3 * d + 5 * b + 7 * k)
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[n // 2] = 0.0                                            # a zero-norm row rides along (distance 0.0)
    rows[n // 3] = rows[n // 3 + 1]                               # and an exact duplicate (ties by id)
    ids = rng.permutation(n * 2)[:n].astype(np.uint32)
    st = make_store(cs, rows, ids)
    qs = rng.standard_normal((b, d)).astype(np.float32)
    qs[1] = 0.0                                                   # zero-norm query: distance 0.0 everywhere
    qs[2] = rows[n // 3]                                          # query equal to the duplicated row
    l0 = _lib.load().csgpu_kernel_launches()
    oi, od, on = st.search_batch_ids(qs, k)
    assert _lib.load().csgpu_kernel_launches() - l0 < b           # not one scan per query
    k_eff = min(k, n)
    for j in list(range(min(b, 5))) + [b // 2, b - 1]:
        ri, rd, r64 = oracle.np_search(rows, qs[j], k + MARGIN, ids=ids)
        assert on[j] == k_eff
        check_topk(oi[j, :k_eff], od[j, :k_eff], ri, rd, r64, k_eff)
        gi, gd = st.search_ids(qs[j], k)                          # the scan kernel on the same index
        assert np.abs(gd - od[j, :k_eff]).max() <= 2e-6
    # the duplicate pair comes back first for query 2, smaller id first, equal distances
    a, c = sorted((int(ids[n // 3]), int(ids[n // 3 + 1])))
    if k_eff >= 3:                                                # (the zero-norm row, distance 0.0, may sit among them)
        top3 = oi[2, :3].tolist()
        assert a in top3 and c in top3 and top3.index(a) < top3.index(c)
        assert od[2, top3.index(a)] == od[2, top3.index(c)]


def test_gemm_batch_sorted_corpus_overflow_path(cs, oracle):
    """Adversarial order (every later row beats the threshold): candidate buffers overflow, the host halves
    the phase and retries; results stay exact."""
    rng = np.random.default_rng(5)
    n, d, b, k = 60000, 128, 64, 10
    q = rng.standard_normal(d).astype(np.float32)
    rows = rng.standard_normal((n, d)).astype(np.float32)
    cos = (rows @ q) / np.linalg.norm(rows, axis=1)
    rows = rows[np.argsort(cos)]                                  # ascending similarity: best rows last
    st = make_store(cs, rows)
    qs = np.tile(q, (b, 1)) + 0.01 * rng.standard_normal((b, d)).astype(np.float32)
    oi, od, on = st.search_batch_ids(qs, k)
    for j in (0, 17, b - 1):
        ri, rd, r64 = oracle.np_search(rows, qs[j], k + MARGIN)
        check_topk(oi[j], od[j], ri, rd, r64, k)


def test_gemm_batch_after_mutation(cs, oracle):
    """append / delete / rebuild re-encodes the TMA map over the (moved) row matrix."""
    rng = np.random.default_rng(6)
    d, b, k = 384, 64, 20
    rows = rng.standard_normal((9000, d)).astype(np.float32)
    st = cs.VectorStore.new(None, d)
    st.append_rows(rows[:3000], np.arange(3000, dtype=np.uint32))
    st.build_index()
    qs = rng.standard_normal((b, d)).astype(np.float32)
    st.search_batch_ids(qs, k)
    st.delete_chunks(list(range(100, 200)))
    st.append_rows(rows[3000:], np.arange(3000, 9000, dtype=np.uint32))   # forces the row matrix to grow
    with pytest.raises(cs.CsgpuError) as e:
        st.search_batch_ids(qs, k)
    assert e.value.code == 2
    st.build_index()
    keep = np.ones(9000, bool); keep[100:200] = False
    oi, od, on = st.search_batch_ids(qs, k)
    for j in (0, b - 1):
        ri, rd, r64 = oracle.np_search(rows[keep], qs[j], k + MARGIN, ids=np.arange(9000, dtype=np.uint32)[keep])
        check_topk(oi[j], od[j], ri, rd, r64, k)


def test_gemm_batch_synthetic_200k(cs, oracle):
    """The bench corpus generator at a size the oracle finishes in seconds; B = 256, k = 100."""
    n, d, b, k = 200_000, 384, 256, 100
    st = cs.VectorStore.new(None, d)
    st.append_synthetic(1234, 0, n)
    st.build_index()
    rows = oracle.synth_rows(1234, 0, n, d)
    qs = oracle.synth_rows(4321, 0, b, d)
    oi, od, on = st.search_batch_ids(qs, k)
    swaps = 0
    for j in range(0, b, 16):
        ri, rd, r64 = oracle.search(rows, qs[j], k + MARGIN)
        swaps += check_topk(oi[j], od[j], ri, rd, r64, k)
    assert swaps == 0
