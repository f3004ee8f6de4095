"""Snapshot / hydrate of the device index (csgpu_save / csgpu_load) and the reference's own
persistence test restated (src/vectordb/store.rs:995-1028). GPU only."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cs():
    import codesearch_b200 as m
    m.load_library()
    return m


def test_persistence(cs, tmp_path):
    db = str(tmp_path / "test.db")                                  # store.rs:995-1028
    st = cs.VectorStore.new(db, 4)
    st.insert_chunks([cs.EmbeddedChunk(cs.Chunk("fn test() {}", 0, 1, "Function", "test.rs"), [1.0, 0.0, 0.0, 0.0])])
    st.build_index()
    st.close()
    st2 = cs.VectorStore.new(db, 4)
    assert st2.stats().total_chunks == 1
    assert st2.get_chunk(0) is not None
    assert st2.is_indexed() and st2.next_id == 1
    r = st2.search([0.9, 0.1, 0.0, 0.0], 2)
    assert len(r) == 1 and r[0].id == 0 and r[0].content == "fn test() {}"
    # next ids continue after the maximum key (store.rs:141-144), and a rebuild re-publishes the snapshot
    ids = st2.insert_chunks_with_ids([cs.EmbeddedChunk(cs.Chunk("fn b() {}", 2, 3, "Function", "b.rs"), [0.0, 1.0, 0.0, 0.0])])
    assert ids == [1]
    st2.build_index()
    st2.close()
    ro = cs.VectorStore.open_readonly(db, 4)                        # store.rs:183-250
    assert ro.stats().total_chunks == 2
    with pytest.raises(cs.CsgpuError):
        ro.delete_chunks([0])
    ro.close()
    st3 = cs.VectorStore.new(db, 4)
    st3.clear()                                                     # store.rs:690-706
    st3.close()
    st4 = cs.VectorStore.new(db, 4)
    assert st4.stats().total_chunks == 0 and not st4.is_indexed()


def test_file_table_survives_delete_reinsert_reopen(cs, tmp_path):
    """The row tags carry the file ids assigned at insert time. The reference's incremental reindex flow (delete a
    file's chunks, re-insert them: src/index/mod.rs refresh path over store.rs:548-610,618-686) must not renumber files
    on reopen, or a path filter selects the wrong files' rows (round-1 advisor finding)."""
    db = str(tmp_path / "files.db")
    st = cs.VectorStore.new(db, 4)
    a = st.insert_chunks_with_ids([cs.EmbeddedChunk(cs.Chunk("fn a() {}", 0, 1, "Function", "src/a.rs"), [1.0, 0.0, 0.0, 0.0])])
    b = st.insert_chunks_with_ids([cs.EmbeddedChunk(cs.Chunk("def b(): pass", 0, 1, "Function", "lib/b.py"), [0.0, 1.0, 0.0, 0.0])])
    st.build_index()
    assert st.delete_chunks(a) == 1
    a2 = st.insert_chunks_with_ids([cs.EmbeddedChunk(cs.Chunk("fn a2() {}", 0, 1, "Function", "src/a.rs"), [0.9, 0.1, 0.0, 0.0])])
    st.build_index()
    want_a = [r.id for r in st.search_tagged([0.5, 0.5, 0.0, 0.0], 5, path_prefix="src/")]
    want_b = [r.id for r in st.search_tagged([0.5, 0.5, 0.0, 0.0], 5, path_prefix="lib/")]
    assert want_a == a2 and want_b == b
    st.close()
    st2 = cs.VectorStore.new(db, 4)
    assert st2.files.paths == ["src/a.rs", "lib/b.py"]
    assert [r.id for r in st2.search_tagged([0.5, 0.5, 0.0, 0.0], 5, path_prefix="src/")] == a2
    assert [r.id for r in st2.search_tagged([0.5, 0.5, 0.0, 0.0], 5, path_prefix="lib/")] == b
    assert [r.path for r in st2.search_tagged([0.5, 0.5, 0.0, 0.0], 5, languages=["Python"])] == ["lib/b.py"]
    # every chunk of the FIRST file gone: the second file must keep id 1
    st2.delete_chunks(a2)
    st2.build_index()
    st2.close()
    st3 = cs.VectorStore.new(db, 4)
    assert [r.id for r in st3.search_tagged([0.5, 0.5, 0.0, 0.0], 5, path_prefix="lib/")] == b
    assert st3.search_tagged([0.5, 0.5, 0.0, 0.0], 5, path_prefix="src/") == []
    c = st3.insert_chunks_with_ids([cs.EmbeddedChunk(cs.Chunk("x", 0, 1, "Function", "src/c.rs"), [0.0, 0.0, 1.0, 0.0])])
    st3.build_index()
    assert [r.id for r in st3.search_tagged([0.5, 0.5, 0.5, 0.0], 5, path_prefix="src/")] == c
    st3.close()


@pytest.mark.parametrize("dtype,d", [("fp32", 384), ("fp32", 100), ("bf16", 384)])
def test_snapshot_roundtrip_bit_identical(cs, tmp_path, dtype, d):
    rng = np.random.default_rng(41)
    n = 30000
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[7] = 0.0                                                   # zero-norm row: kept in zero.u32
    ids = rng.permutation(3 * n)[:n].astype(np.uint32)
    db = str(tmp_path / f"snap_{dtype}_{d}.db")
    st = cs.VectorStore.new(db, d, dtype=dtype)
    st.append_rows(rows, ids)
    st.delete_chunks(ids[100:200].tolist())
    st.build_index()                                                # writes <db>/gpu
    meta = json.load(open(os.path.join(db, "gpu", "meta.json")))
    assert meta["rows"] == n - 100 - 1 and meta["zero_ids"] == 1 and meta["dim"] == d
    qs = rng.standard_normal((70, d)).astype(np.float32)
    want = [st.search_ids(q, 50) for q in qs[:5]]
    want_b = st.search_batch_ids(qs, 20)
    st.close()
    st2 = cs.VectorStore.new(db, d, dtype=dtype)
    s = st2.device_stats()
    assert s.built and s.live_rows == n - 100 and s.zero_norm_rows == 1
    for q, (wi, wd) in zip(qs[:5], want):
        gi, gd = st2.search_ids(q, 50)
        assert np.array_equal(gi, wi) and np.array_equal(gd, wd)    # same rows in HBM -> bit-identical
    got_b = st2.search_batch_ids(qs, 20)
    assert all(np.array_equal(a, b) for a, b in zip(got_b, want_b))
    # a loaded index is a normal index: mutate, rebuild, search
    st2.append_rows(rows[:3] * 2, np.array([900000, 900001, 900002], np.uint32))
    st2.build_index()
    assert st2.device_stats().live_rows == n - 100 + 3


def test_snapshot_rejects_corruption_and_mismatch(cs, tmp_path):
    from codesearch_b200 import _lib
    rng = np.random.default_rng(42)
    rows = rng.standard_normal((5000, 128)).astype(np.float32)
    db = str(tmp_path / "c.db")
    st = cs.VectorStore.new(db, 128)
    st.append_rows(rows, np.arange(5000, dtype=np.uint32))
    st.build_index()
    st.close()
    with pytest.raises(cs.CsgpuError) as e:                         # wrong dimensions
        cs.VectorStore.new(db, 64)
    assert e.value.code == _lib.ERR_DIM
    with pytest.raises(cs.CsgpuError):                              # wrong dtype
        cs.VectorStore.new(db, 128, dtype="bf16")
    p = os.path.join(db, "gpu", "rows.f32")
    with open(p, "r+b") as f:                                       # flip one byte -> checksum mismatch
        f.seek(12345)
        b = f.read(1)
        f.seek(12345)
        f.write(bytes([b[0] ^ 0x40]))
    with pytest.raises(cs.CsgpuError) as e:
        cs.VectorStore.new(db, 128)
    assert "checksum" in str(e.value)
