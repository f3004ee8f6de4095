"""Randomised lifecycle test: long seeded sequences of the reference's write/read cycle — insert (with replacements and
zero-norm rows), delete, build_index, reopen from the snapshot, and every search entry point in between — run against
three configurations of the same store (one shard; two shards on device 0 = the whole multi-shard host logic incl. the
fused in-process exchange; one shard with the byte prefilter) and a plain numpy model of what the store must contain.

What is asserted after every build: the three configurations agree bit for bit on ids AND distances for plain, filtered,
tagged, batched and variant searches, and the plain search agrees with the f64 oracle over the model's rows. This is the
state machine of /root/reference/src/vectordb/store.rs (insert :618-686 flips indexed=false, delete :548-610 leaves id
gaps, build_index :386-430, new/open :110-176), where round 1's advisor found the bugs the unit tests had not (file table
after delete + re-insert, uneven shards in the snapshot).
"""
import os

import numpy as np
import pytest

from parity import check_topk

pytestmark = pytest.mark.gpu
os.environ.setdefault("CSGPU_I8_MIN_ROWS", "4096")

MARGIN = 8


@pytest.fixture(scope="module")
def cs():
    import codesearch_b200 as m
    m.load_library()
    return m


def _same(a, b):
    return np.array_equal(a[0], b[0]) and np.array_equal(np.asarray(a[1]).view(np.uint32), np.asarray(b[1]).view(np.uint32))


@pytest.mark.parametrize("seed,d", [(1, 96), (2, 384), (3, 128), (4, 768), (5, 100), (6, 384)])
def test_random_lifecycle_three_configurations_agree(cs, oracle, tmp_path, seed, d):
    from codesearch_b200 import tags as T
    rng = np.random.default_rng(1000 + seed)
    dbs = [str(tmp_path / f"db{i}") for i in range(3)]
    devs = [[0], [0, 0], [0]]

    def open_all():
        out = [cs.VectorStore.new(dbs[i], d, devices=devs[i]) for i in range(3)]
        out[2].set_byte_prefilter(True)
        return out

    stores = open_all()
    model = {}                      # chunk id -> (row, tag)
    next_id = 0
    built = False

    def do_append():
        nonlocal next_id, built
        n = int(rng.choice([1, 7, 300, 5000, 9000]))
        rows = rng.standard_normal((n, d)).astype(np.float32)
        if n >= 7 and rng.random() < 0.5:
            rows[rng.integers(0, n)] = 0.0                                   # a zero-norm row (distance 0.0 side list)
        ids = np.arange(next_id, next_id + n, dtype=np.uint32)
        if model and n >= 7 and rng.random() < 0.4:                          # re-insert some live ids: replace semantics
            live = np.fromiter(model.keys(), dtype=np.uint32)
            take = rng.choice(live, size=min(len(live), 5), replace=False)
            ids[: len(take)] = take
        next_id += n
        tg = ((rng.integers(0, 23, size=n).astype(np.uint32)) << np.uint32(27)) | rng.integers(0, 400, size=n).astype(np.uint32)
        for st in stores:
            st.append_rows(rows, ids, tg)
        for i in range(n):                                                   # within a batch the last occurrence of an id wins
            model[int(ids[i])] = (rows[i], int(tg[i]))
        built = False

    def do_delete():
        nonlocal built
        if not model:
            return
        live = np.fromiter(model.keys(), dtype=np.uint32)
        dead = rng.choice(live, size=max(1, len(live) // int(rng.choice([3, 10, 50]))), replace=False)
        dead = np.concatenate([dead, np.array([next_id + 5], np.uint32)])    # and one id that never existed
        counts = [st.delete_chunks(dead.tolist()) for st in stores]
        assert counts[0] == counts[1] == counts[2] == len(dead) - 1
        for i in dead[:-1]:
            del model[int(i)]
        built = False

    def do_build_and_check():
        nonlocal built
        for st in stores:
            st.build_index()
        built = True
        live_ids = np.array(sorted(model), dtype=np.uint32)
        for st in stores:
            assert st.device_stats().live_rows == len(live_ids)
        if len(live_ids) == 0:
            return
        rows = np.stack([model[int(i)][0] for i in live_ids])
        tg = np.array([model[int(i)][1] for i in live_ids], dtype=np.uint32)
        for st in stores:
            assert np.array_equal(st.get_tags(live_ids[:50]), tg[:50])
        qs = rng.standard_normal((12, d)).astype(np.float32)
        for k in (int(rng.choice([1, 10, 40])), int(rng.choice([100, 256, 700]))):
            res = [st.search_ids(qs[0], k) for st in stores]
            assert _same(res[0], res[1]) and _same(res[0], res[2]), ("plain", k)
            oi, od, o64 = oracle.np_search(rows, qs[0], k + MARGIN, ids=live_ids)
            check_topk(res[0][0], res[0][1], oi, od, o64, min(k, len(live_ids)))
            mask_ids = rng.random(next_id + 10) < float(rng.choice([0.9, 0.3, 0.03]))
            flt = cs.RowFilter.from_mask(mask_ids)
            res = [st.search_ids(qs[1], k, flt) for st in stores]
            assert _same(res[0], res[1]) and _same(res[0], res[2]), ("filtered", k)
            allowed = mask_ids[live_ids]
            oi, od, o64 = oracle.np_search(rows[allowed], qs[1], k + MARGIN, ids=live_ids[allowed])
            check_topk(res[0][0], res[0][1], oi, od, o64, min(k, int(allowed.sum())))
            pred = T.TagPredicate(lang_mask=int(rng.integers(1, 1 << 23)), file_lo=int(rng.integers(0, 100)), file_hi=int(rng.integers(100, 400)))
            res = [st.search_tagged_ids(qs[2], k, pred) for st in stores]
            assert _same(res[0], res[1]) and _same(res[0], res[2]), ("tagged", k)
            ok = pred.passes(tg)
            assert len(res[0][0]) == min(k, int(ok.sum())) and set(res[0][0].tolist()) <= set(live_ids[ok].tolist())
        b, k = int(rng.choice([3, 9, 16])), int(rng.choice([10, 100]))
        bat = [st.search_batch_ids(qs[:b], k) for st in stores]
        assert np.array_equal(bat[0][2], bat[1][2]) and _same(bat[0], bat[1]) and _same(bat[0], bat[2])
        var = [st.search_variants_ids(qs[:b], k) for st in stores]
        assert _same(var[0], var[1]) and _same(var[0], var[2])

    for step in range(14):
        op = rng.choice(["append", "append", "delete", "build", "reopen"]) if step else "append"
        if op == "append":
            do_append()
        elif op == "delete":
            do_delete()
        elif op == "build":
            do_build_and_check()
        elif op == "reopen":
            do_build_and_check()                                 # build_index publishes the snapshot (store.rs: commit)
            for st in stores:
                st.close()
            stores = open_all()                                  # new(): hydrates from <db>/gpu, indexed = snapshot present
            for st in stores:
                assert st.is_indexed() == bool(model) or not model
            if model:
                q = rng.standard_normal(d).astype(np.float32)
                res = [st.search_ids(q, 20) for st in stores]
                assert _same(res[0], res[1]) and _same(res[0], res[2])
        if not built and model and rng.random() < 0.3:           # search on a dirty index: the reference's literal error
            with pytest.raises(cs.CsgpuError) as e:
                stores[int(rng.integers(0, 3))].search_ids(np.ones(d, np.float32), 5)
            assert e.value.code == 2 and "Index not built" in str(e.value)
    do_build_and_check()
    for st in stores:
        st.close()
