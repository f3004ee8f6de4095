"""Pins the oracle against every known answer the reference's own tests hold for the path
 and against published Philox vectors. CPU only."""
import numpy as np
import pytest


def test_philox_known_answers(oracle):
    # Random123 kat_vectors, philox4x32 10 rounds
    kats = [((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
            ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
            ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
             (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1))]
    for ctr, key, want in kats:
        assert tuple(int(x) for x in oracle.c_philox(ctr, key)) == want
        got = oracle.np_philox4x32_10(*[np.uint32(c) for c in ctr], key[0], key[1])
        assert tuple(int(x) for x in got) == want


def test_synth_c_matches_numpy(oracle):
    for seed, first, n, d in [(1234, 0, 64, 384), (4321, 7, 5, 768), (9, (1 << 33) + 5, 3, 8)]:
        a = oracle.synth_rows(seed, first, n, d)
        b = oracle.np_synth_rows(seed, first, n, d)
        assert np.array_equal(a, b)
    r = oracle.synth_rows(1234, 0, 2000, 384)
    assert abs(r.mean()) < 1.0 and 140 < r.std() < 156  # Irwin-Hall(4) over bytes: sigma = 147.8
    assert np.abs(r).max() <= 510


def test_reference_cosine_known_answers(oracle):
    # src/embed/batch.rs:326-340
    for f in (oracle.py_cosine_similarity, lambda a, b: oracle.c_cosine(a, b), lambda a, b: oracle.c_cosine(a, b, True)):
        assert abs(f([1, 0, 0], [1, 0, 0]) - 1.0) < 0.001
        assert abs(f([1, 0, 0], [0, 1, 0]) - 0.0) < 0.001
        s = f([1, 1, 0], [1, 0, 0])
        assert 0.7 < s < 0.72
    # zero-magnitude convention of the guarded helper (batch.rs:320-322); the unguarded one is 0/0
    assert oracle.c_cosine([0, 0, 0], [1, 0, 0], guarded=True) == 0.0
    assert np.isnan(oracle.c_cosine([0, 0, 0], [1, 0, 0]))
    assert np.isnan(oracle.py_cosine_similarity([0, 0, 0], [1, 0, 0]))


def test_reference_insert_and_search_case(oracle):
    # src/vectordb/store.rs:846-893: id0=[1,0,0,0], id1=[0,1,0,0], q=[0.9,0.1,0,0], limit 2
    rows = np.array([[1, 0, 0, 0], [0, 1, 0, 0]], dtype=np.float32)
    q = np.array([0.9, 0.1, 0, 0], dtype=np.float32)
    for mode in (0, 1):
        ids, dist, _ = oracle.search(rows, q, 2, mode=mode)
        assert ids.tolist() == [0, 1]
        score = oracle.score_from_distance(dist)
        assert score[0] > score[1]
        # values implied by the reference's arithmetic
        assert np.allclose(dist, [0.0030581355, 0.44478422], atol=2e-7)
        assert np.allclose(score, [0.99694186, 0.5552158], atol=2e-7)
    assert abs(float(oracle.py_distance_f32(rows[0], q)) - 0.0030581355) < 2e-7
    assert abs(oracle.c_distance_f32(rows[0], q) - float(oracle.py_distance_f32(rows[0], q))) == 0.0
    # min(k, N) results (arroy capacity = count.min(len); pinned by results.len()==2 above)
    ids, _, _ = oracle.search(rows, q, 10)
    assert len(ids) == 2


def test_distance_scale_and_zero_norm(oracle):
    a = np.array([1, 0, 0, 0], dtype=np.float32)
    assert oracle.c_distance_f32(a, a) == 0.0              # cos=1  -> 0
    assert oracle.c_distance_f32(a, -a) == 1.0             # cos=-1 -> 1
    assert oracle.c_distance_f32(a, [0, 1, 0, 0]) == 0.5   # cos=0  -> 0.5
    assert oracle.c_distance_f32([0, 0, 0, 0], a) == 0.0   # pn*qn == 0 -> 0
    assert float(oracle.py_distance_f32([0, 0, 0, 0], a)) == 0.0


def test_c_and_numpy_search_agree(oracle):
    rng = np.random.default_rng(0)
    for n, d, k in [(1, 4, 3), (7, 8, 7), (1000, 384, 10), (5000, 96, 100), (3000, 768, 200)]:
        rows = rng.standard_normal((n, d)).astype(np.float32)
        q = rng.standard_normal(d).astype(np.float32)
        ids = rng.permutation(n * 3)[:n].astype(np.uint32)
        i1, d1, d64 = oracle.search(rows, q, k, ids=ids)
        i2, d2, e64 = oracle.np_search(rows, q, k, ids=ids)
        assert np.array_equal(i1, i2) and np.array_equal(d1, d2)
        assert np.allclose(d64, e64, atol=1e-12)
        assert len(i1) == min(k, n)
        # strict-f32 sequential restatement agrees to within fp32 rounding
        i0, d0, _ = oracle.search(rows, q, k, ids=ids, mode=0)
        assert np.abs(d0 - d1).max() < 5e-7


def test_ties_break_by_id_and_duplicates(oracle):
    base = np.random.default_rng(1).standard_normal((4, 16)).astype(np.float32)
    rows = np.concatenate([base, base, base])          # every row appears 3 times
    ids = np.array([40, 10, 30, 20, 41, 11, 31, 21, 42, 12, 32, 22], dtype=np.uint32)
    i1, d1, _ = oracle.search(rows, base[0], 12, ids=ids)
    assert i1[:3].tolist() == [40, 41, 42] and d1[0] == d1[1] == d1[2]
    for a, b in zip(range(0, 12, 3), range(3, 12, 3)):
        pass
    # groups of three equal distances, ids ascending inside each group
    for g in range(4):
        grp = i1[3 * g:3 * g + 3]
        assert sorted(grp.tolist()) == grp.tolist()
        assert len(set(d1[3 * g:3 * g + 3].tolist())) == 1


def test_filtered_oracle(oracle):
    rng = np.random.default_rng(2)
    rows = rng.standard_normal((500, 32)).astype(np.float32)
    q = rng.standard_normal(32).astype(np.float32)
    allowed = rng.random(500) < 0.25
    bm = np.packbits(np.pad(allowed, (0, 12)).astype(np.uint8).reshape(-1, 8), axis=1, bitorder="little").reshape(-1).view(np.uint64)
    i1, d1, _ = oracle.search(rows, q, 20, bitmap=bm, n_bits=500)
    i2, d2, _ = oracle.np_search(rows, q, 20, allowed=allowed)
    assert np.array_equal(i1, i2) and np.array_equal(d1, d2)
    assert allowed[i1].all()
    # superset-preserving: the unfiltered ranking restricted to allowed ids, same relative order
    iu, _, _ = oracle.search(rows, q, 500)
    assert [i for i in iu if allowed[i]][:20] == i1.tolist()


def test_streaming_synth_oracle_matches_materialised(oracle):
    rows = oracle.synth_rows(1234, 100, 20000, 384)
    qs = oracle.synth_rows(4321, 0, 3, 384)
    oi, od, o64, on = oracle.search_synth(1234, 100, 20000, 384, qs, 10)
    for j in range(3):
        i1, d1, d64 = oracle.search(rows, qs[j], 10, ids=np.arange(100, 20100, dtype=np.uint32))
        assert np.array_equal(oi[j], i1) and np.array_equal(od[j], d1)
        assert np.allclose(o64[j], d64, atol=1e-12)
        assert on[j] == 10


def test_cpu_baseline_matches_oracle(oracle):
    rows = oracle.synth_rows(1234, 0, 50000, 384)
    q = oracle.synth_rows(4321, 0, 1, 384)[0]
    i1, d1, _ = oracle.search(rows, q, 10)
    i3, d3 = oracle.cpu_baseline_search(rows, q, 10)
    assert np.array_equal(i1, i3) and np.abs(d1 - d3).max() < 5e-7


def test_rrf_consumer_contract(oracle):
    # src/rerank/mod.rs:57-59: rank = position, score = 1/(k + rank + 1)
    s = oracle.rrf_scores([7, 3, 9], 60.0)
    assert s[7] == pytest.approx(1 / 61) and s[3] == pytest.approx(1 / 62) and s[9] == pytest.approx(1 / 63)


# ---- committed golden fixtures (tests/golden/, regenerated by tests/golden/make_golden.py) ---------------
def _golden(name):
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name)))


def test_golden_reference_kats(oracle):
    g = _golden("reference_kats.json")
    c = g["insert_and_search"]
    rows, q = np.array(c["rows"], np.float32), np.array(c["query"], np.float32)
    ids, dist, _ = oracle.search(rows, q, c["limit"])
    assert ids.tolist() == c["expect_ids"] and len(ids) == c["expect_len"]
    assert np.allclose(dist, c["implied_distance"], atol=2e-7)
    assert np.allclose(oracle.score_from_distance(dist), c["implied_score"], atol=2e-7)
    for case in g["cosine_similarity"]["cases"]:
        s = float(oracle.py_cosine_similarity(case["a"], case["b"]))
        if "approx" in case:
            assert abs(s - case["approx"]) < case["tol"]
        else:
            assert case["between"][0] < s < case["between"][1]
    z = g["cosine_similarity"]["zero_norm"]
    assert oracle.c_cosine(z["a"], z["b"], guarded=True) == z["guarded_helper_returns"]
    r = g["rrf"]
    s = oracle.rrf_scores(r["ranked_ids"], r["k"])
    assert [s[i] for i in r["ranked_ids"]] == pytest.approx(r["scores"])
    for case in g["distance_scale"]["cases"]:
        a = np.array([1, 0], np.float32)
        b = np.array([case["cos"], np.sqrt(max(0.0, 1 - case["cos"] ** 2))], np.float32)
        assert oracle.c_distance_f32(a, b) == pytest.approx(case["distance"], abs=1e-7)


def test_golden_c1_config0(oracle):
    g = _golden("c1_100k_top10.json")
    rows = oracle.synth_rows(1234, 0, 100_000, 384)
    qs = oracle.synth_rows(4321, 0, len(g["queries"]), 384)
    for q, want in zip(qs, g["queries"]):
        for fn in (oracle.search, oracle.np_search):                 # C and numpy restatements both hit the fixture
            ids, d32, d64 = fn(rows, q, 10)
            assert ids.tolist() == want["ids"]
            assert [float(x).hex() for x in d32] == want["distance_f32_hex"]
            assert np.allclose(d64, want["distance_f64"], atol=1e-12)


def test_dedup_variants_restatement(oracle):
    # src/search/mod.rs:513-590: per id the best score over the variant lists, top-N of the union, best first
    lists = [(np.array([5, 2, 9], np.uint32), np.array([0.10, 0.20, 0.30], np.float32)),
             (np.array([2, 7, 5], np.uint32), np.array([0.05, 0.20, 0.25], np.float32)),
             (np.array([], np.uint32), np.array([], np.float32))]
    ids, dist = oracle.dedup_variants(lists, 3)
    assert ids.tolist() == [2, 5, 7] and dist.tolist() == [np.float32(0.05), np.float32(0.10), np.float32(0.20)]
    ids, dist = oracle.dedup_variants(lists, 10)
    assert ids.tolist() == [2, 5, 7, 9]
