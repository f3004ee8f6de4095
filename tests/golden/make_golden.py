#!/usr/bin/env python
"""Regenerates the fixtures in this directory. Run from the repo root: python tests/golden/make_golden.py

reference_kats.json   the known answers the reference's OWN tests hold for the path, copied by hand
                      from the cited lines, with the numeric values its arithmetic implies (fp32, sequential sums).
c1_100k_top10.json    BASELINE configs[0] (100k x 384 synthetic unit vectors, single query, top-10): ids + distances
                      of the first 8 queries from the f64-accumulating oracle. The reference itself (Rust + arroy
                      0.5.0 + LMDB) cannot be built or imported here, so these are ORACLE outputs ("port"), pinned so
                      that a change in the oracle or in the generator is caught; they are not reference outputs.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402

O.build()

kats = {
    "_source": "hand-copied from /root/reference (file:line in each entry); values = what the reference's fp32 arithmetic yields",
    "insert_and_search": {
        "cite": "src/vectordb/store.rs:846-893",
        "rows": [[1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0]], "query": [0.9, 0.1, 0.0, 0.0], "limit": 2,
        "expect_ids": [0, 1], "expect_len": 2,
        "implied_cos": [0.9938837, 0.11043153], "implied_distance": [0.0030581355, 0.44478422],
        "implied_score": [0.99694186, 0.5552158], "asserted_by_reference": "results.len()==2, results[0] is id 0, score[0] > score[1]",
    },
    "cosine_similarity": {
        "cite": "src/embed/batch.rs:326-340 (helper :316-324)",
        "cases": [{"a": [1, 0, 0], "b": [1, 0, 0], "approx": 1.0, "tol": 0.001},
                  {"a": [1, 0, 0], "b": [0, 1, 0], "approx": 0.0, "tol": 0.001},
                  {"a": [1, 1, 0], "b": [1, 0, 0], "between": [0.70, 0.72]}],
        "zero_norm": {"a": [0, 0, 0], "b": [1, 0, 0], "guarded_helper_returns": 0.0},
    },
    "rrf": {"cite": "src/rerank/mod.rs:57-59", "k": 60.0, "ranked_ids": [7, 3, 9], "scores": [1 / 61, 1 / 62, 1 / 63]},
    "distance_scale": {"cite": "arroy 0.5.0 Cosine::built_distance (SURVEY.md App. A.1); corroborated by src/search/mod.rs:595",
                       "cases": [{"cos": 1.0, "distance": 0.0}, {"cos": 0.0, "distance": 0.5}, {"cos": -1.0, "distance": 1.0}],
                       "zero_norm_distance": 0.0},
}
json.dump(kats, open(os.path.join(HERE, "reference_kats.json"), "w"), indent=1)

n, d, k = 100_000, 384, 10
rows = O.synth_rows(1234, 0, n, d)
qs = O.synth_rows(4321, 0, 8, d)
out = {"config": "BASELINE configs[0]: 100k x 384 synthetic unit vectors (Philox seed 1234), queries seed 4321 rows 0..7, top-10",
       "kind": "oracle (port) output, f64 accumulation; NOT produced by the reference (unbuildable here)",
       "queries": []}
for j in range(8):
    ids, d32, d64 = O.search(rows, qs[j], k)
    out["queries"].append({"ids": ids.tolist(), "distance_f32_hex": [float(x).hex() for x in d32], "distance_f64": d64.tolist()})
json.dump(out, open(os.path.join(HERE, "c1_100k_top10.json"), "w"), indent=1)
print("wrote", os.listdir(HERE))
