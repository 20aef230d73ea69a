#!/usr/bin/env python
"""Generates tests/golden/knn_small.json from the fp64 oracle (oracle/knn_oracle.c) on seeded synthetic inputs.
The reference (Go) cannot run here, so the oracle — itself pinned to the reference's KATs by
tests/test_oracle_kat.py — is the generator.  Inputs are reproduced from (n, d, seed) by the shared
counter-based generator, so only the expected outputs are stored.

    python tests/golden/make_knn_fixture.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

cases = []
for (n, d, Q, k, metric, dtype, seed) in [
    (2000, 64, 4, 5, "cosine", "f32", 101),
    (2000, 64, 4, 5, "dot", "f32", 102),
    (2000, 64, 4, 5, "euclidean", "f32", 103),
    (1500, 100, 3, 8, "cosine", "f16", 104),
    (1500, 104, 2, 8, "euclidean", "f16", 105),
    (4097, 33, 9, 3, "cosine", "f32", 106),
]:
    rows = oracle.fill_uniform(n, d, seed, dtype=dtype)
    q = oracle.fill_uniform(Q, d, seed + 5000)
    idx, sc = oracle.knn_exact64(rows, q, k, metric)
    cases.append(dict(n=n, d=d, Q=Q, k=k, metric=metric, dtype=dtype, seed=seed, qseed=seed + 5000,
                      idx=idx.tolist(), score=sc.tolist()))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "knn_small.json")
with open(out, "w") as f:
    json.dump({"generator": "tests/golden/make_knn_fixture.py (oracle.knn_exact64)", "cases": cases}, f, indent=1)
print(out)
