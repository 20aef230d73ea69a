"""Timing of db.index.vector.queryNodes on the device ("1M-chunk case timed").
1M chunk rows (d=1024) of 250k nodes resident in HBM; one nk_search_groups call per query: segment-max scan + top-k.
Prints ms per query (host-synchronous, includes H2D of the query and D2H of k results) and the HBM rate of the scan."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
from nornicdb_b200.knn import KnnIndex

n, d, nodes, k = 1_000_000, 1024, 250_000, 10
ix = KnnIndex(d, metric="cosine")
ix.fill_uniform(n, 42)
rng = np.random.default_rng(1)
group = np.sort(rng.integers(0, nodes, n)).astype(np.uint32)
ix.set_row_groups(group, nodes)
ix.set_min_score(0.0)
q = rng.uniform(-1, 1, (32, d)).astype(np.float32)
for i in range(3):
    ix.search_groups(q[i], k)
t0 = time.perf_counter()
for i in range(3, 32):
    g, r, s = ix.search_groups(q[i], k)
ms = (time.perf_counter() - t0) / 29 * 1e3
print(json.dumps({"what": "queryNodes best-of-chunks on device", "chunks": n, "nodes": nodes, "dim": d, "k": k,
                  "ms_per_query": ms, "scan_gbs": n * d * 4 / (ms / 1e3) / 1e9, "first": [int(g[0]), int(r[0]), float(s[0])]}))
ix.release()
