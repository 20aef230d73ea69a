"""Times one k-means assignment + update pass on device (N=2M x 1024 fp32 rows, K=1000 centroids)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from nornicdb_b200.knn import KnnIndex  # noqa: E402

n, d, K = 2_000_000, 1024, 1000
ix = KnnIndex(d, metric="cosine")
ix.fill_uniform(n, 42)
cen = ix.read_rows(0, K).copy()
assign = np.zeros(n, dtype=np.int32)
for metric in ("euclidean", "cosine"):
    ix.assign_nearest(cen, assign, metric=metric)  # warm-up (workspaces)
    t0 = time.perf_counter()
    changed = ix.assign_nearest(cen, assign, metric=metric)
    dt = time.perf_counter() - t0
    print(f"assign_nearest {metric}: {dt * 1e3:.1f} ms for {n} rows x {K} centroids ({2.0 * n * K * d / dt / 1e12:.1f} algorithmic TFLOP/s), changed={changed}")
t0 = time.perf_counter()
new, counts = ix.cluster_means(assign, cen)
print(f"cluster_means: {(time.perf_counter() - t0) * 1e3:.1f} ms, {int((counts > 0).sum())} non-empty clusters")
ix.release()
