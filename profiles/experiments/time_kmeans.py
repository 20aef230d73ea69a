"""Times one k-means assignment + update pass on device: N x 1024 fp32 rows, K = 1000 centroids, (a) uniform random rows
(worst case: every row is nearly equidistant from many centroids, so most rows go through the exact fix-up) and (b) the
clustered corpus of SURVEY.md 8(d) (Gaussian mixture, 1000 centres, sigma = 0.1)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from nornicdb_b200.knn import KnnIndex  # noqa: E402

d, K = 1024, 1000


def run(ix, cen, n, label):
    assign = np.zeros(n, dtype=np.int32)
    for metric in ("euclidean", "cosine"):
        ix.assign_nearest(cen, assign, metric=metric)  # warm-up
        t0 = time.perf_counter()
        ix.assign_nearest(cen, assign, metric=metric)
        dt = time.perf_counter() - t0
        print(f"{label}: assign_nearest {metric}: {dt * 1e3:.1f} ms for {n} rows x {K} centroids "
              f"({2.0 * n * K * d / dt / 1e12:.1f} algorithmic TFLOP/s)")
    t0 = time.perf_counter()
    _, counts = ix.cluster_means(assign, cen)
    print(f"{label}: cluster_means {(time.perf_counter() - t0) * 1e3:.1f} ms, {int((counts > 0).sum())} non-empty clusters")


n = 2_000_000
ix = KnnIndex(d, metric="cosine")
ix.fill_uniform(n, 42)
run(ix, ix.read_rows(0, K).copy(), n, "uniform")
ix.release()

n = 1_000_000
rng = np.random.default_rng(0)
mu = rng.uniform(-1, 1, (K, d)).astype(np.float32)
rows = np.empty((n, d), dtype=np.float32)
for lo in range(0, n, 100_000):
    lab = rng.integers(0, K, 100_000)
    rows[lo:lo + 100_000] = mu[lab] + rng.standard_normal((100_000, d), dtype=np.float32) * 0.1
ix = KnnIndex(d, metric="cosine")
ix.upload(rows)
run(ix, (mu + 0.01).astype(np.float32), n, "mixture")
ix.release()
