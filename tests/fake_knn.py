"""A CPU stand-in for nornicdb_b200.knn.KnnIndex, for the `-m "not gpu"` host-logic tests ONLY: same methods, every
distance computed by the oracle (fp64 exact kNN, oracle k-means steps).  It lets the host mirrors of the reference's types
(EmbeddingIndex, VectorIndex / queryNodes, ClusterIndex: id maps, swap-remove bookkeeping, over-selection, best-of-chunks,
Lloyd loop, routing) run where there is no GPU.  Never imported by the product."""
import numpy as np

import oracle
from nornicdb_b200.knn import KnnError


class FakeKnnIndex:
    def __init__(self, dim, metric="cosine", dtype="f32", devices=(0,)):
        self.dim, self.metric = int(dim), metric
        self.np_dtype = np.float16 if dtype in ("f16", "fp16", "float16") else np.float32
        self._rows = np.empty((0, self.dim), dtype=self.np_dtype)
        self._mask = None
        self.searches = 0

    def _r(self, rows):
        a = np.ascontiguousarray(np.asarray(rows, dtype=self.np_dtype))
        a = a.reshape(-1, self.dim) if a.size else a.reshape(0, self.dim)
        return a

    def release(self):
        self._rows = np.empty((0, self.dim), dtype=self.np_dtype)

    close = release

    def __len__(self):
        return self._rows.shape[0]

    def upload(self, rows):
        self._rows, self._mask = self._r(rows).copy(), None

    def upload_from_f32(self, rows=None, ptr=None, n_rows=None):
        if ptr is not None:
            import ctypes as C
            buf = (C.c_char * (n_rows * self.dim * 4)).from_address(ptr)
            rows = np.frombuffer(bytes(buf), dtype="<f4").reshape(n_rows, self.dim)
        self.upload(np.asarray(rows, dtype=np.float32).astype(self.np_dtype))

    def append(self, rows):
        self._rows, self._mask = np.concatenate([self._rows, self._r(rows)]), None

    def update_row(self, row, vec):
        if not 0 <= row < len(self):
            raise KnnError("row out of range")
        self._rows[row] = self._r(vec)[0]

    def remove_swap(self, row):
        if not 0 <= row < len(self):
            raise KnnError("row out of range")
        self._rows[row] = self._rows[-1]
        self._rows, self._mask = self._rows[:-1].copy(), None

    def read_rows(self, row, n_rows):
        return self._rows[row:row + n_rows].copy()

    def set_row_mask(self, keep):
        self._mask = None if keep is None else np.asarray(keep, dtype=bool).reshape(-1).copy()

    def set_path(self, path):
        pass

    def set_metric(self, metric):
        self.metric = metric

    def set_min_score(self, v):
        self._floor = None if v is None else float(v)

    def set_row_groups(self, group_of_row, n_groups=None):
        self._groups = None if group_of_row is None else np.asarray(group_of_row, dtype=np.int64).reshape(-1).copy()

    def search_groups(self, query, k):
        q = np.asarray(query, dtype=np.float32).reshape(1, -1)
        kept = np.arange(len(self)) if self._mask is None else np.nonzero(self._mask)[0]
        if len(kept) == 0:
            return np.empty(0, np.uint32), np.empty(0, np.uint32), np.empty(0, np.float32)
        idx, sc = oracle.knn_exact64(self._rows[kept], q, len(kept), self.metric)
        floor = getattr(self, "_floor", None)
        out, seen = [], set()
        for r, s in zip(kept[idx[0]].tolist(), sc[0].tolist()):
            if floor is not None and (s > floor if self.metric == "euclidean" else s < floor):
                continue
            g = int(self._groups[r])
            if g in seen:
                continue
            seen.add(g)
            out.append((g, r, s))
            if len(out) == k:
                break
        return (np.array([o[0] for o in out], np.uint32), np.array([o[1] for o in out], np.uint32), np.array([o[2] for o in out], np.float32))

    def last_path(self):
        return "fake"

    def stats(self):
        return {"kernel_launches": self.searches, "searches": self.searches, "bytes_h2d": 0, "bytes_d2h": 0}

    def search(self, queries, k):
        q = np.ascontiguousarray(np.asarray(queries, dtype=np.float32))
        if q.ndim == 1:
            q = q.reshape(1, -1)
        if q.shape[1] != self.dim:
            raise KnnError(f"invalid dimensions: query has {q.shape[1]}, index has {self.dim}")
        self.searches += 1
        rows = self._rows
        kept = np.arange(len(self)) if self._mask is None else np.nonzero(self._mask)[0]
        ke = min(int(k), len(kept))
        if ke <= 0 or q.shape[0] == 0:
            return np.empty((q.shape[0], 0), np.uint32), np.empty((q.shape[0], 0), np.float32)
        idx, sc = oracle.knn_exact64(rows[kept], q, ke, self.metric)
        gi, gs = kept[idx].astype(np.uint32), sc.astype(np.float32)
        floor = getattr(self, "_floor", None)
        if floor is not None:  # the device leaves 0xffffffff / 0 in slots below the score floor
            bad = sc > floor if self.metric == "euclidean" else sc < floor
            gi[bad] = 0xFFFFFFFF
            gs[bad] = 0.0
        return gi, gs

    def score_subset(self, query, rows, k=None):
        q = np.ascontiguousarray(np.asarray(query, dtype=np.float32).reshape(-1))
        if q.size != self.dim:
            raise KnnError(f"invalid dimensions: query has {q.size}, index has {self.dim}")
        r = np.asarray(rows, dtype=np.int64).reshape(-1)
        kk = len(r) if k is None else min(int(k), len(r))
        if kk <= 0:
            return np.empty(0, np.uint32), np.empty(0, np.float32)
        idx, sc = oracle.knn_exact64(self._rows[r], q.reshape(1, -1), kk, self.metric)
        return r[idx[0]].astype(np.uint32), sc[0].astype(np.float32)

    def assign_nearest(self, centroids, assign, metric="euclidean"):
        return oracle.kmeans_assign(self._rows.astype(np.float32), centroids, assign, by_cosine=metric == "cosine")

    def cluster_means(self, assign, centroids):
        return oracle.kmeans_update(self._rows.astype(np.float32), assign, centroids)
