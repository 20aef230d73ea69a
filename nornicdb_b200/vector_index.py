"""Front-end semantics of the callers of the hot path, on top of the fused kernel:

  VectorIndex            mirror of search.VectorIndex (pkg/search/vector_index.go:155-361): vectors are normalised on
                         Add (:234), Search normalises the query, scores by dot product (= cosine), drops everything below
                         minSimilarity, sorts descending, truncates to `limit`, returns float64 scores.
  query_nodes            the scoring loop of CALL db.index.vector.queryNodes (pkg/cypher/call_vector.go:177-256): every node
                         owns one or more chunk embeddings, its score is the BEST over its chunks under the index's
                         similarity function (cosine | dot | euclidean as 1/(1+d), similarity.go:152-158), nodes whose best
                         score is negative are dropped (:240-242), top k nodes returned.
Host logic only (id maps, over-selection, de-duplication by node); all distance work is the CUDA scan."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

from .knn import KnnIndex


class ErrDimensionMismatch(ValueError):  # search.ErrDimensionMismatch
    def __init__(self):
        super().__init__("search: dimension mismatch")


def normalize(vec: np.ndarray) -> np.ndarray:
    """vector.Normalize (pkg/math/vector/similarity.go:197-210): fp32 norm, multiply by 1/norm, zero vector stays zero."""
    v = np.asarray(vec, dtype=np.float32)
    n = np.float32(np.sqrt(np.sum(v * v, dtype=np.float32)))
    if n == 0:
        return np.zeros_like(v)
    return (v * (np.float32(1.0) / n)).astype(np.float32)


class VectorIndex:
    def __init__(self, dimensions: int, devices: Sequence[int] = (0,)):
        self.dimensions = int(dimensions)
        self._ids: List[str] = []
        self._pos: Dict[str, int] = {}
        self._ix = KnnIndex(self.dimensions, metric="dot", devices=devices)  # normalised rows: dot == cosine

    def Add(self, id: str, vec) -> None:  # vector_index.go:224-236
        v = np.asarray(vec, dtype=np.float32).reshape(-1)
        if v.size != self.dimensions:
            raise ErrDimensionMismatch()
        v = normalize(v)
        if id in self._pos:
            self._ix.update_row(self._pos[id], v)
        else:
            self._ix.append(v)
            self._pos[id] = len(self._ids)
            self._ids.append(id)

    def Remove(self, id: str) -> None:  # vector_index.go:239-243
        idx = self._pos.pop(id, None)
        if idx is None:
            return
        last = len(self._ids) - 1
        self._ix.remove_swap(idx)
        if idx != last:
            moved = self._ids[last]
            self._ids[idx] = moved
            self._pos[moved] = idx
        self._ids.pop()

    def Count(self) -> int:
        return len(self._ids)

    def HasVector(self, id: str) -> bool:
        return id in self._pos

    def Search(self, query, limit: int, minSimilarity: float) -> List[Tuple[str, float]]:  # vector_index.go:312-361
        q = np.asarray(query, dtype=np.float32).reshape(-1)
        if q.size != self.dimensions:
            raise ErrDimensionMismatch()
        if not self._ids or limit <= 0:
            return []
        idx, sc = self._ix.search(normalize(q), min(limit, len(self._ids)))
        out = []
        for i, s in zip(idx[0].tolist(), sc[0].tolist()):
            if float(s) < minSimilarity:  # sorted descending: everything after is below the cut too
                break
            out.append((self._ids[i], float(s)))
        return out

    def Release(self) -> None:
        self._ix.release()


def query_nodes(node_chunks: Sequence[Tuple[str, Sequence[Sequence[float]]]], query, k: int, similarity: str = "cosine",
                devices: Sequence[int] = (0,)) -> List[Tuple[str, float]]:
    """node_chunks: [(node_id, [chunk embedding, ...]), ...] -> top-k [(node_id, best-of-chunks score float64)]."""
    q = np.asarray(query, dtype=np.float32).reshape(-1)
    d = q.size
    rows, owner = [], []
    for ni, (_, chunks) in enumerate(node_chunks):
        for c in chunks:
            c = np.asarray(c, dtype=np.float32).reshape(-1)
            if c.size == d:  # chunks of another dimension are skipped (call_vector.go:216-218)
                rows.append(c)
                owner.append(ni)
    if not rows or k <= 0:
        return []
    metric = {"euclidean": "euclidean", "dot": "dot"}.get(similarity, "cosine")
    ix = KnnIndex(d, metric=metric, devices=devices)
    try:
        ix.upload(np.stack(rows))
        n = len(rows)
        want = min(n, max(2 * k, 16))
        while True:
            idx, sc = ix.search(q, want)
            best: Dict[int, float] = {}
            for r, s in zip(idx[0].tolist(), sc[0].tolist()):
                s = 1.0 / (1.0 + float(s)) if metric == "euclidean" else float(s)
                ni = owner[r]
                if ni not in best:  # results arrive best-first, so the first chunk of a node is its best
                    best[ni] = s
            ranked = [(node_chunks[ni][0], s) for ni, s in best.items() if s >= 0.0]  # bestScore >= 0 (call_vector.go:240)
            exhausted = want >= n or (metric != "euclidean" and sc[0, -1] < 0.0)
            if len(ranked) >= k or exhausted:
                return ranked[:k]
            want = min(n, want * 4)
    finally:
        ix.release()
