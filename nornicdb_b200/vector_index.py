"""Front-end semantics of the callers of the hot path, evaluated ON THE DEVICE:

  VectorIndex            mirror of search.VectorIndex (pkg/search/vector_index.go:155-361): vectors are normalised on
                         Add (:234), Search normalises the query, scores by dot product (= cosine), drops everything below
                         minSimilarity, sorts descending, truncates to `limit`, returns float64 scores.  The minSimilarity
                         cut is the kernels' score floor (nk_index_set_min_score): it seeds every query's running threshold,
                         so it prunes inside the scan instead of filtering a host-side list.
  NodeVectorIndex        the scoring loop of CALL db.index.vector.queryNodes (pkg/cypher/call_vector.go:177-256) over a
                         RESIDENT corpus of chunk embeddings: every node owns one or more chunk rows, its score is the BEST
                         over its chunks under the index's similarity function (cosine | dot | euclidean surfaced as
                         1/(1+d), similarity.go:152-158), nodes whose best score is negative are dropped (:240-242), the
                         label filter (:177-193) is a row bitmask, top k nodes returned.  One nk_search_groups call: a
                         per-node atomic max inside the scan kernel (segment-max) + a device top-k over the node keys — no
                         upload per query, no host over-select loop.
  query_nodes            the functional form (builds a NodeVectorIndex for one query), kept for the reference's call shape.
Host logic here is only id maps and the float64 / 1/(1+d) surfacing of the returned scores."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

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

    def AddBatch(self, ids: Sequence[str], vecs) -> None:
        """Bulk load: ONE device append for all new ids (Add is one synchronous append per vector), in-place updates for
        ids already present.  The reference loads an index by calling Add in a loop (search.go BuildIndexes)."""
        vs = np.asarray(vecs, dtype=np.float32).reshape(len(ids), -1)
        if vs.shape[1] != self.dimensions:
            raise ErrDimensionMismatch()
        fresh_ids, fresh_rows = [], []
        seen: Dict[str, int] = {}
        for i, v in zip(ids, vs):
            v = normalize(v)
            if i in self._pos:
                self._ix.update_row(self._pos[i], v)
            elif i in seen:
                fresh_rows[seen[i]] = v  # last write wins, like repeated Add calls
            else:
                seen[i] = len(fresh_ids)
                fresh_ids.append(i)
                fresh_rows.append(v)
        if fresh_ids:
            self._ix.append(np.stack(fresh_rows))
            for i in fresh_ids:
                self._pos[i] = len(self._ids)
                self._ids.append(i)

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
        self._ix.set_min_score(float(minSimilarity))  # `sim >= minSimilarity` (:339) evaluated inside the scan kernels
        try:
            idx, sc = self._ix.search(normalize(q), min(limit, len(self._ids)))
        finally:
            self._ix.set_min_score(None)
        out = []
        for i, s in zip(idx[0].tolist(), sc[0].tolist()):
            if i == 0xFFFFFFFF:  # fewer than `limit` rows reach the floor: unused slots
                break
            out.append((self._ids[i], float(s)))
        return out

    def Release(self) -> None:
        self._ix.release()


class NodeVectorIndex:
    """Chunk embeddings of a set of nodes, resident in HBM, for db.index.vector.queryNodes."""

    def __init__(self, dimensions: int, similarity: str = "cosine", devices: Sequence[int] = (0,)):
        self.dimensions = int(dimensions)
        self.metric = {"euclidean": "euclidean", "dot": "dot"}.get(similarity, "cosine")  # default cosine (call_vector.go:230)
        self._ix = KnnIndex(self.dimensions, metric=self.metric, devices=devices)
        self._node_ids: List[str] = []
        self._owner = np.empty(0, dtype=np.uint32)
        self._labels: List[Sequence[str]] = []

    def Load(self, node_chunks: Sequence[Tuple], labels: Optional[Sequence[Sequence[str]]] = None) -> None:
        """node_chunks: [(node_id, [chunk embedding, ...]), ...]; labels[i] = the labels of node i (optional)."""
        rows, owner = [], []
        self._node_ids = [nc[0] for nc in node_chunks]
        self._labels = list(labels) if labels is not None else [()] * len(node_chunks)
        for ni, nc in enumerate(node_chunks):
            for c in nc[1]:
                c = np.asarray(c, dtype=np.float32).reshape(-1)
                if c.size == self.dimensions:  # chunks of another dimension are skipped (call_vector.go:216-218)
                    rows.append(c)
                    owner.append(ni)
        self._owner = np.asarray(owner, dtype=np.uint32)
        if rows:
            self._ix.upload(np.stack(rows))
            self._ix.set_row_groups(self._owner, len(self._node_ids))

    def Query(self, query, k: int, label: str = "") -> List[Tuple[str, float]]:
        q = np.asarray(query, dtype=np.float32).reshape(-1)
        if q.size != self.dimensions:
            raise ErrDimensionMismatch()
        if k <= 0 or self._owner.size == 0:
            return []
        if label:  # label filter of the index (call_vector.go:177-193) as a row bitmask
            node_ok = np.fromiter((label in ls for ls in self._labels), dtype=bool, count=len(self._labels))
            self._ix.set_row_mask(node_ok[self._owner])
        else:
            self._ix.set_row_mask(None)
        # "bestScore >= 0" (call_vector.go:243); euclidean similarity 1/(1+d) is never negative
        self._ix.set_min_score(None if self.metric == "euclidean" else 0.0)
        try:
            grp, _, sc = self._ix.search_groups(q, min(int(k), len(self._node_ids)))
        finally:
            self._ix.set_min_score(None)
        out = []
        for g, s in zip(grp.tolist(), sc.tolist()):
            s = 1.0 / (1.0 + float(s)) if self.metric == "euclidean" else float(s)
            out.append((self._node_ids[g], s))
        return out

    def Release(self) -> None:
        self._ix.release()


def query_nodes(node_chunks: Sequence[Tuple[str, Sequence[Sequence[float]]]], query, k: int, similarity: str = "cosine",
                devices: Sequence[int] = (0,)) -> List[Tuple[str, float]]:
    """node_chunks: [(node_id, [chunk embedding, ...]), ...] -> top-k [(node_id, best-of-chunks score float64)]."""
    q = np.asarray(query, dtype=np.float32).reshape(-1)
    if k <= 0 or not node_chunks:
        return []
    nv = NodeVectorIndex(q.size, similarity, devices)
    try:
        nv.Load(node_chunks)
        return nv.Query(q, k)
    finally:
        nv.Release()
