"""Host logic of the reference-facing mirrors (EmbeddingIndex, VectorIndex / queryNodes, ClusterIndex) on CPU: the device
index is replaced by tests/fake_knn.py (oracle-backed), so id maps, swap-remove bookkeeping, serialisation, over-selection,
best-of-chunks and the k-means control flow are exercised by `pytest -m "not gpu"` as well.  The reference's own test
cases for these types (gpu_test.go, search_test.go, kmeans_test.go) are replayed from tests/golden/reference_kats.json."""
import struct

import numpy as np
import pytest

from fake_knn import FakeKnnIndex


@pytest.fixture()
def fake_device(monkeypatch, oracle_mod):
    import nornicdb_b200.embedding_index as ei
    import nornicdb_b200.vector_index as vi
    monkeypatch.setattr(ei, "KnnIndex", FakeKnnIndex)
    monkeypatch.setattr(vi, "KnnIndex", FakeKnnIndex)
    return oracle_mod


def test_embedding_index_reference_cases(fake_device, kats):
    from nornicdb_b200.embedding_index import EmbeddingIndex, ErrInvalidDimensions
    for t in kats:
        if t["op"] == "gpu.embedding_index_search":   # gpu_test.go:496-533
            ei = EmbeddingIndex(len(t["query"]))
            for nid, v in zip(t["ids"], t["vectors"]):
                ei.Add(nid, v)
            res = ei.Search(t["query"], t["k"])
            assert [r.ID for r in res][:len(t["want_ids"])] == t["want_ids"]
        if t["op"] == "gpu.score_subset":               # gpu_test.go:1592-1621
            ei = EmbeddingIndex(len(t["query"]))
            for nid, v in zip(t["ids"], t["vectors"]):
                ei.Add(nid, v)
            assert [r.ID for r in ei.ScoreSubset(t["query"], t["subset"])] == t["want_ids"]
    ei = EmbeddingIndex(4)
    assert ei.Search([1, 0, 0, 0], 3) is None            # empty index -> nil, nil (gpu.go:1540-1542)
    # lifecycle of gpu_test.go:1403-1480: SyncToGPU sets GPUSynced, Add / Remove clear it, double Release is safe
    life = EmbeddingIndex(4)
    life.Add("a", [1, 0, 0, 0]); life.Add("b", [0, 1, 0, 0]); life.Add("c", [0.9, 0.1, 0, 0])
    assert not life.Stats().GPUSynced
    life.SyncToGPU()
    assert life.Stats().GPUSynced and len(life.Search([1, 0, 0, 0], 2)) == 2
    life.Add("d", [0, 0, 1, 0])
    assert not life.Stats().GPUSynced
    life.SyncToGPU(); life.Remove("d")
    assert not life.IsGPUSynced()
    life.Release(); life.Release()
    with pytest.raises(ErrInvalidDimensions):
        ei.Add("x", [1, 2, 3])
    ei.AddBatch(["a", "b", "c", "a"], [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])  # duplicate id: last wins
    assert ei.Count() == 3 and np.allclose(ei.Get("a")[0], [0, 0, 0, 1])
    assert ei.Remove("a") and not ei.Remove("a") and ei.Count() == 2     # swap-with-last (gpu.go:1437-1471)
    assert ei.nodeIDs == ["c", "b"] and ei.idToIndex == {"c": 0, "b": 1}
    assert [r.ID for r in ei.Search([0, 0, 1, 0], 5)] == ["c", "b"]       # k > n -> n results
    blob = ei.Serialize()
    assert struct.unpack_from("<II", blob, 0) == (4, 2)
    other = EmbeddingIndex(4)
    other.Deserialize(blob)
    assert other.nodeIDs == ei.nodeIDs and other.Serialize() == blob
    with pytest.raises(ErrInvalidDimensions):
        EmbeddingIndex(5).Deserialize(blob)
    ei.Clear()
    assert ei.Count() == 0 and not ei.Has("b")


def test_vector_index_and_query_nodes(fake_device, kats):
    import nornicdb_b200.vector_index as vi
    for t in kats:
        if t["op"] == "search.vector_index":            # search_test.go:25-52
            ix = vi.VectorIndex(len(t["query"]))
            for nid, v in zip(t["ids"], t["vectors"]):
                ix.Add(nid, v)
            res = ix.Search(t["query"], t["limit"], t["min_similarity"])
            assert [r[0] for r in res] == t["want_ids"]
            assert abs(res[0][1] - t["want_first_score"]) <= t["tol"]
            ix.Remove(t["ids"][0])
            assert not ix.HasVector(t["ids"][0]) and ix.Count() == len(t["ids"]) - 1
        if t["op"] == "cypher.query_nodes_score":       # vector_procedures_test.go:538-572
            got = vi.query_nodes([("n", [t["stored"]])], t["query"], 1)
            assert got[0][0] == "n" and got[0][1] > t["want_gt"]
    # best-of-chunks per node, chunks of another dimension skipped, euclidean surfaced as 1/(1+d), negative best dropped
    nodes = [("a", [[1, 0, 0], [0, 1, 0]]), ("b", [[0.9, 0.1, 0]]), ("c", [[1, 0]]), ("d", [[-1, 0, 0]])]
    got = vi.query_nodes(nodes, [1, 0, 0], 10)
    assert [g[0] for g in got] == ["a", "b"] and abs(got[0][1] - 1.0) < 1e-6
    got = vi.query_nodes(nodes, [1, 0, 0], 10, similarity="euclidean")
    assert [g[0] for g in got][:2] == ["a", "b"] and abs(got[0][1] - 1.0) < 1e-6 and abs(got[-1][1] - 1.0 / 3.0) < 1e-6
    assert vi.query_nodes(nodes, [1, 0, 0], 0) == [] and vi.query_nodes([], [1, 0, 0], 3) == []


def test_cluster_index_control_flow(fake_device, kats):
    from nornicdb_b200.cluster_index import ClusterIndex, ErrInvalidDimensions, KMeansConfig
    rng = np.random.default_rng(0)
    mu = rng.uniform(-1, 1, (6, 16)).astype(np.float32)
    lab = rng.integers(0, 6, 600)
    rows = (mu[lab] + rng.standard_normal((600, 16)).astype(np.float32) * 0.05).astype(np.float32)
    ci = ClusterIndex(16, KMeansConfig(NumClusters=6, AutoK=False), rng=np.random.default_rng(1))
    ci.Cluster()                                             # empty: no-op
    assert not ci.IsClustered() and ci.SearchWithClusters(rows[0], 3, 2) is None
    ci.AddBatch([f"n{i}" for i in range(600)], rows)
    ci.Cluster()
    assert ci.IsClustered() and ci.NumClusters() == 6 and 1 <= ci.iterations <= 100
    # Lloyd fixed point: every row sits with its nearest centroid, every centroid is the mean of its members
    d = ((rows[:, None, :].astype(np.float64) - ci.centroids[None].astype(np.float64)) ** 2).sum(-1)
    assert (ci.assignments == d.argmin(1)).all()
    for c, members in ci.clusterMap.items():
        assert np.allclose(ci.centroids[c], rows[members].astype(np.float64).mean(0), rtol=1e-5, atol=1e-6)
    assert sorted(ci.GetClusterMembers(range(6))) == list(range(600))
    assert ci.FindNearestCentroid(rows[7]) == ci.assignments[7]
    near = ci.FindNearestClusters(rows[7], 99)
    assert len(near) == 6 and near[0] == ci.assignments[7]
    res = ci.SearchWithClusters(rows[7], 5, 1)
    assert res[0].ID == "n7" and abs(res[0].Score - 1.0) < 1e-5 and len(res) == 5
    st = ci.ClusterStats()
    assert st.Clustered and st.NumClusters == 6 and st.MinClusterSize >= 1 and st.MaxClusterSize <= 600
    # the reference's SearchCandidates cases (kmeans_test.go:744-800)
    for t in kats:
        if t["op"] != "kmeans.search_candidates":
            continue
        k = ClusterIndex(t["dims"])
        for i in range(t["n"]):
            emb = np.zeros(t["dims"], dtype=np.float32)
            emb[i % t["dims"]] = float(i)
            k.Add("node-" + chr(ord("A") + i), emb)
        if "want_error" in t:
            with pytest.raises(ErrInvalidDimensions):
                k.SearchCandidates(t["query"], t["candidates"], t["topk"])
        else:
            assert len(k.SearchCandidates(t["query"], t["candidates"], t["topk"]) or []) == t["want_len"]
    # real-time updates (kmeans.go:910-1052)
    ci.OnNodeUpdate("n3", rows[int(np.nonzero(lab != lab[3])[0][0])])
    assert ci.assignments[3] == ci.FindNearestCentroid(ci.Get("n3")[0]) and ci.updatesSinceCluster == 1
    ci.OnNodeUpdate("brand-new", rows[5])
    assert len(ci.assignments) == 601 and 600 in ci.clusterMap[int(ci.assignments[600])]
    assert not ci.ShouldRecluster()
    ci.UpdateCentroidsBatch()
    assert ci.pendingUpdates == []
    for i in range(70):
        ci.OnNodeUpdate(f"n{i + 10}", rows[i])
    assert ci.ShouldRecluster()                              # > 10 % of the corpus updated (kmeans.go:988-992)
    ci.Clear()
    assert not ci.IsClustered() and ci.Count() == 0
