"""Front-end semantics of the callers over the fused kernel: search.VectorIndex and the scoring loop
of CALL db.index.vector.queryNodes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_vector_index_basic_kat(knn_lib, kats):
    # pkg/search/search_test.go:25-52 TestVectorIndex_Basic
    from nornicdb_b200.vector_index import ErrDimensionMismatch, VectorIndex
    t = [t for t in kats if t["op"] == "search.vector_index"][0]
    idx = VectorIndex(4)
    for i, v in zip(t["ids"], t["vectors"]):
        idx.Add(i, v)
    assert idx.Count() == 3 and idx.HasVector("doc1") and not idx.HasVector("doc99")
    res = idx.Search(t["query"], t["limit"], t["min_similarity"])
    assert [r[0] for r in res] == t["want_ids"]  # doc3 is orthogonal: below the threshold
    assert abs(res[0][1] - t["want_first_score"]) <= t["tol"] and isinstance(res[0][1], float)
    with pytest.raises(ErrDimensionMismatch):
        idx.Add("bad", [1, 2, 3])
    with pytest.raises(ErrDimensionMismatch):
        idx.Search([1, 0], 5, 0.0)
    idx.Remove("doc1")
    assert [r[0] for r in idx.Search(t["query"], 10, 0.5)] == ["doc2"]
    idx.Release()


def test_vector_index_matches_cpu_semantics(knn_lib, oracle_mod):
    from nornicdb_b200.vector_index import VectorIndex
    rows = oracle_mod.fill_uniform(3000, 96, 7) * 3.0  # un-normalised inputs: Add normalises (vector_index.go:234)
    q = oracle_mod.fill_uniform(1, 96, 8)[0] * 0.5
    idx = VectorIndex(96)
    for i, v in enumerate(rows):
        idx.Add(f"d{i}", v)
    res = idx.Search(q, 25, 0.05)
    cos = np.array([oracle_mod.vec_cosine64(r, q) for r in rows])
    order = np.argsort(-cos, kind="stable")
    want = [(f"d{i}", cos[i]) for i in order[:25] if cos[i] >= 0.05]
    assert [r[0] for r in res] == [w[0] for w in want]
    assert np.allclose([r[1] for r in res], [w[1] for w in want], rtol=1e-4, atol=1e-6)
    idx.Release()


def test_query_nodes_best_of_chunks(knn_lib, kats, oracle_mod):
    from nornicdb_b200.vector_index import query_nodes
    # pkg/cypher/vector_procedures_test.go:538-572: one stored embedding, score > 0.9
    t = [t for t in kats if t["op"] == "cypher.query_nodes_score"][0]
    res = query_nodes([("n", [t["stored"]])], t["query"], 5, "cosine")
    assert len(res) == 1 and res[0][0] == "n" and res[0][1] > t["want_gt"]
    # randomised: nodes with 1..4 chunks, all three similarity functions, against the reference formulas in fp64
    rng = np.random.default_rng(2)
    chunks = oracle_mod.fill_uniform(900, 32, 11)
    nodes, at = [], 0
    while at < 900:
        c = int(rng.integers(1, 5))
        nodes.append((f"node{len(nodes)}", chunks[at:at + c]))
        at += c
    q = oracle_mod.fill_uniform(1, 32, 12)[0]
    for sim in ("cosine", "dot", "euclidean"):
        got = query_nodes(nodes, q, 15, sim)
        ref = []
        for nid, cs in nodes:
            if sim == "cosine":
                best = max(oracle_mod.vec_cosine64(c, q) for c in cs)
            elif sim == "dot":
                best = max(float(np.dot(c.astype(np.float64), q.astype(np.float64))) for c in cs)
            else:
                best = max(1.0 / (1.0 + float(np.sqrt(((c.astype(np.float64) - q) ** 2).sum()))) for c in cs)
            if best >= 0.0:
                ref.append((nid, best))
        ref.sort(key=lambda x: -x[1])
        assert [g[0] for g in got] == [r[0] for r in ref[:15]], sim
        assert np.allclose([g[1] for g in got], [r[1] for r in ref[:15]], rtol=1e-4, atol=1e-6)
    # nodes whose best score is negative are dropped (call_vector.go:240-242)
    res = query_nodes([("pos", [[1, 0]]), ("neg", [[-1, 0]])], [1, 0], 5, "cosine")
    assert [r[0] for r in res] == ["pos"]


def test_node_vector_index_resident_labels_and_batch_add(knn_lib, oracle_mod):
    """queryNodes over a RESIDENT chunk corpus: label filter = row bitmask, best-of-chunks = device segment-max, the
    `bestScore >= 0` rule = the kernels' score floor; VectorIndex.AddBatch = one device append."""
    from nornicdb_b200.vector_index import NodeVectorIndex, VectorIndex
    rng = np.random.default_rng(4)
    d, n_nodes = 48, 700
    chunks = oracle_mod.fill_uniform(2500, d, 21)
    nodes, labels, at = [], [], 0
    for i in range(n_nodes):
        c = int(rng.integers(1, 6))
        nodes.append((f"n{i}", chunks[at:at + c]))
        labels.append(("Doc",) if i % 3 else ("Doc", "Memory"))
        at += c
    q = oracle_mod.fill_uniform(1, d, 22)[0]
    for sim in ("cosine", "dot", "euclidean"):
        nv = NodeVectorIndex(d, sim)
        nv.Load(nodes, labels)
        for label in ("", "Memory"):
            got = nv.Query(q, 12, label)
            ref = []
            for (nid, cs), ls in zip(nodes, labels):
                if label and label not in ls:
                    continue
                x = cs.astype(np.float64)
                if sim == "cosine":
                    best = max(oracle_mod.vec_cosine64(c, q) for c in cs)
                elif sim == "dot":
                    best = float((x @ q.astype(np.float64)).max())
                else:
                    best = float((1.0 / (1.0 + np.sqrt(((x - q) ** 2).sum(1)))).max())
                if best >= 0.0:
                    ref.append((nid, best))
            ref.sort(key=lambda t: -t[1])
            assert [g[0] for g in got] == [r[0] for r in ref[:12]], (sim, label)
            assert np.allclose([g[1] for g in got], [r[1] for r in ref[:12]], rtol=1e-4, atol=1e-6)
        nv.Release()
    # AddBatch == Add in a loop
    rows = oracle_mod.fill_uniform(500, d, 23)
    a, b = VectorIndex(d), VectorIndex(d)
    for i, v in enumerate(rows):
        a.Add(f"d{i}", v)
    b.AddBatch([f"d{i}" for i in range(500)], rows)
    b.AddBatch(["d3", "new"], [rows[7], rows[9]])  # update in place + one fresh id
    a.Add("d3", rows[7]); a.Add("new", rows[9])
    assert a.Search(q, 20, 0.1) == b.Search(q, 20, 0.1)
    a.Release(); b.Release()
