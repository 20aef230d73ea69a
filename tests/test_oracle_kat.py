"""Pins the CPU oracle against every known-answer vector the reference's own tests hold for the kNN path
(tests/golden/reference_kats.json, transcribed from the cited reference test files)."""
import math

import numpy as np
import pytest


def _close(got, want, tol):
    return abs(got - want) <= tol if tol else got == want


def test_simd_scalar_kats(kats, oracle_mod):
    o = oracle_mod
    fns = {"simd.dot": o.dot, "simd.cosine": o.cosine, "simd.euclid": o.euclid, "gpu.cosine_flat": o.cosine_flat,
           "vector.cosine64": o.vec_cosine64, "vector.dot": o.vec_dot, "vector.euclid_sim": o.vec_euclid_sim}
    n = 0
    for t in kats:
        if t["op"] in fns:
            got = fns[t["op"]](t["a"], t["b"])
            assert _close(got, t["want"], t["tol"]), (t, got)
            n += 1
    assert n >= 45


def test_vector_cosine64_exact_doc_value(oracle_mod):
    # pkg/math/vector/similarity.go:32 documents 0.9746318461970762 for ([1,2,3],[4,5,6]).
    assert oracle_mod.vec_cosine64([1, 2, 3], [4, 5, 6]) == 0.9746318461970762


def test_simd_fast_variant_agrees_on_kats(kats, oracle_mod):
    """The AVX2 -ffast-math restatement (timing baseline) meets the same KATs within the reference's eps."""
    import ctypes as C
    L = oracle_mod.lib()
    fns = {"simd.dot": L.sb_dot, "simd.cosine": L.sb_cosine, "simd.euclid": L.sb_euclid}
    for t in kats:
        if t["op"] in fns and len(t["a"]) == len(t["b"]) and len(t["a"]) > 0:
            a = np.asarray(t["a"], np.float32)
            b = np.asarray(t["b"], np.float32)
            got = fns[t["op"]](a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), a.size)
            assert abs(got - t["want"]) <= max(t["tol"], 1e-5), (t, got)


def test_norm_and_normalize(kats, oracle_mod):
    o = oracle_mod
    for t in kats:
        if t["op"] == "simd.norm":
            assert _close(o.norm(t["v"]), t["want"], t["tol"])
        if t["op"] == "simd.normalize":
            v = np.asarray(t["v"], np.float32).copy()
            o.normalize_inplace(v)
            assert np.allclose(v, t["want"], atol=t["tol"])
            if o.norm(t["v"]) > 0:
                assert abs(o.norm(v) - 1.0) < 1e-5
            # vector.Normalize: zero vector -> zero vector (similarity.go:197-210)
            assert np.allclose(o.vec_normalize(t["v"]), t["want"], atol=1e-5)


def test_large_pattern_vectors(kats, oracle_mod):
    o = oracle_mod
    for t in kats:
        if t["op"] != "simd.large_pattern":
            continue
        size = t["size"]
        i = np.arange(size)
        a = ((i % 10) / 10.0).astype(np.float32)
        b = (((i + 5) % 10) / 10.0).astype(np.float32)
        d, c, e, nr = o.dot(a, b), o.cosine(a, b), o.euclid(a, b), o.norm(a)
        assert math.isfinite(d) and -1.0 <= c <= 1.0 and e >= 0 and nr >= 0
        # independent fp64 check of the values themselves
        a64, b64 = a.astype(np.float64), b.astype(np.float64)
        assert abs(d - a64 @ b64) <= 1e-4 * max(1.0, abs(a64 @ b64))
        assert abs(c - (a64 @ b64) / math.sqrt((a64 @ a64) * (b64 @ b64))) <= 1e-5
        assert abs(e - math.sqrt(((a64 - b64) ** 2).sum())) <= 1e-4 * max(1.0, e)


def test_edge_cases(kats, oracle_mod):
    o = oracle_mod
    for t in kats:
        if t["op"] == "simd.cosine_not_nan":
            assert not math.isnan(o.cosine(t["a"], t["b"]))
        if t["op"] == "kmeans.squared_euclidean":
            assert abs(o.euclid(t["a"], t["b"]) ** 2 - t["want"]) <= 1e-3
            assert abs(o.sq_euclid64(t["a"], t["b"]) - t["want"]) <= t["tol"]
        if t["op"] == "cypher.query_nodes_score":
            assert o.vec_cosine64(t["stored"], t["query"]) > t["want_gt"]


def test_topk_and_partial_sort(kats, oracle_mod):
    o = oracle_mod
    for t in kats:
        if t["op"] == "cuda.topk":
            idx, sc = o.topk_insertion(t["scores"], t["k"])
            assert idx.tolist() == t["want_idx"]
            assert np.allclose(sc, t["want_scores"])
        if t["op"] == "gpu.partial_sort":
            idx = o.partial_sort(t["scores"], t["k"])
            assert [t["scores"][i] for i in idx[: t["k"]]] == t["want_scores"]
    # ties: lowest index first (strict '>' forward scan, cuda_bridge.go:356-371)
    idx, _ = o.topk_insertion([0.5, 0.9, 0.5, 0.9, 0.1], 4)
    assert idx.tolist() == [1, 3, 0, 2]
    # k > n clamps (cuda_bridge.go:332-334); k == 0 -> nothing
    idx, _ = o.topk_insertion([0.3, 0.1], 10)
    assert idx.tolist() == [0, 1]
    idx, _ = o.topk_insertion([0.3, 0.1], 0)
    assert idx.size == 0
    # partialSort k >= n -> full sort (gpu_test.go:1111-1145)
    idx = o.partial_sort([0.1, 0.9, 0.5], 5)
    assert idx.tolist() == [1, 2, 0]


def test_batch_shapes(oracle_mod):
    o = oracle_mod
    emb = np.array([1, 0, 0, 0, 1, 0, 0.6, 0.8, 0], np.float32)
    q = np.array([1, 0, 0], np.float32)
    assert np.allclose(o.batch("cosine", emb, q), [1.0, 0.0, 0.6], atol=1e-6)
    assert np.allclose(o.batch("dot", emb, q), [1.0, 0.0, 0.6], atol=1e-6)
    assert np.allclose(o.batch("euclid", emb, q), [0.0, math.sqrt(2), math.sqrt(0.16 + 0.64)], atol=1e-6)
    # scores slice too short -> silently untouched (simd.go:155-157)
    assert np.isnan(o.batch("cosine", emb, q, n_scores=2)).all()
    v = np.array([3, 4, 0, 0, 0, 0], np.float32)
    o.batch_normalize(v, 2, 3)
    assert np.allclose(v, [0.6, 0.8, 0, 0, 0, 0])


def test_exact_knn_matches_numpy_and_simd(oracle_mod):
    o = oracle_mod
    rows = o.fill_uniform(3000, 96, 42)
    q = o.fill_uniform(5, 96, 1337)
    r64, q64 = rows.astype(np.float64), q.astype(np.float64)
    for metric in ("cosine", "dot", "euclidean"):
        idx, sc = o.knn_exact64(rows, q, 7, metric)
        if metric == "cosine":
            full = (q64 @ r64.T) / np.linalg.norm(q64, axis=1)[:, None] / np.linalg.norm(r64, axis=1)[None, :]
        elif metric == "dot":
            full = q64 @ r64.T
        else:
            full = -np.sqrt(((q64[:, None, :] - r64[None, :, :]) ** 2).sum(-1))
        want = np.argsort(-full, axis=1, kind="stable")[:, :7]
        assert (idx == want).all()
        ref = np.take_along_axis(full, want, 1)
        assert np.allclose(sc, np.abs(ref) if metric == "euclidean" else ref, rtol=1e-12, atol=1e-12)
        for threads in (1, 3):
            si, ss = o.simd_knn(rows, q, 7, metric, threads=threads)
            assert (si == idx).all()
            assert np.allclose(ss, sc, rtol=1e-4, atol=1e-6)


def test_generator_is_counter_based(oracle_mod):
    o = oracle_mod
    a = o.fill_uniform(10, 16, 7)
    b = o.fill_uniform(4, 16, 7, row_base=6)
    assert (a[6:] == b).all()
    assert a.min() >= -1.0 and a.max() < 1.0
    h = o.fill_uniform(10, 16, 7, dtype="f16")
    assert (h == a.astype(np.float16)).all()


def test_kmeans_restatement(kats, oracle_mod):
    """pkg/gpu/kmeans.go: optimalK KATs (host mirror), assignment / update steps against plain float64 numpy."""
    from nornicdb_b200.cluster_index import optimalK
    seen = 0
    for t in kats:
        if t["op"] == "kmeans.optimal_k":
            assert t["want_min"] <= optimalK(t["n"]) <= t["want_max"], t
            seen += 1
    assert seen == 6
    rng = np.random.default_rng(0)
    rows = rng.uniform(-1, 1, (500, 24)).astype(np.float32)
    cen = rng.uniform(-1, 1, (7, 24)).astype(np.float32)
    a = np.zeros(500, dtype=np.int32)
    changed = oracle_mod.kmeans_assign(rows, cen, a)
    d = ((rows[:, None, :].astype(np.float64) - cen[None].astype(np.float64)) ** 2).sum(-1)
    assert (a == d.argmin(1)).all() and changed == int((d.argmin(1) != 0).sum())
    assert oracle_mod.kmeans_assign(rows, cen, a) == 0
    ac = np.zeros(500, dtype=np.int32)
    oracle_mod.kmeans_assign(rows, cen, ac, by_cosine=True)
    cos = (rows @ cen.T) / np.sqrt((rows * rows).sum(1)[:, None] * (cen * cen).sum(1)[None])
    assert (ac == cos.argmax(1)).mean() > 0.995
    new, counts = oracle_mod.kmeans_update(rows, a, cen)
    for c in range(7):
        assert counts[c] == (a == c).sum()
        if counts[c]:
            assert np.allclose(new[c], rows[a == c].astype(np.float64).mean(0), rtol=1e-6, atol=1e-7)
