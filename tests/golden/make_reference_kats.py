#!/usr/bin/env python
"""Transcribes the known-answer vectors of the reference's OWN tests for the kNN hot path into
tests/golden/reference_kats.json.  Each entry cites the reference file:line it was read
from.  The reference is Go and cannot be executed here (no toolchain), so these literals — not outputs
of a run — are the golden data; re-run this script after editing to regenerate the JSON.

    python tests/golden/make_reference_kats.py
"""
import json
import os

K = []


def add(op, src, **kw):
    kw.update(op=op, src=src)
    K.append(kw)


# ---- pkg/simd/simd_test.go:14-79 TestDotProduct (eps 1e-5)
S = "pkg/simd/simd_test.go:14-79"
add("simd.dot", S, a=[1, 2, 3], b=[4, 5, 6], want=32, tol=1e-5)
add("simd.dot", S, a=[0, 0, 0], b=[0, 0, 0], want=0, tol=1e-5)
add("simd.dot", S, a=[], b=[], want=0, tol=1e-5)
add("simd.dot", S, a=[1, 0, 0], b=[0, 1, 0], want=0, tol=1e-5)
add("simd.dot", S, a=[3, 4], b=[3, 4], want=25, tol=1e-5)
add("simd.dot", S, a=[-1, -2, -3], b=[4, 5, 6], want=-32, tol=1e-5)
add("simd.dot", S, a=[1] * 256, b=[1] * 256, want=256, tol=1e-5)

# ---- pkg/simd/simd_test.go:81-140 TestCosineSimilarity
S = "pkg/simd/simd_test.go:81-140"
add("simd.cosine", S, a=[1, 2, 3], b=[1, 2, 3], want=1.0, tol=1e-5)
add("simd.cosine", S, a=[1, 0, 0], b=[-1, 0, 0], want=-1.0, tol=1e-5)
add("simd.cosine", S, a=[1, 0, 0], b=[0, 1, 0], want=0.0, tol=1e-5)
add("simd.cosine", S, a=[0, 0, 0], b=[1, 2, 3], want=0.0, tol=1e-5)
add("simd.cosine", S, a=[1, 2, 3], b=[0, 0, 0], want=0.0, tol=1e-5)
add("simd.cosine", S, a=[], b=[], want=0.0, tol=1e-5)
add("simd.cosine", S, a=[1, 2, 3], b=[2, 4, 6], want=1.0, tol=1e-5)

# ---- pkg/simd/simd_test.go:142-189 TestEuclideanDistance
S = "pkg/simd/simd_test.go:142-189"
add("simd.euclid", S, a=[0, 0], b=[3, 4], want=5.0, tol=1e-5)
add("simd.euclid", S, a=[1, 2, 3], b=[1, 2, 3], want=0.0, tol=1e-5)
add("simd.euclid", S, a=[], b=[], want=0.0, tol=1e-5)
add("simd.euclid", S, a=[0, 0, 0], b=[1, 0, 0], want=1.0, tol=1e-5)
add("simd.euclid", S, a=[-3, -4], b=[0, 0], want=5.0, tol=1e-5)

# ---- pkg/simd/simd_test.go:191-232 TestNorm
S = "pkg/simd/simd_test.go:191-232"
add("simd.norm", S, v=[3, 4], want=5.0, tol=1e-5)
add("simd.norm", S, v=[1, 0, 0], want=1.0, tol=1e-5)
add("simd.norm", S, v=[0, 0, 0], want=0.0, tol=1e-5)
add("simd.norm", S, v=[], want=0.0, tol=1e-5)
add("simd.norm", S, v=[-3, -4], want=5.0, tol=1e-5)

# ---- pkg/simd/simd_test.go:234-278 TestNormalizeInPlace
S = "pkg/simd/simd_test.go:234-278"
add("simd.normalize", S, v=[3, 4], want=[0.6, 0.8], tol=1e-5)
add("simd.normalize", S, v=[1, 0, 0], want=[1, 0, 0], tol=1e-5)
add("simd.normalize", S, v=[0, 0, 0], want=[0, 0, 0], tol=1e-5)

# ---- pkg/simd/simd_test.go:304-340 TestLargeVectors: pattern vectors, results finite and in range
for size in [16, 32, 64, 128, 256, 512, 768, 1024, 1536]:
    add("simd.large_pattern", "pkg/simd/simd_test.go:304-340", size=size)

# ---- pkg/simd/simd_test.go:343-381 TestEdgeCases
for size in range(1, 18):
    add("simd.dot", "pkg/simd/simd_test.go:345-359", a=[1.0] * size, b=[1.0] * size, want=size, tol=1e-5)
add("simd.cosine_not_nan", "pkg/simd/simd_test.go:361-369", a=[1e-20] * 3, b=[1e-20] * 3)
add("simd.cosine", "pkg/simd/simd_test.go:371-380", a=[1e5] * 3, b=[1e5] * 3, want=1.0, tol=1e-3)

# ---- pkg/math/vector/similarity_test.go
add("vector.cosine64", "pkg/math/vector/similarity_test.go:38-43,111-121 (doc value similarity.go:32)",
    a=[1, 2, 3], b=[4, 5, 6], want=0.9746318461970762, tol=1e-3)
add("vector.dot", "pkg/math/vector/similarity_test.go:123-158", a=[1, 2, 3], b=[4, 5, 6], want=32.0, tol=0)
add("vector.dot", "pkg/math/vector/similarity_test.go:123-158", a=[1, 2, 3], b=[4, 5], want=0.0, tol=0)
add("vector.euclid_sim", "pkg/math/vector/similarity_test.go:160-192", a=[1, 2, 3], b=[1, 2, 3], want=1.0, tol=1e-9)
add("vector.euclid_sim", "pkg/math/vector/similarity_test.go:160-192", a=[0, 0], b=[3, 4], want=1.0 / 6.0, tol=1e-6)

# ---- pkg/gpu/cuda/cuda_test.go (real-CUDA tests of the boundary being replaced)
add("cuda.normalize_vectors", "pkg/gpu/cuda/cuda_test.go:215-253", data=[3, 4, 0, 1, 0, 0], n=2, dims=3,
    want=[0.6, 0.8, 0, 1, 0, 0], tol=1e-3)
add("cuda.cosine_similarity", "pkg/gpu/cuda/cuda_test.go:255-312",
    embeddings=[1, 0, 0, 0, 1, 0, 0.6, 0.8, 0], query=[1, 0, 0], n=3, dims=3, normalized=True,
    want=[1.0, 0.0, 0.6], tol=1e-3)
add("cuda.topk", "pkg/gpu/cuda/cuda_test.go:314-353", scores=[0.1, 0.8, 0.3, 0.9, 0.2], k=3, want_idx=[3, 1, 2],
    want_scores=[0.9, 0.8, 0.3])
add("cuda.search", "pkg/gpu/cuda/cuda_test.go:355-400",
    embeddings=[1, 0, 0, 0, 1, 0, 0, 0, 1, 0.6, 0.8, 0, 0.7, 0.7, 0.14], query=[0.6, 0.8, 0.0], n=5, dims=3, k=2,
    want_len=2, want_first_idx=3, want_first_score=1.0, tol=1e-3)
add("cuda.search_zero_k", "pkg/gpu/cuda/cuda_test.go:402-431", embeddings=[1, 0, 0], query=[1, 0, 0], n=1, dims=3, k=0)
add("cuda.search_k_gt_n", "pkg/gpu/cuda/cuda_test.go:433-462", embeddings=[1, 0, 0, 0, 1, 0], query=[1, 0, 0], n=2,
    dims=3, k=10, want_len=2)

# ---- pkg/gpu/gpu_test.go
add("gpu.embedding_index_search", "pkg/gpu/gpu_test.go:496-533",
    ids=["node-1", "node-2", "node-3", "node-4"],
    vectors=[[1, 0, 0, 0], [0, 1, 0, 0], [0.9, 0.1, 0, 0], [0, 0, 1, 0]], query=[1, 0, 0, 0], k=2,
    want_ids=["node-1", "node-3"], want_first_score_ge=0.99)
add("gpu.partial_sort", "pkg/gpu/gpu_test.go:841-857", scores=[0.1, 0.9, 0.5, 0.3, 0.7], k=3, want_scores=[0.9, 0.7, 0.5])
add("gpu.cosine_flat", "pkg/gpu/gpu_test.go:859-883,1088-1109", a=[1, 0, 0], b=[1, 0, 0], want=1.0, tol=1e-3)
add("gpu.cosine_flat", "pkg/gpu/gpu_test.go:859-883,1088-1109", a=[1, 0, 0], b=[0, 1, 0], want=0.0, tol=1e-3)
add("gpu.cosine_flat", "pkg/gpu/gpu_test.go:1088-1109", a=[1, 2, 3], b=[1, 2], want=0.0, tol=0)
add("gpu.cosine_flat", "pkg/gpu/gpu_test.go:1088-1109", a=[0, 0, 0], b=[0, 0, 0], want=0.0, tol=0)
add("gpu.score_subset", "pkg/gpu/gpu_test.go:1592-1621",
    ids=["a", "b", "c"], vectors=[[1, 0, 0], [0, 1, 0], [1, 1, 0]], query=[1, 0, 0],
    subset=["b", "c", "a", "missing"], want_ids=["a", "c", "b"])

# ---- pkg/search/search_test.go:25-52 (VectorIndex ordering + threshold)
add("search.vector_index", "pkg/search/search_test.go:25-52",
    ids=["doc1", "doc2", "doc3"], vectors=[[1, 0, 0, 0], [0.9, 0.1, 0, 0], [0, 1, 0, 0]], query=[1, 0, 0, 0],
    limit=10, min_similarity=0.5, want_ids=["doc1", "doc2"], want_first_score=1.0, tol=0.01)

# ---- pkg/gpu/kmeans_test.go:698-717 (next-row KAT: squared euclidean)
add("kmeans.squared_euclidean", "pkg/gpu/kmeans_test.go:698-717", a=[1, 2, 3], b=[4, 5, 6], want=27.0, tol=1e-4)
add("kmeans.squared_euclidean", "pkg/gpu/kmeans_test.go:698-717", a=[3, 4, 0], b=[0, 0, 0], want=25.0, tol=1e-4)
add("kmeans.squared_euclidean", "pkg/gpu/kmeans_test.go:698-717", a=[0, 0, 0], b=[0, 0, 0], want=0.0, tol=1e-4)
add("kmeans.squared_euclidean", "pkg/gpu/kmeans_test.go:698-717", a=[1, 0, 0], b=[0, 0, 0], want=1.0, tol=1e-4)
add("kmeans.squared_euclidean", "pkg/gpu/kmeans_test.go:698-717", a=[1, 1, 1], b=[0, 0, 0], want=3.0, tol=1e-4)

# ---- pkg/gpu/kmeans_test.go:675-696 TestOptimalK (k = clamp(sqrt(n/2), 10, 1000), kmeans.go:323-332)
for n, lo, hi in [(100, 10, 10), (200, 10, 10), (800, 10, 30), (2000, 20, 50), (10000, 50, 100), (2000000, 1000, 1000)]:
    add("kmeans.optimal_k", "pkg/gpu/kmeans_test.go:675-696", n=n, want_min=lo, want_max=hi)

# ---- pkg/gpu/kmeans_test.go:744-800 TestClusterIndex_SearchCandidates: 20 embeddings, emb_i[i % dims] = i (testDims = 64,
# kmeans_test.go:60), query e0; result length = min(topK, |candidates|), empty candidates -> 0, wrong dims -> error
S = "pkg/gpu/kmeans_test.go:744-800"
E0, Z = [1.0] + [0.0] * 63, [0.0] * 64
add("kmeans.search_candidates", S, dims=64, n=20, query=E0, candidates=[0, 1, 2, 3, 4], topk=3, want_len=3)
add("kmeans.search_candidates", S, dims=64, n=20, query=Z, candidates=[], topk=3, want_len=0)
add("kmeans.search_candidates", S, dims=64, n=20, query=Z, candidates=[0, 1, 2], topk=100, want_len=3)
add("kmeans.search_candidates", S, dims=64, n=20, query=Z + [0.0], candidates=[0, 1], topk=1, want_error="ErrInvalidDimensions")

# ---- pkg/cypher/vector_procedures_test.go:538-572
add("cypher.query_nodes_score", "pkg/cypher/vector_procedures_test.go:538-572", stored=[0.7, 0.2, 0.05, 0.05],
    query=[0.65, 0.25, 0.05, 0.05], want_gt=0.9)

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")
with open(out, "w") as f:
    json.dump(K, f, indent=1)
print(f"{len(K)} KATs -> {out}")
