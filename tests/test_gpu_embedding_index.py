"""gpu.EmbeddingIndex semantics on the device, mirroring pkg/gpu/gpu_test.go."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_add_and_search(knn_lib, kats):
    # pkg/gpu/gpu_test.go:496-533 TestEmbeddingIndexAddAndSearch
    from nornicdb_b200.embedding_index import EmbeddingIndex
    t = [t for t in kats if t["op"] == "gpu.embedding_index_search"][0]
    ei = EmbeddingIndex(4)
    for nid, v in zip(t["ids"], t["vectors"]):
        ei.Add(nid, v)
    assert ei.Count() == 4
    res = ei.Search(t["query"], t["k"])
    assert len(res) == 2
    assert [r.ID for r in res] == t["want_ids"]
    assert res[0].Score >= 0.99 and abs(res[0].Distance - (1 - res[0].Score)) < 1e-6
    ei.Release()


def test_invalid_dimensions_and_empty(knn_lib):
    from nornicdb_b200.embedding_index import EmbeddingIndex, ErrInvalidDimensions
    ei = EmbeddingIndex(3)
    with pytest.raises(ErrInvalidDimensions):
        ei.Add("a", [1, 2])  # gpu.go:1379-1381
    assert ei.Search([1, 0, 0], 5) is None  # empty index -> nil, nil (gpu.go:1540-1542)
    ei.Add("a", [1, 0, 0])
    with pytest.raises(ErrInvalidDimensions):
        ei.Search([1, 0], 1)  # gpu.go:1533-1535
    assert len(ei.Search([1, 0, 0], 10)) == 1  # k > n -> n
    ei.Release()


def test_update_remove_has_get(knn_lib, oracle_mod):
    # pkg/gpu/gpu_test.go:1403-1480 lifecycle + swap-with-last removal (gpu.go:1437-1471)
    from nornicdb_b200.embedding_index import EmbeddingIndex
    rows = oracle_mod.fill_uniform(200, 32, 5)
    ids = [f"n{i}" for i in range(200)]
    ei = EmbeddingIndex(32)
    ei.AddBatch(ids[:150], rows[:150])
    for i in range(150, 200):
        ei.Add(ids[i], rows[i])
    assert ei.Count() == 200 and ei.Has("n7") and not ei.Has("zz")
    assert not ei.IsGPUSynced() and not ei.Stats().GPUSynced   # "should not be synced after add" (gpu_test.go:1446-1454)
    ei.SyncToGPU()
    assert ei.IsGPUSynced() and ei.Stats().GPUSynced
    v, ok = ei.Get("n7")
    assert ok and (v == rows[7]).all()
    ei.Add("n7", rows[8] * 3.0)  # update in place -> now parallel to n8
    res = ei.Search(rows[8], 2)
    assert {r.ID for r in res} == {"n7", "n8"} and res[0].Score > 0.999999 and res[1].Score > 0.999999
    assert ei.Remove("n8") and not ei.Remove("n8")
    assert ei.Count() == 199 and not ei.Has("n8")
    # the last row (n199) moved into n8's slot; results still map to the right ids
    res = ei.Search(rows[199], 1)
    assert res[0].ID == "n199" and res[0].Score > 0.999999
    host = {nid: (rows[8] * 3.0 if nid == "n7" else rows[int(nid[1:])]) for nid in ids if nid != "n8"}
    q = oracle_mod.fill_uniform(1, 32, 77)[0]
    want = sorted(host, key=lambda nid: -oracle_mod.vec_cosine64(host[nid], q))[:5]
    assert [r.ID for r in ei.Search(q, 5)] == want
    ei.Release()


def test_score_subset(knn_lib, kats):
    # pkg/gpu/gpu_test.go:1592-1621
    from nornicdb_b200.embedding_index import EmbeddingIndex
    t = [t for t in kats if t["op"] == "gpu.score_subset"][0]
    ei = EmbeddingIndex(3)
    for nid, v in zip(t["ids"], t["vectors"]):
        ei.Add(nid, v)
    res = ei.ScoreSubset(t["query"], t["subset"])
    assert [r.ID for r in res] == t["want_ids"]
    assert ei.ScoreSubset(t["query"], []) is None and ei.ScoreSubset(t["query"], ["missing"]) is None
    ei.Release()


def test_serialize_roundtrip(knn_lib, oracle_mod):
    # gpu.go:2373-2454 blob: LE [dims][count][len-prefixed ids][fp32 vectors]
    import struct
    from nornicdb_b200.embedding_index import EmbeddingIndex, ErrInvalidDimensions
    rows = oracle_mod.fill_uniform(50, 16, 9)
    ei = EmbeddingIndex(16)
    ei.AddBatch([f"id-{i}" for i in range(50)], rows)
    blob = ei.Serialize()
    assert struct.unpack_from("<II", blob, 0) == (16, 50)
    assert blob[8:12] == struct.pack("<I", 4) and blob[12:16] == b"id-0"
    assert np.frombuffer(blob[-50 * 16 * 4:], "<f4").reshape(50, 16).tobytes() == rows.tobytes()
    other = EmbeddingIndex(16)
    other.Deserialize(blob)
    q = oracle_mod.fill_uniform(1, 16, 10)[0]
    assert [r.ID for r in other.Search(q, 5)] == [r.ID for r in ei.Search(q, 5)]
    with pytest.raises(ErrInvalidDimensions):
        EmbeddingIndex(8).Deserialize(blob)
    ei.Release(); other.Release()


def test_cold_start_feed_large_unaligned_blob_and_fp16_conversion(knn_lib, oracle_mod):
    """a serialized index whose fp32 payload starts at an odd byte offset and spans several 32 MB staging
    chunks loads bit-exactly through the pinned double buffer; an fp16 index converts on the device while loading."""
    import struct
    import time
    from nornicdb_b200.embedding_index import EmbeddingIndex
    n, d = 70_000, 384  # 107 MB of vectors: 4 staging chunks
    rows = oracle_mod.fill_uniform(n, d, 77)
    ids = [f"n{i}" for i in range(n)]
    blob = bytearray(struct.pack("<II", d, n))
    for s in ids:
        b = s.encode()
        blob += struct.pack("<I", len(b)) + b
    assert len(blob) % 4 != 0
    blob += rows.tobytes()
    blob = bytes(blob)
    ei = EmbeddingIndex(d)
    t0 = time.perf_counter()
    ei.Deserialize(blob)
    dt = time.perf_counter() - t0
    assert ei.Count() == n and ei.nodeIDs[-1] == ids[-1]
    for r in (0, 21_845, 43_690, n - 1):  # rows around the chunk boundaries
        assert ei._ix.read_rows(r, 1).tobytes() == rows[r:r + 1].tobytes()
    assert ei.Serialize() == blob
    q = oracle_mod.fill_uniform(1, d, 5)[0]
    want = [r.ID for r in ei.Search(q, 10)]
    h = EmbeddingIndex(d, dtype="f16")
    h.Deserialize(blob)  # fp32 blob -> fp16 rows, converted on the device
    assert h._ix.read_rows(12_345, 3).tobytes() == rows[12_345:12_348].astype(np.float16).tobytes()
    got = [r.ID for r in h.Search(q, 10)]
    assert len(set(got) & set(want)) >= 9  # fp16 rows: at most a boundary swap
    print(f"cold-start feed: {len(blob) / dt / 1e9:.2f} GB/s through Deserialize (ids parsed on the host)")
    ei.Release(); h.Release()


def test_search_batch_uses_one_fused_call(knn_lib, oracle_mod):
    from nornicdb_b200.embedding_index import EmbeddingIndex
    rows = oracle_mod.fill_uniform(3000, 64, 3)
    ei = EmbeddingIndex(64)
    ei.AddBatch([str(i) for i in range(3000)], rows)
    q = oracle_mod.fill_uniform(40, 64, 4)
    before = ei._ix.stats()["searches"]
    batch = ei.SearchBatch(q, 5)
    assert ei._ix.stats()["searches"] == before + 1
    oi, _ = oracle_mod.knn_exact64(rows, q, 5, "cosine")
    assert [[int(r.ID) for r in b] for b in batch] == oi.tolist()
    ei.Release()
