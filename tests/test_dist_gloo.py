"""N>1 path on CPU (world_size 2, gloo): shard plan + candidate-key codec + all-gather + merge give the same
answer as the unsharded oracle, for every metric, independent of how rows are partitioned.
The per-shard scan is played by the CPU oracle here; on a GPU box the same flow runs with the CUDA scan
(tests/test_gpu_knn.py::test_stats_and_device_api, bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nornicdb_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, d, Q, k, metric, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    lo, hi = sharding.shard_range(n, world, rank)
    rows = oracle.fill_uniform(hi - lo, d, 42, row_base=lo)  # this rank's rows of the global stream
    q = oracle.fill_uniform(Q, d, 1337)
    idx, sc = oracle.knn_exact64(rows, q, k, metric, row_base=lo)
    ke = idx.shape[1]
    keys = np.zeros((Q, k), dtype=np.uint64)
    s32 = sc.astype(np.float32)
    keys[:, :ke] = sharding.pack_keys(-(s32 * s32) if metric == "euclidean" else s32, idx)
    mine = torch.from_numpy(keys.view(np.int64).copy())
    gathered = torch.empty((world, Q, k), dtype=torch.int64)
    dist.all_gather_into_tensor(gathered.view(-1), mine.view(-1))
    merged = sharding.merge_keys_host(gathered.numpy().view(np.uint64), k)
    if rank == 0:
        np.save(os.path.join(out_dir, "merged.npy"), merged)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("metric", ["cosine", "dot", "euclidean"])
def test_two_rank_shard_merge_matches_unsharded(tmp_path, oracle_mod, metric):
    n, d, Q, k = 5001, 48, 5, 10
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n, d, Q, k, metric, str(tmp_path)), nprocs=2, join=True)
    merged = np.load(tmp_path / "merged.npy")
    rows_all = oracle_mod.fill_uniform(n, d, 42)
    q = oracle_mod.fill_uniform(Q, d, 1337)
    oi, os_ = oracle_mod.knn_exact64(rows_all, q, k, metric)
    gi, gs = sharding.unpack_keys(merged, euclidean=(metric == "euclidean"))
    assert (gi == oi).all()
    assert np.allclose(gs, os_, rtol=1e-5, atol=1e-6)


def test_key_codec_roundtrip_and_order():
    rng = np.random.default_rng(0)
    s = np.concatenate([rng.standard_normal(1000).astype(np.float32), np.array([0.0, -0.0, np.inf, -np.inf, 1e-38, -1e-38], np.float32)])
    r = rng.integers(0, 2**32 - 2, s.size, dtype=np.uint64).astype(np.uint32)
    keys = sharding.pack_keys(s, r)
    rr, ss = sharding.unpack_keys(keys)
    assert (rr == r).all() and (ss.view(np.uint32) == s.view(np.uint32)).all()
    # descending key order == (score desc, row asc)
    order = np.argsort(keys)[::-1]
    ref = np.lexsort((r, -s.astype(np.float64)))
    s_sorted = s[order]
    assert (np.diff(s_sorted.astype(np.float64)) <= 0).all()
    same = s[:, None] == s[None, :]
    tie = np.array([0.5, 0.5, 0.5], np.float32)
    kk = sharding.pack_keys(tie, np.array([7, 3, 5], np.uint32))
    assert np.argsort(kk)[::-1].tolist() == [1, 2, 0]
    assert (kk > 0).all()
    del ref, same


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 1000, 10_000_001):
        for g in (1, 2, 3, 8):
            rs = [sharding.shard_range(n, g, i) for i in range(g)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(g - 1))
