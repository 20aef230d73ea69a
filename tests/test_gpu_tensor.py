"""GPU parity tests of the tcgen05 / TMEM / TMA scan (path="tensor") against the fp64 oracle."""
import numpy as np
import pytest

from parity import check_parity

pytestmark = pytest.mark.gpu


def run_tc(oracle, n, d, Q, k, metric, seed=7, mutate=None):
    from nornicdb_b200.knn import KnnIndex
    rows = oracle.fill_uniform(n, d, seed)
    q = oracle.fill_uniform(Q, d, seed + 999)
    if mutate:
        rows, q = mutate(rows, q)
    ix = KnnIndex(d, metric=metric)
    try:
        ix.set_path("tensor")
        ix.upload(rows)
        gi, gs = ix.search(q, k)
        # the CUDA-core scan on the same index must agree (two independent kernels)
        ix.set_path("simt")
        si, ss = ix.search(q, k)
    finally:
        ix.release()
    oi, os_ = oracle.knn_exact64(rows, q, k, metric)
    swaps = check_parity(rows, q, k, metric, gi, gs, oi, os_)
    check_parity(rows, q, k, metric, si, ss, oi, os_)
    return swaps


@pytest.mark.parametrize("metric", ["cosine", "dot"])
def test_tensor_small(knn_lib, oracle_mod, metric):
    run_tc(oracle_mod, 5000, 256, 64, 10, metric)


@pytest.mark.parametrize("Q", [1, 17, 64, 65, 130])
def test_tensor_query_counts(knn_lib, oracle_mod, Q):
    run_tc(oracle_mod, 3000, 128, Q, 10, "cosine")


@pytest.mark.parametrize("n", [1, 100, 255, 256, 257, 511, 513, 40_000])
def test_tensor_ragged_rows(knn_lib, oracle_mod, n):
    run_tc(oracle_mod, n, 64, 8, 10, "dot")


@pytest.mark.parametrize("d", [32, 36, 100, 768, 1024, 1536])
def test_tensor_dims(knn_lib, oracle_mod, d):
    run_tc(oracle_mod, 4000, d, 32, 10, "cosine")


@pytest.mark.parametrize("k", [1, 100, 300, 767])
def test_tensor_k(knn_lib, oracle_mod, k):
    run_tc(oracle_mod, 30_000, 64, 16, k, "cosine")


def test_tensor_config2_shape_subsampled(knn_lib, oracle_mod):
    # BASELINE.json configs[1] shape (d=1024 Q=64 k=10 cosine) at an oracle-sized N
    assert run_tc(oracle_mod, 60_000, 1024, 64, 10, "cosine") <= 2


def test_tensor_zero_vectors_and_ties(knn_lib, oracle_mod):
    def mutate(rows, q):
        rows[3] = 0.0
        rows[300] = 0.0
        rows[700] = rows[10]  # duplicate row: tie broken by lowest index
        q[2] = 0.0
        return rows, q
    run_tc(oracle_mod, 2000, 96, 5, 2000 if False else 500, "cosine", mutate=mutate)
    from nornicdb_b200.knn import KnnIndex
    base = oracle_mod.fill_uniform(16, 64, 5)
    rows = np.tile(base, (64, 1))
    ix = KnnIndex(64, metric="dot")
    ix.set_path("tensor")
    ix.upload(rows)
    gi, gs = ix.search(base[3:4], 20)
    ix.release()
    assert gi[0].tolist() == [3 + 16 * j for j in range(20)]


def test_tensor_unsupported_shapes_fail_loudly(knn_lib, oracle_mod):
    from nornicdb_b200.knn import KnnIndex, KnnError
    ix = KnnIndex(30, metric="cosine")  # dim % 4 != 0 -> no TMA path
    ix.set_path("tensor")
    ix.upload(oracle_mod.fill_uniform(100, 30, 1))
    with pytest.raises(KnnError):
        ix.search(oracle_mod.fill_uniform(1, 30, 2), 5)
    ix.release()
